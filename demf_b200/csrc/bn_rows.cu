// Training-mode BatchNorm (+ ReLU) on point-major rows x (R, C) -- the layer that follows every 1x1
// convolution of the PointNet++ shared MLPs, VoteModule and prediction heads (mmcv ConvModule:
// conv -> BN -> ReLU; mmdet3d ops/pointnet_modules/point_sa_module.py, models/model_utils/vote_module.py).
//
// Upstream (and the layer-by-layer path here) runs, per layer, forward: collect_statistics, update_stats,
// transform_input, clamp (ReLU); backward: threshold (ReLU), backward_reduce, backward_elemt -- eight
// passes over activation tensors that are the largest objects of a training step (537 MB for the three
// layers of SA1 at batch 4). Here a layer is two launches forward and two backward, five passes:
//
//   bn_stats_kernel        per-channel sum / sum of squares: fp32 inside a block, double-precision atomics
//                          across blocks; the LAST block to finish (atomic ticket) writes mean and
//                          1/sqrt(var+eps), updates the running statistics (momentum, unbiased variance)
//                          and re-zeroes the accumulators for the next launch
//   bn_apply_kernel        y = [relu]((x - mean) * invstd * gamma + beta)
//   bn_bwd_reduce_kernel   with dy' = dy * (y > 0): sum dy', sum dy' (x - mean); the last block derives
//                          grad_gamma, grad_beta and the two per-channel coefficients of the input gradient
//   bn_bwd_apply_kernel    dx = (dy' - mean(dy') - (x - mean) * invstd^2 * mean(dy' (x - mean))) * invstd * gamma
//
// Layout: C/4 threads cover one row with float4 loads (C = 4 * 2^k <= 1024), 256/(C/4) rows per block
// iteration; every access is a full coalesced line. All HBM-bound: R*C*4 bytes per pass.
#include "common.cuh"

namespace demf {
namespace {

constexpr int kBnThreads = 256;
constexpr int kBnMaxBlocks = kNumSMs * 4;

__device__ __forceinline__ float4 f4_add(float4 a, float4 b) {
  return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}

// Reduce (a, b) over the row lanes of the block: threads with the same tx; result valid for ty == 0.
__device__ __forceinline__ void block_reduce_rows(float4& a, float4& b, int tx, int ty, int lanes, int cq,
                                                  float4* s_a, float4* s_b) {
  s_a[ty * cq + tx] = a;
  s_b[ty * cq + tx] = b;
  __syncthreads();
  for (int half = lanes >> 1; half > 0; half >>= 1) {
    if (ty < half) {
      s_a[ty * cq + tx] = f4_add(s_a[ty * cq + tx], s_a[(ty + half) * cq + tx]);
      s_b[ty * cq + tx] = f4_add(s_b[ty * cq + tx], s_b[(ty + half) * cq + tx]);
    }
    __syncthreads();
  }
  a = s_a[tx];
  b = s_b[tx];
}

// Returns true in every thread of the last block to pass (after all partials are visible).
__device__ __forceinline__ bool last_block_ticket(unsigned* counter) {
  __shared__ bool s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned t = atomicAdd(counter, 1u);
    s_last = t == gridDim.x - 1;
    if (s_last) *counter = 0;  // ready for the next launch (graph replays included)
  }
  __syncthreads();
  if (s_last) __threadfence();
  return s_last;
}

__device__ __forceinline__ void accumulate4(double* acc, int c4, float4 v) {
  atomicAdd(acc + c4, (double)v.x);
  atomicAdd(acc + c4 + 1, (double)v.y);
  atomicAdd(acc + c4 + 2, (double)v.z);
  atomicAdd(acc + c4 + 3, (double)v.w);
}

// accum: (2, C) doubles = (sum, sum of squares); zero on entry, zero again on exit
__global__ void __launch_bounds__(kBnThreads) bn_stats_kernel(const float* __restrict__ x, long R, int C,
                                                              float eps, float momentum, double* accum,
                                                              unsigned* counter, float* __restrict__ mean,
                                                              float* __restrict__ invstd,
                                                              float* __restrict__ running_mean,
                                                              float* __restrict__ running_var) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* s_a = reinterpret_cast<float4*>(smem_raw);
  float4* s_b = s_a + kBnThreads;
  const int cq = C >> 2, lanes = kBnThreads / cq;
  const int tx = threadIdx.x % cq, ty = threadIdx.x / cq;
  const long rows_per_block = (R + gridDim.x - 1) / gridDim.x;
  const long r0 = (long)blockIdx.x * rows_per_block, r1 = min(R, r0 + rows_per_block);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
#pragma unroll 4
  for (long r = r0 + ty; r < r1; r += lanes) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x + r * C) + tx);
    s = f4_add(s, v);
    q.x = fmaf(v.x, v.x, q.x); q.y = fmaf(v.y, v.y, q.y); q.z = fmaf(v.z, v.z, q.z); q.w = fmaf(v.w, v.w, q.w);
  }
  block_reduce_rows(s, q, tx, ty, lanes, cq, s_a, s_b);
  if (ty == 0) {
    accumulate4(accum, tx * 4, s);
    accumulate4(accum + C, tx * 4, q);
  }
  if (!last_block_ticket(counter)) return;
  for (int c = threadIdx.x; c < C; c += kBnThreads) {
    const double ds = __ldcg(accum + c), dq = __ldcg(accum + C + c);
    accum[c] = 0.0;
    accum[C + c] = 0.0;
    const double m = ds / (double)R;
    double var = dq / (double)R - m * m;
    if (var < 0.0) var = 0.0;
    mean[c] = (float)m;
    invstd[c] = 1.0f / sqrtf((float)var + eps);
    if (running_mean) {
      const double unbiased = R > 1 ? var * ((double)R / (double)(R - 1)) : var;
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)m;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
  }
}

__global__ void __launch_bounds__(kBnThreads) bn_apply_kernel(const float* __restrict__ x, long total4, int cq,
                                                              const float* __restrict__ mean,
                                                              const float* __restrict__ invstd,
                                                              const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, int relu,
                                                              float* __restrict__ y) {
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total4; e += (long)gridDim.x * blockDim.x) {
    const int c = (int)(e % cq);
    const float4 v = __ldg(reinterpret_cast<const float4*>(x) + e);
    const float4 m = __ldg(reinterpret_cast<const float4*>(mean) + c);
    const float4 is = __ldg(reinterpret_cast<const float4*>(invstd) + c);
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c);
    const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + c);
    float4 o;
    o.x = (v.x - m.x) * is.x * g.x + b.x;
    o.y = (v.y - m.y) * is.y * g.y + b.y;
    o.z = (v.z - m.z) * is.z * g.z + b.z;
    o.w = (v.w - m.w) * is.w * g.w + b.w;
    if (relu) {
      o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
    }
    reinterpret_cast<float4*>(y)[e] = o;
  }
}

// coef: (2, C) = mean(dy'), invstd^2 * mean(dy' * (x - mean))
__global__ void __launch_bounds__(kBnThreads) bn_bwd_reduce_kernel(
    const float* __restrict__ dy, const float* __restrict__ y, const float* __restrict__ x, long R, int C,
    const float* __restrict__ mean, const float* __restrict__ invstd, int relu, double* accum,
    unsigned* counter, float* __restrict__ grad_gamma, float* __restrict__ grad_beta, float* __restrict__ coef) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* s_a = reinterpret_cast<float4*>(smem_raw);
  float4* s_b = s_a + kBnThreads;
  const int cq = C >> 2, lanes = kBnThreads / cq;
  const int tx = threadIdx.x % cq, ty = threadIdx.x / cq;
  const long rows_per_block = (R + gridDim.x - 1) / gridDim.x;
  const long r0 = (long)blockIdx.x * rows_per_block, r1 = min(R, r0 + rows_per_block);
  const float4 m = __ldg(reinterpret_cast<const float4*>(mean) + tx);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
#pragma unroll 2
  for (long r = r0 + ty; r < r1; r += lanes) {
    float4 g = __ldg(reinterpret_cast<const float4*>(dy + r * C) + tx);
    if (relu) {
      const float4 o = __ldg(reinterpret_cast<const float4*>(y + r * C) + tx);
      g.x = o.x > 0.f ? g.x : 0.f; g.y = o.y > 0.f ? g.y : 0.f;
      g.z = o.z > 0.f ? g.z : 0.f; g.w = o.w > 0.f ? g.w : 0.f;
    }
    const float4 v = __ldg(reinterpret_cast<const float4*>(x + r * C) + tx);
    s = f4_add(s, g);
    q.x = fmaf(g.x, v.x - m.x, q.x); q.y = fmaf(g.y, v.y - m.y, q.y);
    q.z = fmaf(g.z, v.z - m.z, q.z); q.w = fmaf(g.w, v.w - m.w, q.w);
  }
  block_reduce_rows(s, q, tx, ty, lanes, cq, s_a, s_b);
  if (ty == 0) {
    accumulate4(accum, tx * 4, s);
    accumulate4(accum + C, tx * 4, q);
  }
  if (!last_block_ticket(counter)) return;
  for (int c = threadIdx.x; c < C; c += kBnThreads) {
    const double ds = __ldcg(accum + c), dq = __ldcg(accum + C + c);
    accum[c] = 0.0;
    accum[C + c] = 0.0;
    const float is = invstd[c];
    grad_beta[c] = (float)ds;
    grad_gamma[c] = (float)(dq * (double)is);
    coef[c] = (float)(ds / (double)R);
    coef[C + c] = (float)(dq / (double)R * (double)is * (double)is);
  }
}

__global__ void __launch_bounds__(kBnThreads) bn_bwd_apply_kernel(
    const float* __restrict__ dy, const float* __restrict__ y, const float* __restrict__ x, long total4, int cq,
    const float* __restrict__ mean, const float* __restrict__ invstd, const float* __restrict__ gamma,
    const float* __restrict__ coef, int relu, float* __restrict__ dx) {
  const int C = cq * 4;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total4; e += (long)gridDim.x * blockDim.x) {
    const int c = (int)(e % cq);
    float4 g = __ldg(reinterpret_cast<const float4*>(dy) + e);
    if (relu) {
      const float4 o = __ldg(reinterpret_cast<const float4*>(y) + e);
      g.x = o.x > 0.f ? g.x : 0.f; g.y = o.y > 0.f ? g.y : 0.f;
      g.z = o.z > 0.f ? g.z : 0.f; g.w = o.w > 0.f ? g.w : 0.f;
    }
    const float4 v = __ldg(reinterpret_cast<const float4*>(x) + e);
    const float4 m = __ldg(reinterpret_cast<const float4*>(mean) + c);
    const float4 is = __ldg(reinterpret_cast<const float4*>(invstd) + c);
    const float4 w = __ldg(reinterpret_cast<const float4*>(gamma) + c);
    const float4 k1 = __ldg(reinterpret_cast<const float4*>(coef) + c);
    const float4 k2 = __ldg(reinterpret_cast<const float4*>(coef + C) + c);
    float4 o;
    o.x = (g.x - k1.x - (v.x - m.x) * k2.x) * is.x * w.x;
    o.y = (g.y - k1.y - (v.y - m.y) * k2.y) * is.y * w.y;
    o.z = (g.z - k1.z - (v.z - m.z) * k2.z) * is.z * w.z;
    o.w = (g.w - k1.w - (v.w - m.w) * k2.w) * is.w * w.w;
    reinterpret_cast<float4*>(dx)[e] = o;
  }
}

// ---- last layer of a set-abstraction MLP in training: BatchNorm + ReLU + max over the ns rows of a centre ----
// pooled[m,c] = max_r relu(bn(x[m*ns + r, c])), arg[m,c] = first r attaining it. The normalised (R,C) tensor is
// never written: backward needs x, the pooled value (ReLU mask) and arg only.
__global__ void __launch_bounds__(kBnThreads) bn_apply_max_kernel(const float* __restrict__ x, long M, int ns, int cq,
                                                                  const float* __restrict__ mean,
                                                                  const float* __restrict__ invstd,
                                                                  const float* __restrict__ gamma,
                                                                  const float* __restrict__ beta,
                                                                  float* __restrict__ pooled,
                                                                  uint8_t* __restrict__ arg) {
  const long total = M * cq;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const int c = (int)(e % cq);
    const long m = e / cq;
    const float4 mu = __ldg(reinterpret_cast<const float4*>(mean) + c);
    const float4 is = __ldg(reinterpret_cast<const float4*>(invstd) + c);
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c);
    const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + c);
    float4 best = make_float4(-1.f, -1.f, -1.f, -1.f);  // below every ReLU output
    uchar4 at = make_uchar4(0, 0, 0, 0);
    const float4* row = reinterpret_cast<const float4*>(x) + (m * ns) * cq + c;
#pragma unroll 4
    for (int r = 0; r < ns; ++r) {
      const float4 v = __ldg(row + (long)r * cq);
      const float ox = fmaxf((v.x - mu.x) * is.x * g.x + b.x, 0.f), oy = fmaxf((v.y - mu.y) * is.y * g.y + b.y, 0.f),
                  oz = fmaxf((v.z - mu.z) * is.z * g.z + b.z, 0.f), ow = fmaxf((v.w - mu.w) * is.w * g.w + b.w, 0.f);
      if (ox > best.x) { best.x = ox; at.x = (unsigned char)r; }
      if (oy > best.y) { best.y = oy; at.y = (unsigned char)r; }
      if (oz > best.z) { best.z = oz; at.z = (unsigned char)r; }
      if (ow > best.w) { best.w = ow; at.w = (unsigned char)r; }
    }
    reinterpret_cast<float4*>(pooled)[e] = best;
    reinterpret_cast<uchar4*>(arg)[e] = at;
  }
}

// dy is nonzero only at (m*ns + arg[m,c], c) and only where pooled > 0: the two channel sums come from M*C
// elements and as many gathered x values instead of two passes over (R,C).
__global__ void __launch_bounds__(kBnThreads) bn_max_bwd_reduce_kernel(
    const float* __restrict__ gp, const float* __restrict__ pooled, const uint8_t* __restrict__ arg,
    const float* __restrict__ x, long M, int ns, int C, const float* __restrict__ mean,
    const float* __restrict__ invstd, double* accum, unsigned* counter, float* __restrict__ grad_gamma,
    float* __restrict__ grad_beta, float* __restrict__ coef) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* s_a = reinterpret_cast<float4*>(smem_raw);
  float4* s_b = s_a + kBnThreads;
  const int cq = C >> 2, lanes = kBnThreads / cq;
  const int tx = threadIdx.x % cq, ty = threadIdx.x / cq;
  const long per_block = (M + gridDim.x - 1) / gridDim.x;
  const long m0 = (long)blockIdx.x * per_block, m1 = min(M, m0 + per_block);
  const float4 mu = __ldg(reinterpret_cast<const float4*>(mean) + tx);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
  for (long m = m0 + ty; m < m1; m += lanes) {
    float4 g = __ldg(reinterpret_cast<const float4*>(gp + m * C) + tx);
    const float4 o = __ldg(reinterpret_cast<const float4*>(pooled + m * C) + tx);
    const uchar4 at = __ldg(reinterpret_cast<const uchar4*>(arg + m * C) + tx);
    g.x = o.x > 0.f ? g.x : 0.f; g.y = o.y > 0.f ? g.y : 0.f;
    g.z = o.z > 0.f ? g.z : 0.f; g.w = o.w > 0.f ? g.w : 0.f;
    const float* xb = x + (m * ns) * C + tx * 4;
    const float vx = __ldg(xb + (long)at.x * C), vy = __ldg(xb + (long)at.y * C + 1),
                vz = __ldg(xb + (long)at.z * C + 2), vw = __ldg(xb + (long)at.w * C + 3);
    s = f4_add(s, g);
    q.x = fmaf(g.x, vx - mu.x, q.x); q.y = fmaf(g.y, vy - mu.y, q.y);
    q.z = fmaf(g.z, vz - mu.z, q.z); q.w = fmaf(g.w, vw - mu.w, q.w);
  }
  block_reduce_rows(s, q, tx, ty, lanes, cq, s_a, s_b);
  if (ty == 0) {
    accumulate4(accum, tx * 4, s);
    accumulate4(accum + C, tx * 4, q);
  }
  if (!last_block_ticket(counter)) return;
  const double R = (double)M * ns;
  for (int c = threadIdx.x; c < C; c += kBnThreads) {
    const double ds = __ldcg(accum + c), dq = __ldcg(accum + C + c);
    accum[c] = 0.0;
    accum[C + c] = 0.0;
    const float is = invstd[c];
    grad_beta[c] = (float)ds;
    grad_gamma[c] = (float)(dq * (double)is);
    coef[c] = (float)(ds / R);
    coef[C + c] = (float)(dq / R * (double)is * (double)is);
  }
}

__global__ void __launch_bounds__(kBnThreads) bn_max_bwd_apply_kernel(
    const float* __restrict__ gp, const float* __restrict__ pooled, const uint8_t* __restrict__ arg,
    const float* __restrict__ x, long M, int ns, int cq, const float* __restrict__ mean,
    const float* __restrict__ invstd, const float* __restrict__ gamma, const float* __restrict__ coef,
    float* __restrict__ dx) {
  const int C = cq * 4;
  const long total = M * cq;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const int c = (int)(e % cq);
    const long m = e / cq;
    float4 g = __ldg(reinterpret_cast<const float4*>(gp) + e);
    const float4 o = __ldg(reinterpret_cast<const float4*>(pooled) + e);
    const uchar4 at = __ldg(reinterpret_cast<const uchar4*>(arg) + e);
    g.x = o.x > 0.f ? g.x : 0.f; g.y = o.y > 0.f ? g.y : 0.f;
    g.z = o.z > 0.f ? g.z : 0.f; g.w = o.w > 0.f ? g.w : 0.f;
    const float4 mu = __ldg(reinterpret_cast<const float4*>(mean) + c);
    const float4 is = __ldg(reinterpret_cast<const float4*>(invstd) + c);
    const float4 w = __ldg(reinterpret_cast<const float4*>(gamma) + c);
    const float4 k1 = __ldg(reinterpret_cast<const float4*>(coef) + c);
    const float4 k2 = __ldg(reinterpret_cast<const float4*>(coef + C) + c);
    const float4* xr = reinterpret_cast<const float4*>(x) + (m * ns) * cq + c;
    float4* dr = reinterpret_cast<float4*>(dx) + (m * ns) * cq + c;
#pragma unroll 4
    for (int r = 0; r < ns; ++r) {
      const float4 v = __ldg(xr + (long)r * cq);
      float4 d;
      d.x = ((r == at.x ? g.x : 0.f) - k1.x - (v.x - mu.x) * k2.x) * is.x * w.x;
      d.y = ((r == at.y ? g.y : 0.f) - k1.y - (v.y - mu.y) * k2.y) * is.y * w.y;
      d.z = ((r == at.z ? g.z : 0.f) - k1.z - (v.z - mu.z) * k2.z) * is.z * w.z;
      d.w = ((r == at.w ? g.w : 0.f) - k1.w - (v.w - mu.w) * k2.w) * is.w * w.w;
      dr[(long)r * cq] = d;
    }
  }
}

inline bool bn_shape_ok(int C) {
  const int cq = C / 4;
  return C % 4 == 0 && cq >= 1 && cq <= kBnThreads && (cq & (cq - 1)) == 0;
}

inline unsigned bn_blocks(long R, int C) {
  const long per_iter = kBnThreads / (C / 4);
  long blocks = (R + per_iter * 8 - 1) / (per_iter * 8);  // at least 8 iterations of work per block
  if (blocks > kBnMaxBlocks) blocks = kBnMaxBlocks;
  return (unsigned)(blocks < 1 ? 1 : blocks);
}

inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }


// out[c] += sum_r x[r, c]  (bias gradient of a Linear / 1x1 convolution): one launch instead of a library reduction
// plus an accumulate. Block = 32 columns x 8 row lanes; a warp reads 128 contiguous bytes of one row.
__global__ void __launch_bounds__(256) col_sum_add_kernel(const float* __restrict__ x, long R, int N, long ld,
                                                           long rows_per_block, float* __restrict__ out) {
  __shared__ float part[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const long r0 = (long)blockIdx.y * rows_per_block;
  const long r1 = min(R, r0 + rows_per_block);
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  if (c < N) {
    long r = r0 + threadIdx.y;
    for (; r + 24 < r1; r += 32) {
      a0 += __ldg(x + r * ld + c);
      a1 += __ldg(x + (r + 8) * ld + c);
      a2 += __ldg(x + (r + 16) * ld + c);
      a3 += __ldg(x + (r + 24) * ld + c);
    }
    for (; r < r1; r += 8) a0 += __ldg(x + r * ld + c);
  }
  part[threadIdx.y][threadIdx.x] = (a0 + a1) + (a2 + a3);
  __syncthreads();
  if (threadIdx.y == 0 && c < N) {
    float t = 0.f;
#pragma unroll
    for (int y = 0; y < 8; ++y) t += part[y][threadIdx.x];
    atomicAdd(out + c, t);
  }
}

}  // namespace
}  // namespace demf

using namespace demf;

extern "C" {

int demf_bn_rows_supported(int C) { return bn_shape_ok(C) ? 1 : 0; }

// bytes of the persistent accumulator block of one BatchNorm layer: (2, C) doubles + the ticket counter.
// It must be ZERO before the first call; every call leaves it zero again.
long demf_bn_rows_state_bytes(int C) { return (long)2 * C * 8 + 16; }

int demf_bn_rows_fwd(const float* x, long R, int C, const float* gamma, const float* beta, float eps,
                     float momentum, int relu, float* running_mean, float* running_var, void* state,
                     float* save_mean, float* save_invstd, float* y, void* stream) {
  DEMF_REQUIRE_PTR(x);
  DEMF_REQUIRE_PTR(gamma);
  DEMF_REQUIRE_PTR(beta);
  DEMF_REQUIRE_PTR(state);
  DEMF_REQUIRE_PTR(save_mean);
  DEMF_REQUIRE_PTR(save_invstd);
  DEMF_REQUIRE_PTR(y);
  double* accum = static_cast<double*>(state);
  unsigned* counter = reinterpret_cast<unsigned*>(accum + 2 * (long)C);
  DEMF_REQUIRE(R > 0 && C > 0, DEMF_E_SIZE);
  DEMF_REQUIRE(bn_shape_ok(C), DEMF_E_UNSUPPORTED);
  DEMF_REQUIRE((running_mean == nullptr) == (running_var == nullptr), DEMF_E_SIZE);
  DEMF_REQUIRE(al16(x) && al16(y) && al16(gamma) && al16(beta) && al16(save_mean) && al16(save_invstd) &&
                   al16(state),
               DEMF_E_UNSUPPORTED);
  cudaStream_t st = as_stream(stream);
  const unsigned blocks = bn_blocks(R, C);
  bn_stats_kernel<<<blocks, kBnThreads, kBnThreads * 32, st>>>(x, R, C, eps, momentum, accum, counter, save_mean,
                                                              save_invstd, running_mean, running_var);
  if (int rc = after_launch("bn_stats_kernel")) return rc;
  const long total4 = R * (C / 4);
  long ab = (total4 + kBnThreads - 1) / kBnThreads;
  if (ab > (long)kNumSMs * 16) ab = (long)kNumSMs * 16;
  bn_apply_kernel<<<(unsigned)ab, kBnThreads, 0, st>>>(x, total4, C / 4, save_mean, save_invstd, gamma, beta, relu,
                                                      y);
  return after_launch("bn_apply_kernel");
}

// The second half of demf_bn_rows_fwd / demf_bn_max_rows_fwd alone, for a layer whose statistics came out of the
// producing GEMM's epilogue (demf_gemm_rows_fwd with bn_state, then demf_bn_finalize): one pass over x.
int demf_bn_rows_apply(const float* x, long R, int C, const float* gamma, const float* beta, const float* mean,
                       const float* invstd, int relu, float* y, void* stream) {
  DEMF_REQUIRE_PTR(x);
  DEMF_REQUIRE_PTR(gamma);
  DEMF_REQUIRE_PTR(beta);
  DEMF_REQUIRE_PTR(mean);
  DEMF_REQUIRE_PTR(invstd);
  DEMF_REQUIRE_PTR(y);
  DEMF_REQUIRE(R > 0 && C > 0, DEMF_E_SIZE);
  DEMF_REQUIRE(bn_shape_ok(C), DEMF_E_UNSUPPORTED);
  DEMF_REQUIRE(al16(x) && al16(y) && al16(gamma) && al16(beta) && al16(mean) && al16(invstd), DEMF_E_UNSUPPORTED);
  const long total4 = R * (C / 4);
  long ab = (total4 + kBnThreads - 1) / kBnThreads;
  if (ab > (long)kNumSMs * 16) ab = (long)kNumSMs * 16;
  bn_apply_kernel<<<(unsigned)ab, kBnThreads, 0, as_stream(stream)>>>(x, total4, C / 4, mean, invstd, gamma, beta,
                                                                     relu, y);
  return after_launch("bn_apply_kernel");
}

// The input-gradient pass of demf_bn_rows_bwd alone: g = dL/dz with the ReLU mask already applied and the two
// reductions already finalised into `coef` (demf_gemm_rows_dgrad_bn + demf_bn_bwd_finalize).
int demf_bn_rows_bwd_apply(const float* g, const float* x, long R, int C, const float* gamma, const float* mean,
                           const float* invstd, const float* coef, float* grad_x, void* stream) {
  DEMF_REQUIRE_PTR(g);
  DEMF_REQUIRE_PTR(x);
  DEMF_REQUIRE_PTR(gamma);
  DEMF_REQUIRE_PTR(mean);
  DEMF_REQUIRE_PTR(invstd);
  DEMF_REQUIRE_PTR(coef);
  DEMF_REQUIRE_PTR(grad_x);
  DEMF_REQUIRE(R > 0 && C > 0, DEMF_E_SIZE);
  DEMF_REQUIRE(bn_shape_ok(C), DEMF_E_UNSUPPORTED);
  DEMF_REQUIRE(al16(g) && al16(x) && al16(grad_x) && al16(gamma) && al16(mean) && al16(invstd) && al16(coef),
               DEMF_E_UNSUPPORTED);
  const long total4 = R * (C / 4);
  long ab = (total4 + kBnThreads - 1) / kBnThreads;
  if (ab > (long)kNumSMs * 16) ab = (long)kNumSMs * 16;
  bn_bwd_apply_kernel<<<(unsigned)ab, kBnThreads, 0, as_stream(stream)>>>(g, nullptr, x, total4, C / 4, mean, invstd,
                                                                         gamma, coef, 0, grad_x);
  return after_launch("bn_bwd_apply_kernel");
}

int demf_bn_max_rows_apply(const float* x, long M, int ns, int C, const float* gamma, const float* beta,
                           const float* mean, const float* invstd, float* pooled, uint8_t* arg, void* stream) {
  DEMF_REQUIRE_PTR(x);
  DEMF_REQUIRE_PTR(gamma);
  DEMF_REQUIRE_PTR(beta);
  DEMF_REQUIRE_PTR(mean);
  DEMF_REQUIRE_PTR(invstd);
  DEMF_REQUIRE_PTR(pooled);
  DEMF_REQUIRE_PTR(arg);
  DEMF_REQUIRE(M > 0 && ns > 0 && ns <= 255 && C > 0, DEMF_E_SIZE);
  DEMF_REQUIRE(bn_shape_ok(C), DEMF_E_UNSUPPORTED);
  DEMF_REQUIRE(al16(x) && al16(pooled) && al16(gamma) && al16(beta) && al16(mean) && al16(invstd) &&
                   (reinterpret_cast<uintptr_t>(arg) & 3u) == 0,
               DEMF_E_UNSUPPORTED);
  const long total = M * (C / 4);
  long ab = (total + kBnThreads - 1) / kBnThreads;
  if (ab > (long)kNumSMs * 16) ab = (long)kNumSMs * 16;
  bn_apply_max_kernel<<<(unsigned)ab, kBnThreads, 0, as_stream(stream)>>>(x, M, ns, C / 4, mean, invstd, gamma, beta,
                                                                         pooled, arg);
  return after_launch("bn_apply_max_kernel");
}

int demf_bn_rows_bwd(const float* grad_y, const float* y, const float* x, long R, int C, const float* gamma,
                     const float* save_mean, const float* save_invstd, int relu, void* state, float* coef,
                     float* grad_x, float* grad_gamma, float* grad_beta, void* stream) {
  DEMF_REQUIRE_PTR(grad_y);
  DEMF_REQUIRE_PTR(x);
  DEMF_REQUIRE_PTR(gamma);
  DEMF_REQUIRE_PTR(save_mean);
  DEMF_REQUIRE_PTR(save_invstd);
  DEMF_REQUIRE_PTR(state);
  DEMF_REQUIRE_PTR(coef);
  DEMF_REQUIRE_PTR(grad_x);
  DEMF_REQUIRE_PTR(grad_gamma);
  DEMF_REQUIRE_PTR(grad_beta);
  if (relu) {
    DEMF_REQUIRE_PTR(y);
  }
  DEMF_REQUIRE(R > 0 && C > 0, DEMF_E_SIZE);
  DEMF_REQUIRE(bn_shape_ok(C), DEMF_E_UNSUPPORTED);
  DEMF_REQUIRE(al16(grad_y) && al16(y) && al16(x) && al16(grad_x) && al16(gamma) && al16(save_mean) &&
                   al16(save_invstd) && al16(coef) && al16(state),
               DEMF_E_UNSUPPORTED);
  double* accum = static_cast<double*>(state);
  unsigned* counter = reinterpret_cast<unsigned*>(accum + 2 * (long)C);
  cudaStream_t st = as_stream(stream);
  const unsigned blocks = bn_blocks(R, C);
  bn_bwd_reduce_kernel<<<blocks, kBnThreads, kBnThreads * 32, st>>>(grad_y, y, x, R, C, save_mean, save_invstd,
                                                                   relu, accum, counter, grad_gamma, grad_beta, coef);
  if (int rc = after_launch("bn_bwd_reduce_kernel")) return rc;
  const long total4 = R * (C / 4);
  long ab = (total4 + kBnThreads - 1) / kBnThreads;
  if (ab > (long)kNumSMs * 16) ab = (long)kNumSMs * 16;
  bn_bwd_apply_kernel<<<(unsigned)ab, kBnThreads, 0, st>>>(grad_y, y, x, total4, C / 4, save_mean, save_invstd,
                                                          gamma, coef, relu, grad_x);
  return after_launch("bn_bwd_apply_kernel");
}

int demf_bn_max_rows_fwd(const float* x, long M, int ns, int C, const float* gamma, const float* beta, float eps,
                         float momentum, float* running_mean, float* running_var, void* state, float* save_mean,
                         float* save_invstd, float* pooled, uint8_t* arg, void* stream) {
  DEMF_REQUIRE_PTR(x);
  DEMF_REQUIRE_PTR(gamma);
  DEMF_REQUIRE_PTR(beta);
  DEMF_REQUIRE_PTR(state);
  DEMF_REQUIRE_PTR(save_mean);
  DEMF_REQUIRE_PTR(save_invstd);
  DEMF_REQUIRE_PTR(pooled);
  DEMF_REQUIRE_PTR(arg);
  DEMF_REQUIRE(M > 0 && ns > 0 && ns <= 255 && C > 0, DEMF_E_SIZE);
  DEMF_REQUIRE(bn_shape_ok(C), DEMF_E_UNSUPPORTED);
  DEMF_REQUIRE((running_mean == nullptr) == (running_var == nullptr), DEMF_E_SIZE);
  DEMF_REQUIRE(al16(x) && al16(pooled) && al16(gamma) && al16(beta) && al16(save_mean) && al16(save_invstd) &&
                   al16(state) && (reinterpret_cast<uintptr_t>(arg) & 3u) == 0,
               DEMF_E_UNSUPPORTED);
  cudaStream_t st = as_stream(stream);
  double* accum = static_cast<double*>(state);
  unsigned* counter = reinterpret_cast<unsigned*>(accum + 2 * (long)C);
  const long R = M * ns;
  bn_stats_kernel<<<bn_blocks(R, C), kBnThreads, kBnThreads * 32, st>>>(x, R, C, eps, momentum, accum, counter,
                                                                      save_mean, save_invstd, running_mean,
                                                                      running_var);
  if (int rc = after_launch("bn_stats_kernel")) return rc;
  const long total = M * (C / 4);
  long ab = (total + kBnThreads - 1) / kBnThreads;
  if (ab > (long)kNumSMs * 16) ab = (long)kNumSMs * 16;
  bn_apply_max_kernel<<<(unsigned)ab, kBnThreads, 0, st>>>(x, M, ns, C / 4, save_mean, save_invstd, gamma, beta,
                                                          pooled, arg);
  return after_launch("bn_apply_max_kernel");
}

int demf_bn_max_rows_bwd(const float* grad_pooled, const float* pooled, const uint8_t* arg, const float* x, long M,
                         int ns, int C, const float* gamma, const float* save_mean, const float* save_invstd,
                         void* state, float* coef, float* grad_x, float* grad_gamma, float* grad_beta,
                         void* stream) {
  DEMF_REQUIRE_PTR(grad_pooled);
  DEMF_REQUIRE_PTR(pooled);
  DEMF_REQUIRE_PTR(arg);
  DEMF_REQUIRE_PTR(x);
  DEMF_REQUIRE_PTR(gamma);
  DEMF_REQUIRE_PTR(save_mean);
  DEMF_REQUIRE_PTR(save_invstd);
  DEMF_REQUIRE_PTR(state);
  DEMF_REQUIRE_PTR(coef);
  DEMF_REQUIRE_PTR(grad_x);
  DEMF_REQUIRE_PTR(grad_gamma);
  DEMF_REQUIRE_PTR(grad_beta);
  DEMF_REQUIRE(M > 0 && ns > 0 && ns <= 255 && C > 0, DEMF_E_SIZE);
  DEMF_REQUIRE(bn_shape_ok(C), DEMF_E_UNSUPPORTED);
  DEMF_REQUIRE(al16(grad_pooled) && al16(pooled) && al16(x) && al16(grad_x) && al16(gamma) && al16(save_mean) &&
                   al16(save_invstd) && al16(coef) && al16(state) && (reinterpret_cast<uintptr_t>(arg) & 3u) == 0,
               DEMF_E_UNSUPPORTED);
  cudaStream_t st = as_stream(stream);
  double* accum = static_cast<double*>(state);
  unsigned* counter = reinterpret_cast<unsigned*>(accum + 2 * (long)C);
  bn_max_bwd_reduce_kernel<<<bn_blocks(M, C), kBnThreads, kBnThreads * 32, st>>>(
      grad_pooled, pooled, arg, x, M, ns, C, save_mean, save_invstd, accum, counter, grad_gamma, grad_beta, coef);
  if (int rc = after_launch("bn_max_bwd_reduce_kernel")) return rc;
  const long total = M * (C / 4);
  long ab = (total + kBnThreads - 1) / kBnThreads;
  if (ab > (long)kNumSMs * 16) ab = (long)kNumSMs * 16;
  bn_max_bwd_apply_kernel<<<(unsigned)ab, kBnThreads, 0, st>>>(grad_pooled, pooled, arg, x, M, ns, C / 4, save_mean,
                                                              save_invstd, gamma, coef, grad_x);
  return after_launch("bn_max_bwd_apply_kernel");
}

/* out[c] += sum over the R rows of x (R, N) with row stride ld floats: the bias gradient of a Linear / 1x1
 * convolution accumulated in place (replaces mmcv's / autograd's `grad.sum(0)` + accumulate). */
int demf_col_sum_add(const float* x, long R, int N, long ld, float* out, void* stream) {
  DEMF_REQUIRE_PTR(x);
  DEMF_REQUIRE_PTR(out);
  DEMF_REQUIRE(R > 0 && N > 0 && ld >= N, DEMF_E_SIZE);
  const int gx = (N + 31) / 32;
  long gy = (2L * kNumSMs + gx - 1) / gx;
  if (gy > (R + 63) / 64) gy = (R + 63) / 64;
  if (gy < 1) gy = 1;
  const long rows_per_block = (R + gy - 1) / gy;
  col_sum_add_kernel<<<dim3((unsigned)gx, (unsigned)gy), dim3(32, 8), 0, as_stream(stream)>>>(x, R, N, ld,
                                                                                          rows_per_block, out);
  return after_launch("col_sum_add_kernel");
}

}  // extern "C"
