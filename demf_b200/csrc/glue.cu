// Small fused kernels for the glue between the big ops of an eval forward. Each one replaces a
// chain of 6-12 tiny library launches (index conversions, element-wise arithmetic, concatenations):
// with several forwards in flight a step is bound by the NUMBER of launches the GPU front end can
// retire (~2 us per kernel node), not by their work.
//
//   chain_indices_kernel    : PointNet2SASSG's `sa_indices[i+1] = gather(sa_indices[i], 1, idx.long())`
//                             for every level (mmdet3d models/backbones/pointnet2_sa_ssg.py), one launch
//   interp_cat_rows_kernel  : PointFPModule's  sqrt -> 1/(d+1e-8) -> normalise -> three_interpolate ->
//                             cat([interpolated, skip])  (mmdet3d ops/pointnet_modules/point_fp_module.py)
//   decode_boxes_kernel     : softmax(obj)[..., -1], softmax(sem), argmax / gather / class2angle /
//                             wrap / cat of DeMFClassAgnosticBBoxCoder.decode
//                             (demf/core/bbox/coders/class_agnostic_bbox_coder.py:168-194) for one
//                             prediction stage
//   bias_layer_norm_rows_kernel : LayerNorm(x [+ bias] [+ residual]) over rows, one warp per row held in
//                             registers: the `dropout(out) + identity` add and the LayerNorm that follow
//                             every attention / FFN block of a post-norm transformer layer (mmcv
//                             BaseTransformerLayer) in one pass -- read once, write once
#include "common.cuh"

namespace demf {
namespace {

struct ChainArgs {
  const int32_t* idx[4];
  int64_t* out[4];
  int m[4];
  int levels;
};

// One CTA per scene. out[0] = idx[0]; out[l][i] = out[l-1][idx[l][i]].
__global__ void __launch_bounds__(1024) chain_indices_kernel(const ChainArgs a) {
  const int b = blockIdx.x;
  for (int l = 0; l < a.levels; ++l) {
    const int32_t* idx = a.idx[l] + (long)b * a.m[l];
    int64_t* out = a.out[l] + (long)b * a.m[l];
    const int64_t* prev = l ? a.out[l - 1] + (long)b * a.m[l - 1] : nullptr;
    for (int i = threadIdx.x; i < a.m[l]; i += blockDim.x) {
      const int k = __ldg(idx + i);
      out[i] = prev ? prev[k] : (int64_t)k;
    }
    __syncthreads();  // this CTA's writes of level l are visible to its reads at level l+1
  }
}

// out[b,i,:] = [ sum_k w_k * src[b, idx[b,i,k], :]  |  skip[b,i,:] ],  w_k = r_k / (r_0+r_1+r_2),
// r_k = 1 / (sqrt(dist2_k) + 1e-8)   (the torch expression of PointFPModule, same operation order)
__global__ void __launch_bounds__(256) interp_cat_rows_kernel(const float* __restrict__ src,
                                                              const float* __restrict__ skip,
                                                              const int32_t* __restrict__ idx,
                                                              const float* __restrict__ dist2, int m, int n,
                                                              int C1q, int C2q, long total, float* __restrict__ out) {
  const int Cq = C1q + C2q;
  for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const int j = (int)(e % Cq);
    const long bi = e / Cq;
    float4 o;
    if (j < C1q) {
      const long b = bi / n;
      const float4* f = reinterpret_cast<const float4*>(src + b * (long)m * C1q * 4);
      const int i0 = __ldg(idx + bi * 3), i1 = __ldg(idx + bi * 3 + 1), i2 = __ldg(idx + bi * 3 + 2);
      const float r0 = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(__ldg(dist2 + bi * 3)), 1e-8f));
      const float r1 = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(__ldg(dist2 + bi * 3 + 1)), 1e-8f));
      const float r2 = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(__ldg(dist2 + bi * 3 + 2)), 1e-8f));
      const float norm = __fadd_rn(__fadd_rn(r0, r1), r2);
      const float w0 = __fdiv_rn(r0, norm), w1 = __fdiv_rn(r1, norm), w2 = __fdiv_rn(r2, norm);
      const float4 p0 = __ldg(f + (long)i0 * C1q + j), p1 = __ldg(f + (long)i1 * C1q + j),
                   p2 = __ldg(f + (long)i2 * C1q + j);
      o.x = __fmaf_rn(w2, p2.x, __fmaf_rn(w0, p0.x, __fmul_rn(w1, p1.x)));
      o.y = __fmaf_rn(w2, p2.y, __fmaf_rn(w0, p0.y, __fmul_rn(w1, p1.y)));
      o.z = __fmaf_rn(w2, p2.z, __fmaf_rn(w0, p0.z, __fmul_rn(w1, p1.z)));
      o.w = __fmaf_rn(w2, p2.w, __fmaf_rn(w0, p0.w, __fmul_rn(w1, p1.w)));
    } else {
      o = __ldg(reinterpret_cast<const float4*>(skip) + bi * C2q + (j - C1q));
    }
    reinterpret_cast<float4*>(out)[e] = o;
  }
}

struct DecodeArgs {
  const float *center, *size, *dir_class, *dir_res, *obj, *sem;
  int s_center, s_size, s_dir_class, s_dir_res, s_obj, s_sem;  // row strides in floats
  int Q, bins, classes;
  int out_rows, out_offset;  // rows per scene of the output tensors and this stage's first row
  long total;                // B * Q
  float* box;                // (B, out_rows, 7)
  float* obj_prob;           // (B, out_rows)
  float* sem_prob;           // (B, out_rows, classes)
};

__global__ void __launch_bounds__(256) decode_boxes_kernel(const DecodeArgs a) {
  const float two_pi = 6.283185307179586f, pi = 3.141592653589793f;
  for (long r = blockIdx.x * (long)blockDim.x + threadIdx.x; r < a.total; r += (long)gridDim.x * blockDim.x) {
    const long b = r / a.Q;
    const long q = r - b * a.Q;
    const long o = b * a.out_rows + a.out_offset + q;
    // heading: first arg-max bin, its residual, class2angle, wrap to (-pi, pi], then python % 2pi
    const float* dc = a.dir_class + r * a.s_dir_class;
    int best = 0;
    float bv = __ldg(dc);
    for (int k = 1; k < a.bins; ++k) {
      const float v = __ldg(dc + k);
      if (v > bv) {
        bv = v;
        best = k;
      }
    }
    const float per = (float)(6.283185307179586 / (double)a.bins);
    float ang = __fadd_rn(__fmul_rn((float)best, per), __ldg(a.dir_res + r * a.s_dir_res + best));
    if (ang > pi) ang = __fsub_rn(ang, two_pi);
    float rem = fmodf(ang, two_pi);
    if (rem != 0.f && rem < 0.f) rem = __fadd_rn(rem, two_pi);
    float* box = a.box + o * 7;
    const float* c = a.center + r * a.s_center;
    const float* s = a.size + r * a.s_size;
    box[0] = __ldg(c);
    box[1] = __ldg(c + 1);
    box[2] = __ldg(c + 2);
    box[3] = __ldg(s);
    box[4] = __ldg(s + 1);
    box[5] = __ldg(s + 2);
    box[6] = rem;
    // objectness: softmax over 2 logits, probability of the last one
    const float* ob = a.obj + r * a.s_obj;
    const float o0 = __ldg(ob), o1 = __ldg(ob + 1);
    const float om = fmaxf(o0, o1);
    const float e0 = expf(o0 - om), e1 = expf(o1 - om);
    a.obj_prob[o] = e1 / (e0 + e1);
    // semantic scores: softmax over `classes` logits
    const float* sm = a.sem + r * a.s_sem;
    float mx = __ldg(sm);
    for (int k = 1; k < a.classes; ++k) mx = fmaxf(mx, __ldg(sm + k));
    float sum = 0.f;
    for (int k = 0; k < a.classes; ++k) sum += expf(__ldg(sm + k) - mx);
    float* sp = a.sem_prob + o * a.classes;
    for (int k = 0; k < a.classes; ++k) sp[k] = expf(__ldg(sm + k) - mx) / sum;
  }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline unsigned blocks_for(long total, int threads) {
  long blocks = (total + threads - 1) / threads;
  if (blocks > (long)kNumSMs * 16) blocks = (long)kNumSMs * 16;
  return (unsigned)(blocks < 1 ? 1 : blocks);
}

// VoteModule tail (mmdet3d models/model_utils/vote_module.py forward, vote_per_seed = 1): from the
// conv_out rows votes (R, ldv) = [offset(3) | residual(C)] form
//   offset' = clamp(offset, +-range)      vote_xyz = seed_xyz + offset'
//   vote_feat = (seed_feat + residual) / ||seed_feat + residual||_2      (norm_feats)
// one warp per row, the row in registers (C = 128 * kVec).
template <int kVec>
__global__ void __launch_bounds__(256) vote_tail_kernel(const float* __restrict__ votes, int ldv,
                                                        const float* __restrict__ seed_xyz,
                                                        const float* __restrict__ seed_rows, long rows,
                                                        float rx, float ry, float rz, int norm_feats,
                                                        float* __restrict__ vote_xyz, float* __restrict__ offset,
                                                        float* __restrict__ vote_rows) {
  constexpr int C = 128 * kVec;
  const unsigned lane = lane_id();
  const long warps = (long)gridDim.x * (blockDim.x >> 5);
  for (long r = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < rows; r += warps) {
    const float* v = votes + r * ldv;
    if (lane < 3) {
      const float lim = lane == 0 ? rx : (lane == 1 ? ry : rz);
      float o = v[lane];
      if (lim >= 0.f) o = fminf(fmaxf(o, -lim), lim);
      offset[r * 3 + lane] = o;
      vote_xyz[r * 3 + lane] = __fadd_rn(seed_xyz[r * 3 + lane], o);
    }
    float x[kVec * 4];
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < kVec; ++i) {
      const int c = (i * 32 + lane) * 4;
      const float4 s4 = *reinterpret_cast<const float4*>(seed_rows + r * C + c);
      // the residual starts at column 3 of the vote row: not 16-byte aligned, scalar loads
      x[i * 4 + 0] = __fadd_rn(s4.x, v[3 + c]);
      x[i * 4 + 1] = __fadd_rn(s4.y, v[4 + c]);
      x[i * 4 + 2] = __fadd_rn(s4.z, v[5 + c]);
      x[i * 4 + 3] = __fadd_rn(s4.w, v[6 + c]);
      sq += (x[i * 4] * x[i * 4] + x[i * 4 + 1] * x[i * 4 + 1]) + (x[i * 4 + 2] * x[i * 4 + 2] + x[i * 4 + 3] * x[i * 4 + 3]);
    }
    float scale = 1.f;
    if (norm_feats) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
      scale = sqrtf(sq);
    }
#pragma unroll
    for (int i = 0; i < kVec; ++i) {
      const int c = (i * 32 + lane) * 4;
      float4 y;
      y.x = norm_feats ? __fdiv_rn(x[i * 4 + 0], scale) : x[i * 4 + 0];
      y.y = norm_feats ? __fdiv_rn(x[i * 4 + 1], scale) : x[i * 4 + 1];
      y.z = norm_feats ? __fdiv_rn(x[i * 4 + 2], scale) : x[i * 4 + 2];
      y.w = norm_feats ? __fdiv_rn(x[i * 4 + 3], scale) : x[i * 4 + 3];
      *reinterpret_cast<float4*>(vote_rows + r * C + c) = y;
    }
  }
}

// Proposal centres -> normalised image coordinates (DeMFVoteHead.get_reference_points,
// demf/modeling/heads/class_agnostic_vote_head.py:524-547, with the per-scene chain of inverse 3D
// augmentation, depth2img, 2D augmentation and normalisation folded on the host into mats (B,3,4) and
// affs (B,4) = (su, sv, ou, ov)):  h = M [x y z 1]^T;  uv = clamp((h.xy / h.z) * s + o, 0, 1).
// One thread per proposal, fp32 FMAs (the library path was a batched GEMM + four element-wise launches).
__global__ void __launch_bounds__(256) project_points_kernel(const float* __restrict__ xyz,
                                                             const float* __restrict__ mats,
                                                             const float* __restrict__ affs, int Q, long total,
                                                             float* __restrict__ out) {
  for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const long b = e / Q;
    const float* m = mats + b * 12;
    const float* a = affs + b * 4;
    const float x = __ldg(xyz + e * 3), y = __ldg(xyz + e * 3 + 1), z = __ldg(xyz + e * 3 + 2);
    const float hu = fmaf(z, __ldg(m + 2), fmaf(y, __ldg(m + 1), fmaf(x, __ldg(m + 0), __ldg(m + 3))));
    const float hv = fmaf(z, __ldg(m + 6), fmaf(y, __ldg(m + 5), fmaf(x, __ldg(m + 4), __ldg(m + 7))));
    const float hw = fmaf(z, __ldg(m + 10), fmaf(y, __ldg(m + 9), fmaf(x, __ldg(m + 8), __ldg(m + 11))));
    const float u = fmaf(__fdiv_rn(hu, hw), __ldg(a + 0), __ldg(a + 2));
    const float v = fmaf(__fdiv_rn(hv, hw), __ldg(a + 1), __ldg(a + 3));
    // torch.clamp semantics: NaN stays NaN
    out[e * 2] = u != u ? u : fminf(fmaxf(u, 0.f), 1.f);
    out[e * 2 + 1] = v != v ? v : fminf(fmaxf(v, 0.f), 1.f);
  }
}

// Pyramid levels (B,C,H_l*W_l) -> token rows (B,S,C), all levels in one launch: 32x32 tiles through
// shared memory so that both the reads (along the pixels) and the writes (along the channels) are
// coalesced 128-byte lines. blockIdx.x enumerates (level, pixel tile), blockIdx.y channel tiles,
// blockIdx.z the batch.
struct LevelsArgs {
  const float* src[8];
  int hw[8];
  int start[8];       // first token row of the level
  int tile_start[9];  // prefix sum of ceil(hw/32) over the levels
  int levels;
};

__global__ void __launch_bounds__(256) levels_to_rows_kernel(const LevelsArgs a, int C, int S,
                                                             float* __restrict__ out) {
  __shared__ float tile[32][33];
  int l = 0;
  while (l + 1 < a.levels && (int)blockIdx.x >= a.tile_start[l + 1]) ++l;
  const int p0 = ((int)blockIdx.x - a.tile_start[l]) * 32;
  const int c0 = blockIdx.y * 32;
  const int b = blockIdx.z;
  const int hw = a.hw[l];
  const float* src = a.src[l] + (long)b * C * hw;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = c0 + ty + i * 8, p = p0 + tx;
    if (c < C && p < hw) tile[ty + i * 8][tx] = __ldg(src + (long)c * hw + p);
  }
  __syncthreads();
  float* dst = out + ((long)b * S + a.start[l]) * C;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int p = p0 + ty + i * 8, c = c0 + tx;
    if (c < C && p < hw) dst[(long)p * C + c] = tile[tx][ty + i * 8];
  }
}

// y[r,:] = (t - mean(t)) * rsqrt(var(t) + eps) * gamma + beta,  t = x[r,:] (+ bias) (+ res[r,:]).
// kVec float4 per lane: C = 128 * kVec. Two-pass statistics on the register copy (biased variance).
template <int kVec>
__global__ void __launch_bounds__(256) bias_layer_norm_rows_kernel(
    const float* __restrict__ x, const float* __restrict__ bias, const float* __restrict__ res,
    const float* __restrict__ gamma, const float* __restrict__ beta, long rows, float eps,
    float* __restrict__ out, const float* __restrict__ post_add, float* __restrict__ out2) {
  constexpr int C = 128 * kVec;
  const unsigned lane = lane_id();
  const long warps = (long)gridDim.x * (blockDim.x >> 5);
  for (long r = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < rows; r += warps) {
    float4 v[kVec];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < kVec; ++i) {
      const int c = (i * 32 + lane) * 4;
      v[i] = *reinterpret_cast<const float4*>(x + r * C + c);
      if (bias) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(bias + c));
        v[i].x += b.x; v[i].y += b.y; v[i].z += b.z; v[i].w += b.w;
      }
      if (res) {
        const float4 q = *reinterpret_cast<const float4*>(res + r * C + c);
        v[i].x += q.x; v[i].y += q.y; v[i].z += q.z; v[i].w += q.w;
      }
      sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum * (1.f / C);
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < kVec; ++i) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      sq += (a * a + b * b) + (c * c + d * d);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq * (1.f / C) + eps);
#pragma unroll
    for (int i = 0; i < kVec; ++i) {
      const int c = (i * 32 + lane) * 4;
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
      const float4 b = __ldg(reinterpret_cast<const float4*>(beta + c));
      float4 y;
      y.x = (v[i].x - mean) * rstd * g.x + b.x;
      y.y = (v[i].y - mean) * rstd * g.y + b.y;
      y.z = (v[i].z - mean) * rstd * g.z + b.z;
      y.w = (v[i].w - mean) * rstd * g.w + b.w;
      *reinterpret_cast<float4*>(out + r * C + c) = y;
      if (out2) {  // the next attention's query = y + positional embedding, written alongside
        const float4 p = *reinterpret_cast<const float4*>(post_add + r * C + c);
        *reinterpret_cast<float4*>(out2 + r * C + c) = make_float4(y.x + p.x, y.y + p.y, y.z + p.z, y.w + p.w);
      }
    }
  }
}

}  // namespace
}  // namespace demf

using namespace demf;

extern "C" {

int demf_chain_indices(int B, int levels, const int32_t* idx0, int m0, const int32_t* idx1, int m1,
                       const int32_t* idx2, int m2, const int32_t* idx3, int m3, int64_t* out0,
                       int64_t* out1, int64_t* out2, int64_t* out3, void* stream) {
  DEMF_REQUIRE(B >= 0 && levels >= 1 && levels <= 4, DEMF_E_SIZE);
  ChainArgs a{};
  const int32_t* in[4] = {idx0, idx1, idx2, idx3};
  int64_t* out[4] = {out0, out1, out2, out3};
  const int m[4] = {m0, m1, m2, m3};
  for (int l = 0; l < levels; ++l) {
    DEMF_REQUIRE_PTR(in[l]);
    DEMF_REQUIRE_PTR(out[l]);
    DEMF_REQUIRE(m[l] > 0, DEMF_E_SIZE);
    a.idx[l] = in[l];
    a.out[l] = out[l];
    a.m[l] = m[l];
  }
  a.levels = levels;
  if (B == 0) return 0;
  chain_indices_kernel<<<B, 1024, 0, as_stream(stream)>>>(a);
  return after_launch("chain_indices_kernel");
}

int demf_interp_cat_rows_fwd(const float* src_rows, const float* skip_rows, const int32_t* idx,
                             const float* dist2, int B, int C1, int C2, int m, int n, float* out,
                             void* stream) {
  DEMF_REQUIRE_PTR(src_rows);
  DEMF_REQUIRE_PTR(idx);
  DEMF_REQUIRE_PTR(dist2);
  DEMF_REQUIRE_PTR(out);
  if (C2 > 0) DEMF_REQUIRE_PTR(skip_rows);
  DEMF_REQUIRE(B >= 0 && C1 > 0 && C2 >= 0 && m > 0 && n >= 0, DEMF_E_SIZE);
  DEMF_REQUIRE(C1 % 4 == 0 && C2 % 4 == 0 && aligned16(src_rows) && aligned16(out) &&
                   (C2 == 0 || aligned16(skip_rows)),
               DEMF_E_UNSUPPORTED);
  const long total = (long)B * n * ((C1 + C2) / 4);
  if (total == 0) return 0;
  interp_cat_rows_kernel<<<blocks_for(total, 256), 256, 0, as_stream(stream)>>>(
      src_rows, skip_rows, idx, dist2, m, n, C1 / 4, C2 / 4, total, out);
  return after_launch("interp_cat_rows_kernel");
}

int demf_decode_boxes(const float* center, int s_center, const float* size, int s_size,
                      const float* dir_class, int s_dir_class, const float* dir_res, int s_dir_res,
                      const float* obj, int s_obj, const float* sem, int s_sem, int B, int Q, int bins,
                      int classes, int out_rows, int out_offset, float* box, float* obj_prob,
                      float* sem_prob, void* stream) {
  DEMF_REQUIRE_PTR(center);
  DEMF_REQUIRE_PTR(size);
  DEMF_REQUIRE_PTR(dir_class);
  DEMF_REQUIRE_PTR(dir_res);
  DEMF_REQUIRE_PTR(obj);
  DEMF_REQUIRE_PTR(sem);
  DEMF_REQUIRE_PTR(box);
  DEMF_REQUIRE_PTR(obj_prob);
  DEMF_REQUIRE_PTR(sem_prob);
  DEMF_REQUIRE(B >= 0 && Q > 0 && bins > 0 && classes > 0 && out_offset >= 0 && out_offset + Q <= out_rows,
               DEMF_E_SIZE);
  if (B == 0) return 0;
  DecodeArgs a{center, size, dir_class, dir_res, obj, sem, s_center, s_size, s_dir_class, s_dir_res,
               s_obj, s_sem, Q, bins, classes, out_rows, out_offset, (long)B * Q, box, obj_prob, sem_prob};
  decode_boxes_kernel<<<blocks_for(a.total, 256), 256, 0, as_stream(stream)>>>(a);
  return after_launch("decode_boxes_kernel");
}

int demf_vote_tail(const float* votes, int ldv, const float* seed_xyz, const float* seed_rows, long rows, int C,
                   const float* xyz_range, int norm_feats, float* vote_xyz, float* offset, float* vote_rows,
                   void* stream) {
  DEMF_REQUIRE_PTR(votes);
  DEMF_REQUIRE_PTR(seed_xyz);
  DEMF_REQUIRE_PTR(seed_rows);
  DEMF_REQUIRE_PTR(vote_xyz);
  DEMF_REQUIRE_PTR(offset);
  DEMF_REQUIRE_PTR(vote_rows);
  DEMF_REQUIRE(rows >= 0 && C > 0 && ldv >= C + 3, DEMF_E_SIZE);
  DEMF_REQUIRE(C % 128 == 0 && C <= 512, DEMF_E_UNSUPPORTED);
  DEMF_REQUIRE(((reinterpret_cast<uintptr_t>(seed_rows) | reinterpret_cast<uintptr_t>(vote_rows)) & 15u) == 0,
               DEMF_E_UNSUPPORTED);
  if (rows == 0) return 0;
  // xyz_range: HOST pointer to 3 floats or NULL (no clamp)
  const float rx = xyz_range ? xyz_range[0] : -1.f, ry = xyz_range ? xyz_range[1] : -1.f,
              rz = xyz_range ? xyz_range[2] : -1.f;
  long blocks = (rows + 7) / 8;
  if (blocks > (long)kNumSMs * 16) blocks = (long)kNumSMs * 16;
  cudaStream_t st = as_stream(stream);
  switch (C / 128) {
#define DEMF_CASE(n)                                                                                          \
  case n:                                                                                                     \
    vote_tail_kernel<n><<<(unsigned)blocks, 256, 0, st>>>(votes, ldv, seed_xyz, seed_rows, rows, rx, ry, rz,   \
                                                          norm_feats, vote_xyz, offset, vote_rows);           \
    break;
    DEMF_CASE(1) DEMF_CASE(2) DEMF_CASE(3) DEMF_CASE(4)
#undef DEMF_CASE
  }
  return after_launch("vote_tail_kernel");
}

int demf_project_points(const float* xyz, const float* mats, const float* affs, int B, int Q, float* out,
                        void* stream) {
  DEMF_REQUIRE_PTR(xyz);
  DEMF_REQUIRE_PTR(mats);
  DEMF_REQUIRE_PTR(affs);
  DEMF_REQUIRE_PTR(out);
  DEMF_REQUIRE(B >= 0 && Q >= 0, DEMF_E_SIZE);
  const long total = (long)B * Q;
  if (total == 0) return 0;
  project_points_kernel<<<blocks_for(total, 256), 256, 0, as_stream(stream)>>>(xyz, mats, affs, Q, total, out);
  return after_launch("project_points_kernel");
}

int demf_levels_to_rows(const float* const* levels, const int* hw, int num_levels, int B, int C, float* out,
                        void* stream) {
  DEMF_REQUIRE_PTR(levels);
  DEMF_REQUIRE_PTR(hw);
  DEMF_REQUIRE_PTR(out);
  DEMF_REQUIRE(num_levels >= 1 && num_levels <= 8 && B >= 0 && C > 0, DEMF_E_SIZE);
  DEMF_REQUIRE(B <= 65535 && (C + 31) / 32 <= 65535, DEMF_E_SIZE);
  LevelsArgs a{};
  int S = 0, tiles = 0;
  for (int l = 0; l < num_levels; ++l) {
    DEMF_REQUIRE_PTR(levels[l]);
    DEMF_REQUIRE(hw[l] > 0, DEMF_E_SIZE);
    a.src[l] = levels[l];
    a.hw[l] = hw[l];
    a.start[l] = S;
    a.tile_start[l] = tiles;
    S += hw[l];
    tiles += (hw[l] + 31) / 32;
  }
  a.tile_start[num_levels] = tiles;
  a.levels = num_levels;
  if (B == 0) return 0;
  dim3 grid(tiles, (C + 31) / 32, B);
  levels_to_rows_kernel<<<grid, 256, 0, as_stream(stream)>>>(a, C, S, out);
  return after_launch("levels_to_rows_kernel");
}

int demf_bias_layer_norm_rows(const float* x, const float* bias, const float* residual, const float* gamma,
                              const float* beta, long rows, int C, float eps, float* out, const float* post_add,
                              float* out2, void* stream) {
  DEMF_REQUIRE((post_add == nullptr) == (out2 == nullptr), DEMF_E_SIZE);
  DEMF_REQUIRE(((reinterpret_cast<uintptr_t>(post_add) | reinterpret_cast<uintptr_t>(out2)) & 15u) == 0,
               DEMF_E_UNSUPPORTED);
  DEMF_REQUIRE_PTR(x);
  DEMF_REQUIRE_PTR(gamma);
  DEMF_REQUIRE_PTR(beta);
  DEMF_REQUIRE_PTR(out);
  DEMF_REQUIRE(rows >= 0 && C > 0, DEMF_E_SIZE);
  DEMF_REQUIRE(C % 128 == 0 && C <= 1024, DEMF_E_UNSUPPORTED);
  DEMF_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out) |
                 reinterpret_cast<uintptr_t>(bias) | reinterpret_cast<uintptr_t>(residual) |
                 reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(beta)) & 15u) == 0,
               DEMF_E_UNSUPPORTED);
  if (rows == 0) return 0;
  long blocks = (rows + 7) / 8;
  if (blocks > (long)kNumSMs * 16) blocks = (long)kNumSMs * 16;
  cudaStream_t st = as_stream(stream);
  switch (C / 128) {
#define DEMF_CASE(n)                                                                                      \
  case n:                                                                                                 \
    bias_layer_norm_rows_kernel<n><<<(unsigned)blocks, 256, 0, st>>>(x, bias, residual, gamma, beta, rows, \
                                                                     eps, out, post_add, out2);           \
    break;
    DEMF_CASE(1) DEMF_CASE(2) DEMF_CASE(3) DEMF_CASE(4) DEMF_CASE(5) DEMF_CASE(6) DEMF_CASE(7) DEMF_CASE(8)
#undef DEMF_CASE
  }
  return after_launch("bias_layer_norm_rows_kernel");
}

}  // extern "C"
