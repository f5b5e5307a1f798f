// Point-major ("row") layout kernels: the layout the B200 backbone actually runs on.
//
// Upstream keeps features channel-major, (B,C,N), so every grouping gather reads C scattered
// 4-byte words per neighbour. Here the backbone keeps features point-major, (B,N,C): a
// neighbour is ONE contiguous row, gathered with 16-byte loads, and the grouped tensor is
// emitted as GEMM-ready rows
//     out[b, m, s, :] = [ feat[b, idx, 0:C] | 0-pad to a multiple of 4 | (xyz[idx]-centre)*1/r | 0 ]
// of K = roundup(C,4)+4 floats (16-byte aligned rows; the 1x1-conv weight is permuted to this
// column order on the host side, demf_b200/mm/pointnet_modules.py). The index rows are the same
// "first nsample hits in index order" rows as demf_ball_query (ball_query.cu), produced in the
// same launch and kept in shared memory.
//
// Roofline: HBM. Algorithmic bytes per scene = N*12 (cloud) + M*12 + M*ns*4 (idx) +
// M*ns*C*4 (row gather, L2-resident source) + M*ns*K*4 (rows written).
#include "ball_grid.cuh"
#include "common.cuh"

namespace demf {
namespace {

constexpr int kWarps = 16;  // centres per CTA
constexpr int kThreads = kWarps * 32;
constexpr int kTile = 2048;  // cloud points per shared-memory tile (24 KB)

__device__ __forceinline__ void red_add_v4(float* addr, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}

// kQuery: run the ball query (idx is an output); otherwise idx is an input.
template <bool kQuery>
__global__ void __launch_bounds__(kThreads) group_rows_fwd_kernel(
    const float* __restrict__ xyz, const float* __restrict__ feat, const float* __restrict__ new_xyz,
    int N, int M, int C, float min_r2, float max_r2, float inv_radius, int ns, int normalize_xyz,
    const void* __restrict__ grid, int32_t* __restrict__ idx, float* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* tile = reinterpret_cast<float*>(smem_raw);
  int32_t* rows = reinterpret_cast<int32_t*>(smem_raw + (kQuery ? kTile * 3 * 4 : 0));

  const unsigned lane = lane_id();
  const int warp = threadIdx.x >> 5;
  const int b = blockIdx.y;
  const int m = blockIdx.x * kWarps + warp;
  const bool active = m < M;
  const float* cloud = xyz + (long)b * N * 3;

  float cx = 0.f, cy = 0.f, cz = 0.f;
  if (active) {
    const float* c = new_xyz + ((long)b * M + m) * 3;
    cx = __ldg(c + 0);
    cy = __ldg(c + 1);
    cz = __ldg(c + 2);
  }
  int32_t* row = rows + warp * ns;
  int32_t* grow = idx + ((long)b * M + (active ? m : 0)) * ns;

  if (kQuery && grid != nullptr) {
    // exact grid query (ball_grid.cuh): only the 3x3x3 cell neighbourhood of the centre is tested
    if (!active) return;
    int* scratch = reinterpret_cast<int*>(tile) + warp * (kGridCap + kGridHist);
    const BallGridView g = ball_grid_view(grid, b, N);
    ball_query_warp(g, cloud, N, cx, cy, cz, min_r2, max_r2, ns, row, scratch, scratch + kGridCap, lane);
    for (int l = lane; l < ns; l += 32) grow[l] = row[l];
  } else if (kQuery) {
    int cnt = 0, first = 0;
    bool done = !active;
    for (int base = 0; base < N; base += kTile) {
      const int npts = min(kTile, N - base);
      __syncthreads();
      for (int i = threadIdx.x; i < npts * 3; i += kThreads)
        tile[i] = __ldg(cloud + (long)base * 3 + i);
      __syncthreads();
      if (!done) {
        for (int j = 0; j < npts; j += 32) {
          const int p = j + lane;
          bool hit = false;
          if (p < npts) {
            const float d2 = sqdist(cx, cy, cz, tile[p * 3 + 0], tile[p * 3 + 1], tile[p * 3 + 2]);
            hit = (d2 == 0.f) || (d2 >= min_r2 && d2 < max_r2);
          }
          const unsigned ballot = __ballot_sync(0xffffffffu, hit);
          if (ballot) {
            if (cnt == 0) first = base + j + (__ffs(ballot) - 1);
            const int pos = cnt + __popc(ballot & ((1u << lane) - 1u));
            if (hit && pos < ns) row[pos] = base + p;
            cnt += __popc(ballot);
            if (cnt >= ns) {
              done = true;
              break;
            }
          }
        }
      }
      if (__syncthreads_and(done)) break;
    }
    if (!active) return;
    if (cnt > ns) cnt = ns;
    __syncwarp();
    for (int l = cnt + lane; l < ns; l += 32) row[l] = first;
    __syncwarp();
    for (int l = lane; l < ns; l += 32) grow[l] = row[l];
  } else {
    if (!active) return;
    for (int l = lane; l < ns; l += 32) row[l] = __ldg(grow + l);
    __syncwarp();
  }

  // ---- emit the ns GEMM rows of this centre: consecutive lanes write consecutive float4s
  const int Cp4 = (C + 3) >> 2;  // float4 slots of the feature part
  const int K4 = Cp4 + 1;        // + the xyz slot
  const float* fb = feat ? feat + (long)b * N * C : nullptr;
  float4* o4 = reinterpret_cast<float4*>(out + ((long)b * M + m) * ns * (K4 * 4));
  const bool vec = (C & 3) == 0;
  const float scale = normalize_xyz ? inv_radius : 1.f;
  const int total = ns * K4;
  for (int e = lane; e < total; e += 32) {
    const int s = e / K4;
    const int j = e - s * K4;
    const int k = row[s];
    float4 v;
    if (j < Cp4) {
      if (vec) {
        v = __ldg(reinterpret_cast<const float4*>(fb + (long)k * C) + j);
      } else {
        const float* f = fb + (long)k * C + j * 4;
        const int left = C - j * 4;
        v.x = __ldg(f);
        v.y = left > 1 ? __ldg(f + 1) : 0.f;
        v.z = left > 2 ? __ldg(f + 2) : 0.f;
        v.w = left > 3 ? __ldg(f + 3) : 0.f;
      }
    } else {
      const float* p = cloud + (long)k * 3;
      // separate roundings, as upstream's torch ops: (p - c), then * (1.0f / r)
      v.x = __fsub_rn(__ldg(p + 0), cx);
      v.y = __fsub_rn(__ldg(p + 1), cy);
      v.z = __fsub_rn(__ldg(p + 2), cz);
      if (normalize_xyz) {
        v.x = __fmul_rn(v.x, scale);
        v.y = __fmul_rn(v.y, scale);
        v.z = __fmul_rn(v.z, scale);
      }
      v.w = 0.f;
    }
    o4[e] = v;
  }
}

// One warp per centre: scatter-add of the row gradients.
__global__ void __launch_bounds__(kThreads) group_rows_bwd_kernel(
    const float* __restrict__ grad_out, const int32_t* __restrict__ idx, int N, int M, int C, int ns,
    float scale, float* __restrict__ grad_feat, float* __restrict__ grad_xyz,
    float* __restrict__ grad_centre) {
  const unsigned lane = lane_id();
  const int warp = threadIdx.x >> 5;
  const int b = blockIdx.y;
  const int m = blockIdx.x * kWarps + warp;
  if (m >= M) return;
  const int Cp4 = (C + 3) >> 2;
  const int K4 = Cp4 + 1;
  const int32_t* row = idx + ((long)b * M + m) * ns;
  const float4* g4 = reinterpret_cast<const float4*>(grad_out + ((long)b * M + m) * ns * (K4 * 4));
  float* gf = grad_feat ? grad_feat + (long)b * N * C : nullptr;
  float* gx = grad_xyz ? grad_xyz + (long)b * N * 3 : nullptr;
  const bool vec = (C & 3) == 0;
  float sx = 0.f, sy = 0.f, sz = 0.f;
  const int total = ns * K4;
  for (int e = lane; e < total; e += 32) {
    const int s = e / K4;
    const int j = e - s * K4;
    const int k = __ldg(row + s);
    const float4 g = __ldg(g4 + e);
    if (j < Cp4) {
      if (gf) {
        float* dst = gf + (long)k * C + j * 4;
        if (vec) {
          red_add_v4(dst, g);
        } else {
          const int left = C - j * 4;
          atomicAdd(dst, g.x);
          if (left > 1) atomicAdd(dst + 1, g.y);
          if (left > 2) atomicAdd(dst + 2, g.z);
          if (left > 3) atomicAdd(dst + 3, g.w);
        }
      }
    } else {
      const float ax = g.x * scale, ay = g.y * scale, az = g.z * scale;
      if (gx) {
        atomicAdd(gx + (long)k * 3 + 0, ax);
        atomicAdd(gx + (long)k * 3 + 1, ay);
        atomicAdd(gx + (long)k * 3 + 2, az);
      }
      sx += ax;
      sy += ay;
      sz += az;
    }
  }
  if (grad_centre) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sx += __shfl_xor_sync(0xffffffffu, sx, o);
      sy += __shfl_xor_sync(0xffffffffu, sy, o);
      sz += __shfl_xor_sync(0xffffffffu, sz, o);
    }
    if (lane == 0) {
      float* gc = grad_centre + ((long)b * M + m) * 3;
      gc[0] = -sx;
      gc[1] = -sy;
      gc[2] = -sz;
    }
  }
}

// out[b,i,:] = w0*f[b,i0,:] + w1*f[b,i1,:] + w2*f[b,i2,:]   (C % 4 == 0), upstream fma order
__global__ void __launch_bounds__(256) interp_rows_fwd_kernel(
    const float* __restrict__ feat, const int32_t* __restrict__ idx, const float* __restrict__ weight,
    int C4, int m, int n, long total, float* __restrict__ out) {
  for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total;
       e += (long)gridDim.x * blockDim.x) {
    const int j = (int)(e % C4);
    const long bi = e / C4;
    const long b = bi / n;
    const float4* f = reinterpret_cast<const float4*>(feat + b * (long)m * C4 * 4);
    const int i0 = __ldg(idx + bi * 3), i1 = __ldg(idx + bi * 3 + 1), i2 = __ldg(idx + bi * 3 + 2);
    const float w0 = __ldg(weight + bi * 3), w1 = __ldg(weight + bi * 3 + 1),
                w2 = __ldg(weight + bi * 3 + 2);
    const float4 p0 = __ldg(f + (long)i0 * C4 + j), p1 = __ldg(f + (long)i1 * C4 + j),
                 p2 = __ldg(f + (long)i2 * C4 + j);
    float4 o;
    o.x = __fmaf_rn(w2, p2.x, __fmaf_rn(w0, p0.x, __fmul_rn(w1, p1.x)));
    o.y = __fmaf_rn(w2, p2.y, __fmaf_rn(w0, p0.y, __fmul_rn(w1, p1.y)));
    o.z = __fmaf_rn(w2, p2.z, __fmaf_rn(w0, p0.z, __fmul_rn(w1, p1.z)));
    o.w = __fmaf_rn(w2, p2.w, __fmaf_rn(w0, p0.w, __fmul_rn(w1, p1.w)));
    reinterpret_cast<float4*>(out)[e] = o;
  }
}

__global__ void __launch_bounds__(256) interp_rows_bwd_kernel(
    const float* __restrict__ grad_out, const int32_t* __restrict__ idx,
    const float* __restrict__ weight, int C4, int m, int n, long total,
    float* __restrict__ grad_feat) {
  for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total;
       e += (long)gridDim.x * blockDim.x) {
    const int j = (int)(e % C4);
    const long bi = e / C4;
    const long b = bi / n;
    float* g = grad_feat + b * (long)m * C4 * 4 + j * 4;
    const float4 go = __ldg(reinterpret_cast<const float4*>(grad_out) + e);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int i = __ldg(idx + bi * 3 + k);
      const float w = __ldg(weight + bi * 3 + k);
      red_add_v4(g + (long)i * C4 * 4, make_float4(go.x * w, go.y * w, go.z * w, go.w * w));
    }
  }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace
}  // namespace demf

using namespace demf;

extern "C" {

int demf_group_rows_width(int C) { return ((C + 3) / 4) * 4 + 4; }

int demf_query_and_group_rows_fwd(const float* xyz, const float* feat_rows, const float* new_xyz,
                                  int B, int N, int M, int C, float min_radius, float max_radius,
                                  int ns, int normalize_xyz, int query, const void* grid, int32_t* idx,
                                  float* out, void* stream) {
  DEMF_REQUIRE_PTR(xyz);
  DEMF_REQUIRE_PTR(new_xyz);
  DEMF_REQUIRE_PTR(idx);
  DEMF_REQUIRE_PTR(out);
  if (C > 0) DEMF_REQUIRE_PTR(feat_rows);
  DEMF_REQUIRE(B >= 0 && N > 0 && M >= 0 && ns > 0 && C >= 0, DEMF_E_SIZE);
  DEMF_REQUIRE(B <= 65535, DEMF_E_SIZE);
  DEMF_REQUIRE(aligned16(out) && (C % 4 != 0 || aligned16(feat_rows)), DEMF_E_UNSUPPORTED);
  if (B == 0 || M == 0) return 0;
  const size_t smem = (query ? (size_t)kTile * 3 * 4 : 0) + (size_t)kWarps * ns * 4;
  DEMF_REQUIRE(smem <= 200 * 1024, DEMF_E_UNSUPPORTED);
  const float min_r2 = min_radius * min_radius, max_r2 = max_radius * max_radius;
  const float inv_radius = 1.0f / max_radius;
  dim3 grid_dim((M + kWarps - 1) / kWarps, B);
  cudaStream_t st = as_stream(stream);
  if (query) {
    auto k = group_rows_fwd_kernel<true>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    static_assert(kWarps * (kGridCap + kGridHist) * 4 <= kTile * 3 * 4, "grid scratch fits the cloud tile");
    k<<<grid_dim, kThreads, smem, st>>>(xyz, feat_rows, new_xyz, N, M, C, min_r2, max_r2, inv_radius, ns,
                                        normalize_xyz, ns <= kGridCap ? grid : nullptr, idx, out);
    return after_launch("query_group_rows_fwd_kernel");
  }
  auto k = group_rows_fwd_kernel<false>;
  if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k<<<grid_dim, kThreads, smem, st>>>(xyz, feat_rows, new_xyz, N, M, C, min_r2, max_r2, inv_radius, ns,
                                      normalize_xyz, nullptr, idx, out);
  return after_launch("group_rows_fwd_kernel");
}

int demf_group_rows_bwd(const float* grad_out, const int32_t* idx, int B, int N, int M, int C, int ns,
                        float xyz_scale, float* grad_feat_rows, float* grad_xyz, float* grad_centre,
                        void* stream) {
  DEMF_REQUIRE_PTR(grad_out);
  DEMF_REQUIRE_PTR(idx);
  DEMF_REQUIRE(B >= 0 && N > 0 && M >= 0 && ns > 0 && C >= 0, DEMF_E_SIZE);
  DEMF_REQUIRE(B <= 65535, DEMF_E_SIZE);
  DEMF_REQUIRE(aligned16(grad_out) && (C % 4 != 0 || aligned16(grad_feat_rows)), DEMF_E_UNSUPPORTED);
  if (B == 0 || M == 0) return 0;
  dim3 grid((M + kWarps - 1) / kWarps, B);
  group_rows_bwd_kernel<<<grid, kThreads, 0, as_stream(stream)>>>(
      grad_out, idx, N, M, C, ns, xyz_scale, C > 0 ? grad_feat_rows : nullptr, grad_xyz, grad_centre);
  return after_launch("group_rows_bwd_kernel");
}

int demf_three_interpolate_rows_fwd(const float* feat_rows, const int32_t* idx, const float* weight,
                                    int B, int C, int m, int n, float* out, void* stream) {
  DEMF_REQUIRE_PTR(feat_rows);
  DEMF_REQUIRE_PTR(idx);
  DEMF_REQUIRE_PTR(weight);
  DEMF_REQUIRE_PTR(out);
  DEMF_REQUIRE(B >= 0 && C > 0 && m > 0 && n >= 0, DEMF_E_SIZE);
  DEMF_REQUIRE(C % 4 == 0 && aligned16(feat_rows) && aligned16(out), DEMF_E_UNSUPPORTED);
  const long total = (long)B * n * (C / 4);
  if (total == 0) return 0;
  long blocks = (total + 255) / 256;
  if (blocks > (long)kNumSMs * 16) blocks = (long)kNumSMs * 16;
  interp_rows_fwd_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(feat_rows, idx, weight,
                                                                         C / 4, m, n, total, out);
  return after_launch("interp_rows_fwd_kernel");
}

int demf_three_interpolate_rows_bwd(const float* grad_out, const int32_t* idx, const float* weight,
                                    int B, int C, int n, int m, float* grad_feat_rows, void* stream) {
  DEMF_REQUIRE_PTR(grad_out);
  DEMF_REQUIRE_PTR(idx);
  DEMF_REQUIRE_PTR(weight);
  DEMF_REQUIRE_PTR(grad_feat_rows);
  DEMF_REQUIRE(B >= 0 && C > 0 && m > 0 && n >= 0, DEMF_E_SIZE);
  DEMF_REQUIRE(C % 4 == 0 && aligned16(grad_feat_rows) && aligned16(grad_out), DEMF_E_UNSUPPORTED);
  const long total = (long)B * n * (C / 4);
  if (total == 0) return 0;
  long blocks = (total + 255) / 256;
  if (blocks > (long)kNumSMs * 16) blocks = (long)kNumSMs * 16;
  interp_rows_bwd_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(
      grad_out, idx, weight, C / 4, m, n, total, grad_feat_rows);
  return after_launch("interp_rows_bwd_kernel");
}

}  // extern "C"
