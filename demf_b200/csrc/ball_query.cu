// Ball query and the fused QueryAndGroup for sm_100a.
//
// Upstream (mmdet3d ball_query_kernel) runs one thread per centre, each re-reading the whole
// cloud with a 12-byte stride and a serial early exit. Here a CTA owns kWarps centres of one
// scene and streams the cloud through shared memory in coalesced tiles; one warp scans one
// centre 32 points at a time, compacting hits in index order with a ballot + popc prefix, so
// the "first nsample hits in index order, pad with the first hit" result is identical while the
// cloud is read once per CTA (not once per thread) and every load is a full 128-byte line.
//
// query_and_group_kernel continues in the same launch: the index row just produced (still in
// shared memory) drives the xyz / feature gather, the centre subtraction, the radius
// normalisation and the channel concat, writing the (B, 3+C, M, ns) tensor the SA MLP consumes.
// That replaces upstream's ball_query + 2x group_points + transpose + sub + div + cat launches
// and their intermediate tensors.
#include "common.cuh"

namespace demf {
namespace {

constexpr int kWarps = 16;                 // centres per CTA
constexpr int kThreads = kWarps * 32;
constexpr int kTile = 2048;                // points per shared-memory tile (24 KB)

template <bool kGroup>
__global__ void __launch_bounds__(kThreads) ball_query_kernel(
    const float* __restrict__ xyz, const float* __restrict__ features,
    const float* __restrict__ new_xyz, int N, int M, int C, float min_r2, float max_r2,
    float inv_radius, int ns, int use_xyz, int normalize_xyz, int32_t* __restrict__ idx,
    float* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* tile = reinterpret_cast<float*>(smem_raw);                       // kTile*3 floats
  int32_t* rows = reinterpret_cast<int32_t*>(smem_raw + kTile * 3 * 4);   // kWarps*ns ints

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
  int cnt = 0;
  int first = 0;
  bool done = !active;

  for (int base = 0; base < N; base += kTile) {
    const int npts = min(kTile, N - base);
    __syncthreads();  // previous tile fully consumed
    for (int i = threadIdx.x; i < npts * 3; i += kThreads) tile[i] = __ldg(cloud + (long)base * 3 + i);
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
    if (__syncthreads_and(done)) break;  // every centre of this CTA is full
  }
  if (!active) return;
  if (cnt > ns) cnt = ns;
  __syncwarp();
  // pad with the first hit; an empty ball leaves zeros (upstream's pre-zeroed output)
  for (int l = cnt + lane; l < ns; l += 32) row[l] = first;
  __syncwarp();
  int32_t* grow = idx + ((long)b * M + m) * ns;
  for (int l = lane; l < ns; l += 32) grow[l] = row[l];

  if (kGroup) {
    const int Cx = use_xyz ? 3 : 0;
    const long plane = (long)M * ns;
    float* obase = out + (long)b * (Cx + C) * plane + (long)m * ns;
    if (use_xyz) {
      const float cc[3] = {cx, cy, cz};
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        for (int l = lane; l < ns; l += 32) {
          float v = __fsub_rn(__ldg(cloud + (long)row[l] * 3 + a), cc[a]);
          // torch's CUDA `tensor / python_float` is tensor * (1.0f / float), not a division
          if (normalize_xyz) v = __fmul_rn(v, inv_radius);
          obase[a * plane + l] = v;
        }
      }
    }
    const float* fb = features + (long)b * C * N;
    for (int ch = 0; ch < C; ++ch) {
      const float* f = fb + (long)ch * N;
      for (int l = lane; l < ns; l += 32) obase[(Cx + ch) * plane + l] = __ldg(f + row[l]);
    }
  }
}

int launch(bool group, const float* xyz, const float* features, const float* new_xyz, int B, int N,
           int M, int C, float min_radius, float max_radius, int ns, int use_xyz, int normalize_xyz,
           int32_t* idx, float* out, void* stream) {
  if (B == 0 || M == 0) return 0;
  const size_t smem = (size_t)kTile * 3 * 4 + (size_t)kWarps * ns * 4;
  if (smem > 200 * 1024) {
    set_error("ball_query: nsample=%d needs %zu bytes of shared memory", ns, smem);
    return DEMF_E_UNSUPPORTED;
  }
  const float min_r2 = min_radius * min_radius;  // float*float as upstream
  const float max_r2 = max_radius * max_radius;
  const float inv_radius = 1.0f / max_radius;
  dim3 grid((M + kWarps - 1) / kWarps, B);
  if (group) {
    auto k = ball_query_kernel<true>;
    if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<<<grid, kThreads, smem, as_stream(stream)>>>(xyz, features, new_xyz, N, M, C, min_r2, max_r2,
                                                   inv_radius, ns, use_xyz, normalize_xyz, idx, out);
    return after_launch("query_and_group_kernel");
  }
  auto k = ball_query_kernel<false>;
  if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k<<<grid, kThreads, smem, as_stream(stream)>>>(xyz, nullptr, new_xyz, N, M, 0, min_r2, max_r2,
                                                 inv_radius, ns, 0, 0, idx, nullptr);
  return after_launch("ball_query_kernel");
}

}  // namespace
}  // namespace demf

using namespace demf;

extern "C" {

int demf_ball_query(const float* xyz, const float* new_xyz, int B, int N, int M, float min_radius,
                    float max_radius, int nsample, int32_t* idx, void* stream) {
  DEMF_REQUIRE_PTR(xyz);
  DEMF_REQUIRE_PTR(new_xyz);
  DEMF_REQUIRE_PTR(idx);
  DEMF_REQUIRE(B >= 0 && N > 0 && M >= 0 && nsample > 0, DEMF_E_SIZE);
  DEMF_REQUIRE(B <= 65535, DEMF_E_SIZE);
  return launch(false, xyz, nullptr, new_xyz, B, N, M, 0, min_radius, max_radius, nsample, 0, 0, idx,
                nullptr, stream);
}

int demf_query_and_group_fwd(const float* xyz, const float* features, const float* new_xyz, int B,
                             int N, int M, int C, float min_radius, float max_radius, int ns,
                             int use_xyz, int normalize_xyz, int32_t* idx, float* out, void* stream) {
  DEMF_REQUIRE_PTR(xyz);
  DEMF_REQUIRE_PTR(new_xyz);
  DEMF_REQUIRE_PTR(idx);
  DEMF_REQUIRE_PTR(out);
  DEMF_REQUIRE(B >= 0 && N > 0 && M >= 0 && ns > 0 && C >= 0, DEMF_E_SIZE);
  DEMF_REQUIRE(B <= 65535, DEMF_E_SIZE);
  if (C > 0) DEMF_REQUIRE_PTR(features);
  DEMF_REQUIRE(use_xyz || C > 0, DEMF_E_UNSUPPORTED);
  return launch(true, xyz, features, new_xyz, B, N, M, C, min_radius, max_radius, ns, use_xyz,
                normalize_xyz, idx, out, stream);
}

}  // extern "C"
