// Uniform-grid binning of a point cloud for the exact grid ball query (ball_grid.cuh), and the
// stand-alone ball query through it (same (B,M,ns) rows as demf_ball_query).
#include "ball_grid.cuh"

namespace demf {
namespace {

constexpr int kBuildThreads = 1024;

__device__ __forceinline__ float block_reduce(float v, bool want_min, float* scratch) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float t = __shfl_xor_sync(0xffffffffu, v, o);
    v = want_min ? fminf(v, t) : fmaxf(v, t);
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  v = scratch[threadIdx.x & 31];  // kBuildThreads / 32 == 32 partials
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float t = __shfl_xor_sync(0xffffffffu, v, o);
    v = want_min ? fminf(v, t) : fmaxf(v, t);
  }
  return v;
}

// One CTA per scene: bounding box -> cell edge -> counting sort of the points by cell.
__global__ void __launch_bounds__(kBuildThreads) ball_grid_build_kernel(const float* __restrict__ xyz, int N,
                                                                        float radius, unsigned char* workspace) {
  extern __shared__ int cnt[];  // kGridCells counters, then cursors
  const unsigned long long trace_t0 = trace_begin();
  __shared__ float scratch[32];
  __shared__ int warp_tot[32];
  const int b = blockIdx.x, tid = threadIdx.x;
  const float* cloud = xyz + (long)b * N * 3;
  unsigned char* base = workspace + (size_t)b * ball_grid_scene_bytes(N);
  BallGridHeader* hdr = reinterpret_cast<BallGridHeader*>(base);
  int* cell_start = reinterpret_cast<int*>(base + sizeof(BallGridHeader));
  float4* sorted = reinterpret_cast<float4*>(base + sizeof(BallGridHeader) + (size_t)(kGridCells + 4) * 4);

  float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
  for (int i = tid; i < N; i += kBuildThreads) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float v = __ldg(cloud + i * 3L + a);
      lo[a] = fminf(lo[a], v);
      hi[a] = fmaxf(hi[a], v);
    }
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    lo[a] = block_reduce(lo[a], true, scratch);
    hi[a] = block_reduce(hi[a], false, scratch);
  }
  // cell edge: a little more than the radius (so that |p - c| < r keeps p within +-1 cell of c even
  // after rounding), and large enough for the box to fit the fixed grid
  float h = radius * 1.001f + 1e-6f;
  h = fmaxf(h, (hi[0] - lo[0]) * (1.002f / kGridX));
  h = fmaxf(h, (hi[1] - lo[1]) * (1.002f / kGridY));
  h = fmaxf(h, (hi[2] - lo[2]) * (1.002f / kGridZ));
  const float inv_h = 1.0f / h;
  const int nx = min(kGridX, max(ball_grid_coord(hi[0], lo[0], inv_h), 0) + 1);
  const int ny = min(kGridY, max(ball_grid_coord(hi[1], lo[1], inv_h), 0) + 1);
  const int nz = min(kGridZ, max(ball_grid_coord(hi[2], lo[2], inv_h), 0) + 1);
  if (tid == 0) {
    hdr->x0 = lo[0];
    hdr->y0 = lo[1];
    hdr->z0 = lo[2];
    hdr->inv_h = inv_h;
    hdr->nx = nx;
    hdr->ny = ny;
    hdr->nz = nz;
    hdr->radius = radius;
  }
  for (int c = tid; c < kGridCells; c += kBuildThreads) cnt[c] = 0;
  __syncthreads();
  auto cell_of = [&](float x, float y, float z) {
    const int ix = min(max(ball_grid_coord(x, lo[0], inv_h), 0), nx - 1);
    const int iy = min(max(ball_grid_coord(y, lo[1], inv_h), 0), ny - 1);
    const int iz = min(max(ball_grid_coord(z, lo[2], inv_h), 0), nz - 1);
    return (ix * kGridY + iy) * kGridZ + iz;
  };
  for (int i = tid; i < N; i += kBuildThreads)
    atomicAdd(&cnt[cell_of(__ldg(cloud + i * 3L), __ldg(cloud + i * 3L + 1), __ldg(cloud + i * 3L + 2))], 1);
  __syncthreads();
  // exclusive scan: kGridCells / kBuildThreads = 16 consecutive cells per thread
  constexpr int per = kGridCells / kBuildThreads;
  int local[per], sum = 0;
#pragma unroll
  for (int k = 0; k < per; ++k) {
    local[k] = cnt[tid * per + k];
    sum += local[k];
  }
  int incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if ((tid & 31) >= o) incl += t;
  }
  if ((tid & 31) == 31) warp_tot[tid >> 5] = incl;
  __syncthreads();
  if (tid < 32) {
    int w = warp_tot[tid];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, w, o);
      if (tid >= o) w += t;
    }
    warp_tot[tid] = w;  // inclusive over warps
  }
  __syncthreads();
  int run = incl - sum + ((tid >> 5) ? warp_tot[(tid >> 5) - 1] : 0);
#pragma unroll
  for (int k = 0; k < per; ++k) {
    cnt[tid * per + k] = run;
    cell_start[tid * per + k] = run;
    run += local[k];
  }
  if (tid == kBuildThreads - 1) cell_start[kGridCells] = run;  // == N
  __syncthreads();
  for (int i = tid; i < N; i += kBuildThreads) {
    const float x = __ldg(cloud + i * 3L), y = __ldg(cloud + i * 3L + 1), z = __ldg(cloud + i * 3L + 2);
    const int pos = atomicAdd(&cnt[cell_of(x, y, z)], 1);
    sorted[pos] = make_float4(x, y, z, __int_as_float(i));
  }
  trace_end(4, trace_t0);
}

constexpr int kQueryWarps = 8;

// kBitmap: selection through a per-warp bitmap over the index range (ball_grid_query_bitmap_warp) when N / 8 bytes
// per warp fit; otherwise the hit-buffer form.
template <bool kBitmap>
__global__ void __launch_bounds__(kQueryWarps * 32) ball_query_grid_kernel(
    const float* __restrict__ xyz, const float* __restrict__ new_xyz, const void* __restrict__ grid, int N, int M,
    float min_r2, float max_r2, int ns, int32_t* __restrict__ idx) {
  extern __shared__ __align__(16) int smem_i[];
  const unsigned long long trace_t0 = trace_begin();
  const unsigned lane = lane_id();
  const int warp = threadIdx.x >> 5;
  const int b = blockIdx.y;
  const int m = blockIdx.x * kQueryWarps + warp;
  if (m >= M) return;
  const int scratch = kBitmap ? ball_bitmap_words(N) : kGridCap + kGridHist;
  int* buf = smem_i + warp * scratch;
  int32_t* row = smem_i + kQueryWarps * scratch + warp * ns;
  const float* c = new_xyz + ((long)b * M + m) * 3;
  const BallGridView g = ball_grid_view(grid, b, N);
  const float cx = __ldg(c), cy = __ldg(c + 1), cz = __ldg(c + 2);
  if (kBitmap && max_r2 <= g.h.radius * g.h.radius)
    ball_grid_query_bitmap_warp(g, N, cx, cy, cz, min_r2, max_r2, ns, row, reinterpret_cast<unsigned*>(buf), lane);
  else if (kBitmap)
    ball_scan_warp(xyz + (long)b * N * 3, N, cx, cy, cz, min_r2, max_r2, ns, row, lane);
  else
    ball_query_warp(g, xyz + (long)b * N * 3, N, cx, cy, cz, min_r2, max_r2, ns, row, buf, buf + kGridCap, lane);
  int32_t* out = idx + ((long)b * M + m) * ns;
  for (int l = lane; l < ns; l += 32) out[l] = row[l];
  trace_end(3, trace_t0);
}

}  // namespace

int launch_ball_query_grid(const float* xyz, const float* new_xyz, const void* grid, int B, int N, int M,
                           float min_radius, float max_radius, int ns, int32_t* idx, cudaStream_t stream) {
  // the bitmap form wants N/8 bytes per warp: up to 8 KB (N = 65 536) keeps 3 CTAs of 8 warps on an SM
  const bool bitmap = ball_bitmap_words(N) * 4 <= 8 * 1024;
  const size_t smem = (size_t)kQueryWarps * ((bitmap ? ball_bitmap_words(N) : kGridCap + kGridHist) + ns) * 4;
  auto kernel = bitmap ? ball_query_grid_kernel<true> : ball_query_grid_kernel<false>;
  static size_t configured[2] = {48 * 1024, 48 * 1024};
  if (smem > configured[bitmap]) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    configured[bitmap] = smem;
  }
  dim3 grid_dim((M + kQueryWarps - 1) / kQueryWarps, B);
  kernel<<<grid_dim, kQueryWarps * 32, smem, stream>>>(xyz, new_xyz, grid, N, M, min_radius * min_radius,
                                                      max_radius * max_radius, ns, idx);
  return after_launch("ball_query_grid_kernel");
}

DEMF_DEFINE_TRACE_SETTER(trace_set_grid)
}  // namespace demf

using namespace demf;

extern "C" {

size_t demf_ball_grid_workspace_bytes(int B, int N) {
  if (B <= 0 || N <= 0) return 0;
  return (size_t)B * ball_grid_scene_bytes(N);
}

int demf_ball_grid_build(const float* xyz, int B, int N, float radius, void* workspace, void* stream) {
  DEMF_REQUIRE_PTR(xyz);
  DEMF_REQUIRE_PTR(workspace);
  DEMF_REQUIRE(B >= 0 && N > 0 && radius > 0.f, DEMF_E_SIZE);
  DEMF_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15u) == 0, DEMF_E_UNSUPPORTED);
  if (B == 0) return 0;
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(ball_grid_build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kGridCells * 4);
    configured = true;
  }
  ball_grid_build_kernel<<<B, kBuildThreads, kGridCells * 4, as_stream(stream)>>>(
      xyz, N, radius, static_cast<unsigned char*>(workspace));
  return after_launch("ball_grid_build_kernel");
}

int demf_ball_query_grid(const float* xyz, const float* new_xyz, const void* grid, int B, int N, int M,
                         float min_radius, float max_radius, int ns, int32_t* idx, void* stream) {
  DEMF_REQUIRE_PTR(xyz);
  DEMF_REQUIRE_PTR(new_xyz);
  DEMF_REQUIRE_PTR(grid);
  DEMF_REQUIRE_PTR(idx);
  DEMF_REQUIRE(B >= 0 && N > 0 && M >= 0 && ns > 0 && ns <= kGridCap, DEMF_E_SIZE);
  DEMF_REQUIRE(B <= 65535, DEMF_E_SIZE);
  if (B == 0 || M == 0) return 0;
  return launch_ball_query_grid(xyz, new_xyz, grid, B, N, M, min_radius, max_radius, ns, idx, as_stream(stream));
}

}  // extern "C"
