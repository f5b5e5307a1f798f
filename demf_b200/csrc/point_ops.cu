// Grouping, gather, three_nn and three_interpolate for sm_100a.
//
// These are the small HBM-bound index ops of the PointNet++ SA/FP modules
// (mmdet3d group_points / gather_points / interpolate extensions). One thread per
// output element with the contiguous output axis on threadIdx.x (coalesced stores,
// index rows re-read from L1/L2); backward passes are vector-free float atomics
// (RED.ADD.F32), exactly the accumulation upstream performs.
#include "common.cuh"

namespace demf {
namespace {

constexpr int kThreads = 256;

inline int grid_for(long total, int threads = kThreads) {
  long g = (total + threads - 1) / threads;
  const long cap = static_cast<long>(kNumSMs) * 16;  // grid-stride beyond 16 CTAs per SM
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

// out[b,c,j] = features[b,c,idx[b,j]]   with j over M*ns (grouping) or M (gather)
__global__ void __launch_bounds__(kThreads) gather_rows_fwd_kernel(
    const float* __restrict__ features, const int32_t* __restrict__ idx, int C, int N, long J,
    long total, float* __restrict__ out) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total;
       i += (long)gridDim.x * blockDim.x) {
    const long j = i % J;
    const long bc = i / J;
    const long b = bc / C;
    const int k = __ldg(idx + b * J + j);
    out[i] = __ldg(features + bc * N + k);
  }
}

__global__ void __launch_bounds__(kThreads) gather_rows_bwd_kernel(
    const float* __restrict__ grad_out, const int32_t* __restrict__ idx, int C, int N, long J,
    long total, float* __restrict__ grad_features) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total;
       i += (long)gridDim.x * blockDim.x) {
    const long j = i % J;
    const long bc = i / J;
    const long b = bc / C;
    const int k = __ldg(idx + b * J + j);
    atomicAdd(grad_features + bc * N + k, __ldg(grad_out + i));
  }
}

// ---------------------------------------------------------------------- three_nn --
// One warp per target point. Each lane scans the sources lane, lane+32, ... in index order
// with the upstream strict-'<' cascade (so its private list is sorted by (d2, index)); the
// warp then extracts the three smallest (d2, index) keys with two REDUX.MIN per round. A key
// comparison by (d2 bits, index) reproduces "first in index order wins ties".
__global__ void __launch_bounds__(kThreads) three_nn_kernel(
    const float* __restrict__ unknown, const float* __restrict__ known, int n, int m, long total,
    float* __restrict__ dist2, int32_t* __restrict__ idx) {
  const unsigned lane = lane_id();
  const long warp = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;
  if (warp >= total) return;  // whole warp leaves together
  const long b = warp / n;
  const float ux = __ldg(unknown + warp * 3 + 0);
  const float uy = __ldg(unknown + warp * 3 + 1);
  const float uz = __ldg(unknown + warp * 3 + 2);
  const float* kn = known + b * (long)m * 3;

  // best1..3 start at (float)1e40 = +inf upstream; 'd < best' therefore admits every finite d.
  const float kInf = __int_as_float(0x7f800000);
  float d1 = kInf, d2 = kInf, d3 = kInf;
  int i1 = 0, i2 = 0, i3 = 0;
  for (int k = lane; k < m; k += 32) {
    const float d = sqdist(ux, uy, uz, __ldg(kn + k * 3 + 0), __ldg(kn + k * 3 + 1),
                           __ldg(kn + k * 3 + 2));
    if (d < d1) {
      d3 = d2; i3 = i2;
      d2 = d1; i2 = i1;
      d1 = d; i1 = k;
    } else if (d < d2) {
      d3 = d2; i3 = i2;
      d2 = d; i2 = k;
    } else if (d < d3) {
      d3 = d; i3 = k;
    }
  }
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    // head of this lane's list; an exhausted list offers (+inf, 0x7fffffff) and never wins
    // against a real candidate (real candidates are finite).
    const unsigned hi = __float_as_uint(d1);
    const unsigned lo = (hi == 0x7f800000u) ? 0x7fffffffu : (unsigned)i1;
    const unsigned best_hi = __reduce_min_sync(0xffffffffu, hi);
    const unsigned best_lo = __reduce_min_sync(0xffffffffu, hi == best_hi ? lo : 0xffffffffu);
    if (lane == 0) {
      dist2[warp * 3 + r] = __uint_as_float(best_hi);
      idx[warp * 3 + r] = (best_hi == 0x7f800000u) ? 0 : (int)best_lo;  // unfilled slot: idx 0
    }
    if (hi == best_hi && lo == best_lo && hi != 0x7f800000u) {  // pop the winner's list
      d1 = d2; i1 = i2;
      d2 = d3; i2 = i3;
      d3 = kInf;
    }
  }
}

// out[b,c,i] = w0*f[i0] + w1*f[i1] + w2*f[i2] in nvcc's contraction order for that expression
__global__ void __launch_bounds__(kThreads) three_interpolate_fwd_kernel(
    const float* __restrict__ features, const int32_t* __restrict__ idx,
    const float* __restrict__ weight, int C, int m, int n, long total, float* __restrict__ out) {
  for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total;
       e += (long)gridDim.x * blockDim.x) {
    const long i = e % n;
    const long bc = e / n;
    const long b = bc / C;
    const long t = (b * n + i) * 3;
    const float* f = features + bc * m;
    const float p0 = __ldg(f + __ldg(idx + t + 0));
    const float p1 = __ldg(f + __ldg(idx + t + 1));
    const float p2 = __ldg(f + __ldg(idx + t + 2));
    out[e] = __fmaf_rn(__ldg(weight + t + 2), p2,
                       __fmaf_rn(__ldg(weight + t + 0), p0, __fmul_rn(__ldg(weight + t + 1), p1)));
  }
}

__global__ void __launch_bounds__(kThreads) three_interpolate_bwd_kernel(
    const float* __restrict__ grad_out, const int32_t* __restrict__ idx,
    const float* __restrict__ weight, int C, int n, int m, long total,
    float* __restrict__ grad_features) {
  for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total;
       e += (long)gridDim.x * blockDim.x) {
    const long i = e % n;
    const long bc = e / n;
    const long b = bc / C;
    const long t = (b * n + i) * 3;
    float* g = grad_features + bc * m;
    const float go = __ldg(grad_out + e);
#pragma unroll
    for (int k = 0; k < 3; ++k) atomicAdd(g + __ldg(idx + t + k), go * __ldg(weight + t + k));
  }
}

int check_rows(const void* a, const void* b, const void* c, int B, int C, int N, long J) {
  if (!a || !b || !c) {
    set_error("group/gather: NULL pointer argument");
    return DEMF_E_NULL;
  }
  if (B < 0 || C < 0 || N <= 0 || J < 0 || (long)B * C * (N > J ? N : J) >= (1L << 40)) {
    set_error("group/gather: bad sizes B=%d C=%d N=%d J=%ld", B, C, N, J);
    return DEMF_E_SIZE;
  }
  return 0;
}

}  // namespace
}  // namespace demf

using namespace demf;

extern "C" {

int demf_group_fwd(const float* features, const int32_t* idx, int B, int C, int N, int M, int ns,
                   float* out, void* stream) {
  const long J = (long)M * ns;
  if (int rc = check_rows(features, idx, out, B, C, N, J)) return rc;
  const long total = (long)B * C * J;
  if (total == 0) return 0;
  gather_rows_fwd_kernel<<<grid_for(total), kThreads, 0, as_stream(stream)>>>(features, idx, C, N, J,
                                                                             total, out);
  return after_launch("gather_rows_fwd_kernel");
}

int demf_group_bwd(const float* grad_out, const int32_t* idx, int B, int C, int N, int M, int ns,
                   float* grad_features, void* stream) {
  const long J = (long)M * ns;
  if (int rc = check_rows(grad_out, idx, grad_features, B, C, N, J)) return rc;
  const long total = (long)B * C * J;
  if (total == 0) return 0;
  gather_rows_bwd_kernel<<<grid_for(total), kThreads, 0, as_stream(stream)>>>(
      grad_out, idx, C, N, J, total, grad_features);
  return after_launch("gather_rows_bwd_kernel");
}

int demf_gather_fwd(const float* features, const int32_t* idx, int B, int C, int N, int M,
                    float* out, void* stream) {
  return demf_group_fwd(features, idx, B, C, N, M, 1, out, stream);
}

int demf_gather_bwd(const float* grad_out, const int32_t* idx, int B, int C, int N, int M,
                    float* grad_features, void* stream) {
  return demf_group_bwd(grad_out, idx, B, C, N, M, 1, grad_features, stream);
}

int demf_three_nn(const float* unknown, const float* known, int B, int n, int m, float* dist2,
                  int32_t* idx, void* stream) {
  DEMF_REQUIRE_PTR(unknown);
  DEMF_REQUIRE_PTR(known);
  DEMF_REQUIRE_PTR(dist2);
  DEMF_REQUIRE_PTR(idx);
  DEMF_REQUIRE(B >= 0 && n >= 0 && m > 0, DEMF_E_SIZE);
  const long total = (long)B * n;  // warps
  if (total == 0) return 0;
  const long blocks = (total * 32 + kThreads - 1) / kThreads;
  DEMF_REQUIRE(blocks < (1L << 31), DEMF_E_SIZE);
  three_nn_kernel<<<(unsigned)blocks, kThreads, 0, as_stream(stream)>>>(unknown, known, n, m, total,
                                                                       dist2, idx);
  return after_launch("three_nn_kernel");
}

int demf_three_interpolate_fwd(const float* features, const int32_t* idx, const float* weight, int B,
                               int C, int m, int n, float* out, void* stream) {
  DEMF_REQUIRE_PTR(features);
  DEMF_REQUIRE_PTR(idx);
  DEMF_REQUIRE_PTR(weight);
  DEMF_REQUIRE_PTR(out);
  DEMF_REQUIRE(B >= 0 && C >= 0 && m > 0 && n >= 0, DEMF_E_SIZE);
  const long total = (long)B * C * n;
  if (total == 0) return 0;
  three_interpolate_fwd_kernel<<<grid_for(total), kThreads, 0, as_stream(stream)>>>(
      features, idx, weight, C, m, n, total, out);
  return after_launch("three_interpolate_fwd_kernel");
}

int demf_three_interpolate_bwd(const float* grad_out, const int32_t* idx, const float* weight, int B,
                               int C, int n, int m, float* grad_features, void* stream) {
  DEMF_REQUIRE_PTR(grad_out);
  DEMF_REQUIRE_PTR(idx);
  DEMF_REQUIRE_PTR(weight);
  DEMF_REQUIRE_PTR(grad_features);
  DEMF_REQUIRE(B >= 0 && C >= 0 && m > 0 && n >= 0, DEMF_E_SIZE);
  const long total = (long)B * C * n;
  if (total == 0) return 0;
  three_interpolate_bwd_kernel<<<grid_for(total), kThreads, 0, as_stream(stream)>>>(
      grad_out, idx, weight, C, n, m, total, grad_features);
  return after_launch("three_interpolate_bwd_kernel");
}

}  // extern "C"
