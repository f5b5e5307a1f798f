// Shared host/device helpers of libdemf_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/demf_b200.h"

namespace demf {

// Thread-local message behind demf_last_error_string().
void set_error(const char* fmt, ...);
// cudaGetLastError() after a launch: returns 0 or the cudaError_t, records the message and bumps
// the process-wide launch counter (demf_launch_count()).
int after_launch(const char* kernel_name);

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

#define DEMF_REQUIRE_PTR(p)                                   \
  do {                                                        \
    if ((p) == nullptr) {                                     \
      ::demf::set_error("%s: argument '%s' is NULL", __func__, #p); \
      return DEMF_E_NULL;                                     \
    }                                                         \
  } while (0)

#define DEMF_REQUIRE(cond, code)                                           \
  do {                                                                     \
    if (!(cond)) {                                                         \
      ::demf::set_error("%s: requirement '%s' not met", __func__, #cond); \
      return (code);                                                       \
    }                                                                      \
  } while (0)

#ifdef __CUDACC__
// Squared distance in the rounding order nvcc emits for the upstream source expression
// (a-x)*(a-x)+(b-y)*(b-y)+(c-z)*(c-z):  FMUL(dy,dy) -> FFMA(dx,dx,.) -> FFMA(dz,dz,.)
// Written with intrinsics so that no compiler version can contract it differently; the CPU
// oracle uses the same fmaf chain (oracle/demf_oracle.c: sqdist).
__device__ __forceinline__ float sqdist(float ax, float ay, float az, float bx, float by, float bz) {
  const float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
  return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }
#endif

constexpr int kNumSMs = 148;  // B200

}  // namespace demf
