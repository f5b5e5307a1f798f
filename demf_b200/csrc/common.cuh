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

// ---- CTA-level tracing (development): when a trace buffer is set (demf_trace_set), thread 0 of
// every CTA of the instrumented kernels appends {kernel id, block, SM id, start ns, end ns}.
struct TraceRec {
  int kernel, block, smid, pad;
  unsigned long long t0, t1;
};
struct TraceCtl {
  TraceRec* recs;
  unsigned* count;
  unsigned cap;
};
#ifdef __CUDACC__
// one control block per translation unit (no relocatable device code in this library)
static __device__ TraceCtl g_trace = {nullptr, nullptr, 0};

__device__ __forceinline__ unsigned long long trace_begin() {
  unsigned long long t = 0;
  if (g_trace.recs != nullptr && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void trace_end(int kernel, unsigned long long t0) {
  if (g_trace.recs != nullptr && threadIdx.x == 0) {
    unsigned long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    const unsigned i = atomicAdd(g_trace.count, 1u);
    if (i < g_trace.cap) {
      TraceRec r;
      r.kernel = kernel;
      r.block = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
      r.smid = (int)smid;
      r.pad = 0;
      r.t0 = t0;
      r.t1 = t1;
      g_trace.recs[i] = r;
    }
  }
}
// each instrumented .cu defines  void trace_set_<tu>(const TraceCtl&)  with this macro
#define DEMF_DEFINE_TRACE_SETTER(name) \
  void name(const TraceCtl& c) { cudaMemcpyToSymbol(g_trace, &c, sizeof(TraceCtl)); }
#endif
void trace_set_fps(const TraceCtl& c);
void trace_set_sa(const TraceCtl& c);
void trace_set_grid(const TraceCtl& c);
void trace_set_msda(const TraceCtl& c);

}  // namespace demf
