// Library-wide C ABI pieces: version, error string, launch counter.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace demf {

static thread_local char g_error[512] = "no error";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

int after_launch(const char* kernel_name) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: launch failed: %s (%s)", kernel_name, cudaGetErrorName(e), cudaGetErrorString(e));
    return static_cast<int>(e);
  }
  return 0;
}

}  // namespace demf

extern "C" {

int demf_version(void) { return DEMF_B200_VERSION; }

const char* demf_last_error_string(void) { return demf::g_error; }

uint64_t demf_launch_count(void) { return demf::g_launches.load(std::memory_order_relaxed); }

/* development: CTA-level trace of the instrumented kernels into a caller-owned device buffer of
 * `capacity` 32-byte records {int kernel, block, smid, pad; u64 start_ns, end_ns} plus a device
 * counter; NULL switches tracing off. Kernel ids: 1 fps, 2 sa_fused, 3 ball_query_grid,
 * 4 ball_grid_build, 5 msda_fwd. */
int demf_trace_set(void* records, unsigned* counter, unsigned capacity) {
  demf::TraceCtl c{static_cast<demf::TraceRec*>(records), counter, records ? capacity : 0u};
  demf::trace_set_fps(c);
  demf::trace_set_sa(c);
  demf::trace_set_grid(c);
  demf::trace_set_msda(c);
  return static_cast<int>(cudaGetLastError());
}

}  // extern "C"
