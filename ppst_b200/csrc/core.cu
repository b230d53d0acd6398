// Error plumbing, version, launch counter.
#include <stdarg.h>

#include <atomic>

#include "common.cuh"

namespace ppst {
static thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_check(cudaError_t e, const char* what) {
  if (e == cudaSuccess) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return 0;
  }
  set_error("%s: %s", what, cudaGetErrorString(e));
  return PPST_ERR_CUDA;
}
}  // namespace ppst

extern "C" const char* ppst_last_error(void) { return ppst::g_err; }
extern "C" int ppst_version(void) { return 100; }
extern "C" long long ppst_launch_count(void) { return ppst::g_launches.load(); }
