// Library-wide plumbing: thread-local error string, launch counter, device attributes.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "../../include/vlmclip.h"
#include "common.cuh"

namespace vlmclip {

namespace {
thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};
}  // namespace

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int report_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return 0;
  set_last_error("%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
  return static_cast<int>(e);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int sm_count() {
  static int n = []() {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) return 148;
    return v;
  }();
  return n;
}

}  // namespace vlmclip

extern "C" int vlmclip_abi_version(void) { return VLMCLIP_ABI_VERSION; }
extern "C" const char* vlmclip_last_error(void) { return vlmclip::g_err; }
extern "C" int64_t vlmclip_launch_count(void) { return vlmclip::g_launches.load(std::memory_order_relaxed); }
