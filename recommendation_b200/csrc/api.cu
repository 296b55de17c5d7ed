// Library-level entry points: version string, thread-local error message, device query.
#include "common.cuh"
#include <cstdarg>
#include <cstdio>

namespace gcf {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  // per-device cache; cudaDeviceGetAttribute is cheap but not free on the launch path
  static thread_local int cached_dev = -1;
  static thread_local int cached_sms = 148;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) cached_sms = n;
    cached_dev = dev;
  }
  return cached_sms;
}

}  // namespace gcf

extern "C" const char* gcf_version(void) { return "gcf-b200 0.1.0 (sm_100a)"; }
extern "C" const char* gcf_last_error(void) { return gcf::g_err; }

