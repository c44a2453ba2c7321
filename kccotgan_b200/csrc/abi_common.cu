// Error state, version, device check, launch counter.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace kccot {
static thread_local char g_err[512] = "";
std::atomic<unsigned long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int num_sms() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      return 148;
  }
  return cached;
}
}  // namespace kccot

extern "C" {
int kccot_version(void) { return KCCOT_VERSION; }
const char* kccot_last_error(void) { return kccot::g_err; }
unsigned long long kccot_launch_count(void) { return kccot::g_launches.load(); }
int kccot_device_check(void) {
  int dev = 0;
  cudaDeviceProp p;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&p, dev) != cudaSuccess) {
    kccot::set_error("no CUDA device available (libkccot has no CPU fallback)");
    return KCCOT_ECUDA;
  }
  if (p.major != 10) {
    kccot::set_error("device %s is sm_%d%d; libkccot is built for sm_100a (B200) only", p.name, p.major,
                     p.minor);
    return KCCOT_EUNSUPPORTED;
  }
  return KCCOT_OK;
}
}
