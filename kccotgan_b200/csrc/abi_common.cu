// Error state, version, device check, launch counter.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace kccot {
static thread_local char g_err[512] = "";
std::atomic<unsigned long long> g_launches{0};
thread_local CtxHint t_ctx = {0, 0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int num_sms() {
  // per device (a process may drive several GPUs); a benign race between host threads costs a redundant query
  static int cached[kMaxDevices] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) cached[dev] = n;
    else return 148;
  }
  return cached[dev];
}

// A second stream (plus fork / join events) per host thread and device, for the one place where two
// independent kernels of a call can overlap (martingale adjoint next to the gradient GEMM).  Works under
// stream capture too: waiting on the captured fork event pulls the side stream into the capture, the
// join event brings it back before the call returns.
SideLane* side_lane() {
  constexpr int kMaxDev = 64;
  static thread_local SideLane lanes[kMaxDev];
  static thread_local bool ready[kMaxDev] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDev) return nullptr;
  if (!ready[dev]) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    (void)cs;
    SideLane& l = lanes[dev];
    if (cudaStreamCreateWithFlags(&l.stream, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&l.fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&l.join, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    ready[dev] = true;
  }
  return &lanes[dev];
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
