// Shared helpers for libkccot (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/kccot.h"

namespace kccot {

void set_error(const char* fmt, ...);
extern std::atomic<unsigned long long> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

#define KCCOT_CHECK_ARG(cond, ...)            \
  do {                                        \
    if (!(cond)) {                            \
      ::kccot::set_error(__VA_ARGS__);        \
      return KCCOT_EINVAL;                    \
    }                                         \
  } while (0)

#define KCCOT_CUDA(expr)                                                                  \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      ::kccot::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                         __LINE__);                                                       \
      return KCCOT_ECUDA;                                                                 \
    }                                                                                     \
  } while (0)

#define KCCOT_LAUNCH_CHECK()                                                             \
  do {                                                                                   \
    ::kccot::count_launch();                                                             \
    cudaError_t _e = cudaGetLastError();                                                 \
    if (_e != cudaSuccess) {                                                             \
      ::kccot::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e),      \
                         __FILE__, __LINE__);                                            \
      return KCCOT_ECUDA;                                                                \
    }                                                                                    \
  } while (0)

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
int num_sms();

struct SideLane {
  cudaStream_t stream;
  cudaEvent_t fork, join;
};
SideLane* side_lane();   // per host thread and device; nullptr if the stream / events cannot be created

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_log2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

}  // namespace kccot
