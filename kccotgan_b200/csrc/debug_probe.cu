// Development probes (not part of the product path): raw TMA streaming throughput for the box
// shapes the cost kernels use.  Exposed as kccot_debug_* so that scripts/ can time them on a B200.
#include "cost.cuh"
#include "tc_common.cuh"

namespace kccot {
namespace {
constexpr int kMaxStages = 12;
struct PBars { uint64_t full[kMaxStages], empty[kMaxStages]; };

// every CTA streams its contiguous share of k-blocks ([rows x 32] fp32 boxes) through a ring; a
// consumer warp only releases the slots.  prefetch_dist > 0: L2 prefetch of a [rows x 256] box that far ahead.
__global__ void __launch_bounds__(64, 1)
tma_stream_kernel(const __grid_constant__ CUtensorMap tm, const __grid_constant__ CUtensorMap tmp, int rows,
                  long long K, int nstages, int kbps, int prefetch_dist, float* sink) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  PBars& bars = *reinterpret_cast<PBars*>(base);
  uint8_t* tiles = base + 1024;
  const int warp = threadIdx.x >> 5;
  const int nkb = (int)((K + 31) / 32);
  const int kb0 = blockIdx.x * kbps, kb1 = min(nkb, kb0 + kbps);
  if (threadIdx.x == 0) {
    for (int s = 0; s < nstages; ++s) { tc::mbar_init(&bars.full[s], 1); tc::mbar_init(&bars.empty[s], 1); }
    tc::fence_barrier_init();
  }
  __syncthreads();
  float acc = 0.f;
  if (warp == 0) {
    if (tc::elect_one()) {
      int stage = 0, phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        if (prefetch_dist > 0 && ((kb - kb0) % 8) == 0 && kb + prefetch_dist < kb1) {
          asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global [%0, {%1, %2, %3}];" ::"l"(
                           reinterpret_cast<uint64_t>(&tmp)),
                       "r"((kb + prefetch_dist) * 32), "r"(0), "r"(0)
                       : "memory");
        }
        tc::mbar_wait(&bars.empty[stage], phase ^ 1);
        tc::mbar_arrive_expect_tx(&bars.full[stage], (uint32_t)rows * 128);
        tc::tma_load_3d(&tm, &bars.full[stage], tiles + (size_t)stage * rows * 128, kb * 32, 0, 0);
        if (++stage == nstages) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    int stage = 0, phase = 0;
    for (int kb = kb0; kb < kb1; ++kb) {
      tc::mbar_wait(&bars.full[stage], phase);
      acc += *reinterpret_cast<const float*>(tiles + (size_t)stage * rows * 128 + (threadIdx.x & 31) * 4);
      __syncwarp();
      if ((threadIdx.x & 31) == 0) tc::mbar_arrive(&bars.empty[stage]);
      if (++stage == nstages) { stage = 0; phase ^= 1; }
    }
  }
  if (acc == 123.456f) *sink = acc;
}
}  // namespace
}  // namespace kccot

using namespace kccot;
extern "C" int kccot_debug_tma_stream(const float* x, int rows, long long K, int nstages, int prefetch_dist,
                                      float* sink, void* stream) {
  KCCOT_CHECK_ARG(x && rows >= 8 && rows <= 256 && rows % 8 == 0 && K % 4 == 0 && nstages >= 1 && nstages <= kMaxStages,
                  "bad probe arguments");
  CUtensorMap tm, tmp;
  if (int rc = encode_tmap_3d(&tm, x, (uint64_t)K, (uint64_t)rows, 1, (uint64_t)K * 4, (uint64_t)K * 4 * rows, 32,
                              (uint32_t)rows))
    return rc;
  if (int rc = encode_tmap_3d(&tmp, x, (uint64_t)K, (uint64_t)rows, 1, (uint64_t)K * 4, (uint64_t)K * 4 * rows, 256,
                              (uint32_t)rows, false, true))
    return rc;
  const int nkb = (int)((K + 31) / 32);
  const int sms = num_sms();
  const int kbps = (nkb + sms - 1) / sms;
  const size_t smem = 2048 + (size_t)nstages * rows * 128;
  KCCOT_CUDA(cudaFuncSetAttribute(tma_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  tma_stream_kernel<<<(nkb + kbps - 1) / kbps, 64, smem, (cudaStream_t)stream>>>(tm, tmp, rows, K, nstages, kbps,
                                                                                prefetch_dist, sink);
  KCCOT_LAUNCH_CHECK();
  return KCCOT_OK;
}
