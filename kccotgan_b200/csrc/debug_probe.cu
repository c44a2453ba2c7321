// Development probes (not part of the product path): raw TMA streaming throughput for the box
// shapes the cost kernels use.  Exposed as kccot_debug_* so that scripts/ can time them on a B200.
#ifdef KCCOT_DEV      // compiled only into the development library (python -m kccotgan_b200.build --dev -> libkccot_dev.so)
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

// ---------------------------------------------------------------------------------------------
// Mat-vec iteration probe: the scaling-form Sinkhorn loop at B = 64 with LPR lanes per row
// (LPR = 4 is the product kernel's mapping; 2 and 1 trade shuffles and barrier width against longer
// FMA chains).  Reports clock64 cycles per iteration of CTA 0.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t probe_pin(uint32_t v) {
  uint32_t r;
  asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(v));
  return r;
}
__device__ __forceinline__ float4 probe_lds128(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ void probe_sts(uint32_t a, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory");
}

template <int LPR>
__global__ void __launch_bounds__(64 * LPR) matvec_probe_kernel(const float* __restrict__ C, int iters,
                                                                long long* cycles, float* ab_out) {
  constexpr int B = 64, EPT = B / LPR, PQ = EPT + 4;
  __shared__ __align__(16) float as[LPR * PQ], bs[LPR * PQ];
  __shared__ float hist[2][32][B];
  const int tid = threadIdx.x, i = tid / LPR, q = tid % LPR;
  float Kr[EPT], Kc[EPT];
#pragma unroll
  for (int e = 0; e < EPT; ++e) {
    Kr[e] = exp2f(-0.05f * fabsf(C[i * B + q * EPT + e] - 900.f));
    Kc[e] = exp2f(-0.05f * fabsf(C[(q * EPT + e) * B + i] - 900.f));
  }
  const int ip = (i / EPT) * PQ + (i % EPT);
  for (int t = tid; t < LPR * PQ; t += blockDim.x) { as[t] = 0.f; bs[t] = 0.f; }
  __syncthreads();
  if (q == 0) bs[ip] = 1.f;
  __syncthreads();
  const uint32_t as_q = probe_pin((uint32_t)__cvta_generic_to_shared(as) + q * PQ * 4);
  const uint32_t bs_q = probe_pin((uint32_t)__cvta_generic_to_shared(bs) + q * PQ * 4);
  const uint32_t as_i = probe_pin((uint32_t)__cvta_generic_to_shared(&as[ip]));
  const uint32_t bs_i = probe_pin((uint32_t)__cvta_generic_to_shared(&bs[ip]));
  const uint32_t h_i = probe_pin((uint32_t)__cvta_generic_to_shared(&hist[0][0][i]));
  const float c = 1.f / 64.f;
  auto dot = [&](const float (&Ks)[EPT], uint32_t a) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
    for (int e = 0; e < EPT; e += 4) {
      const float4 b = probe_lds128(a + e * 4);
      s0 = fmaf(Ks[e], b.x, s0); s1 = fmaf(Ks[e + 1], b.y, s1); s2 = fmaf(Ks[e + 2], b.z, s2); s3 = fmaf(Ks[e + 3], b.w, s3);
    }
    float s = (s0 + s1) + (s2 + s3);
#pragma unroll
    for (int o = 1; o < LPR; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    return s;
  };
  const long long t0 = clock64();
  float a_new = 0.f, b_new = 0.f;
  for (int it = 0; it < iters; ++it) {
    const float s = dot(Kr, bs_q);
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(s));
    a_new = c * r;
    float lg;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(s));
    if (q == 0) {
      probe_sts(as_i, a_new);
      probe_sts(h_i + (uint32_t)(it & 31) * (B * 4), -lg);
    }
    __syncthreads();
    const float t = dot(Kc, as_q);
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t));
    b_new = c * r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(t));
    if (q == 0) {
      probe_sts(bs_i, b_new);
      probe_sts(h_i + (uint32_t)(32 + (it & 31)) * (B * 4), -lg);
    }
    __syncthreads();
  }
  const long long t1 = clock64();
  if (tid == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
  if (q == 0) { ab_out[blockIdx.x * 128 + i] = a_new; ab_out[blockIdx.x * 128 + 64 + i] = b_new + hist[1][5][i]; }
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

extern "C" int kccot_debug_matvec_probe(int lpr, int iters, const float* C, long long* cycles, float* ab_out, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (lpr == 1) matvec_probe_kernel<1><<<3, 64, 0, st>>>(C, iters, cycles, ab_out);
  else if (lpr == 2) matvec_probe_kernel<2><<<3, 128, 0, st>>>(C, iters, cycles, ab_out);
  else matvec_probe_kernel<4><<<3, 256, 0, st>>>(C, iters, cycles, ab_out);
  return (int)cudaGetLastError();
}

// large-batch GEMM knobs: drain period in k-blocks (0 = kernel default), CTA-pair kernel on/off
extern "C" void kccot_debug_large_config(int drain_k_blocks, int use_pair) {
  kccot::large_set_drain(drain_k_blocks);
  kccot::large_set_pair(use_pair);
}
#endif  // KCCOT_DEV
