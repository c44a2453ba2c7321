// Tiled tensor-core GEMM of the large-batch cost path (B > 64: BASELINE config 5, B = 8192):
//
//     out[m][n] (+)= alpha * sum_k ( A1[m,k] B1[n,k] + A1[m,k] B2[n,k] + A2[m,k] B1[n,k] )
//
// A1/B1 are the fp16 "hi" halves and A2/B2 the fp16 "lo" halves of fp32 operands split by the pre-pass
// (large_prep.cu): x = hi + lo with 11 + 11 significant bits, the same mantissa width as TF32, so the three
// products reproduce 3xTF32 accuracy at the fp16 rate (2x the tf32 rate) and half the operand bytes.  The
// forward uses it as the Gram of the centred stacked rows (gan_utils.py:14-17 in GEMM form), the backward
// as W' . Z over the transposed split (the adjoint of the same lines).
//
//   tile      : 128 x 256 per CTA, fp32 accumulators in tensor memory: hi.hi in columns [0,256), the two
//               cross products in columns [256,512).  The cross terms are 2^-11 smaller; adding them into the
//               large accumulator would lose them to the tensor core's truncating fp32 adder.
//   pipeline  : warp 0 = TMA producer (4 boxes per k-block of 64 columns: A1, A2 [128 x 64], B1, B2 [256 x 64],
//               128-byte swizzle, 96 KB per stage, 2 stages), warp 1 = MMA issuer (12 tcgen05.mma.kind::f16
//               M128 x N256 x K16 per k-block), warps 2-9 = drain / epilogue.
//   chunks    : the tensor core accumulates with truncation, so a chain of thousands of updates into one
//               accumulator drifts (measured in round 1).  Every `drain` k-blocks the hi.hi
//               accumulator is read back and added to the output tile in global memory with round-to-nearest
//               fp32 adds (the tile stays L2-resident between chunks); the cross accumulator is drained once
//               per work item.
//   work item : (job, tile, k-split).  Up to 3 jobs share the operand tensor maps and differ in row offsets
//               (the xy, xx, yy blocks of the stacked Gram); symmetric jobs enumerate only tiles that touch
//               the upper triangle.  Tiles are ordered column-major so that concurrently running CTAs share
//               the B tile and walk down A (L2 reuse).  Persistent CTAs, static round-robin.
#include <cuda_fp16.h>

#include "cost.cuh"
#include "tc_common.cuh"

namespace kccot {

namespace {
constexpr int BM = G3_BM, BN = G3_BN, BK = G3_BK;
constexpr int kABytes = BM * BK * 2;                       // 16 KB
constexpr int kBBytes = BN * BK * 2;                       // 32 KB
constexpr int kStageBytes = 2 * kABytes + 2 * kBBytes;     // 96 KB
constexpr int kStages = 2;
constexpr int kEpiWarps = 8;
constexpr int kThreads = 64 + kEpiWarps * 32;              // 320
constexpr int kTmemCols = 512;
constexpr int kAccX = 256;                                 // TMEM column of the cross accumulator

struct Bars {
  uint64_t full[kStages], empty[kStages];
  uint64_t acc_full, acc_free;
  uint32_t tmem_base;
};

struct Item {
  int job, m0, n0, ks;
};

__device__ __forceinline__ bool decode_item(const G3Params& P, int w, Item& it) {
  int tile = w / P.ksplit;
  it.ks = w - tile * P.ksplit;
  for (int j = 0; j < P.njobs; ++j) {
    const G3Job& jb = P.job[j];
    if (tile < jb.ntiles) {
      it.job = j;
      int tni, tmi;
      if (!jb.tri) {
        tni = tile / jb.tm;
        tmi = tile - tni * jb.tm;
      } else {
        // symmetric block: column tile t needs the row tiles that reach the diagonal, tmi < min(tm, 2t + 2)
        tni = 0;
        for (;; ++tni) {
          const int cnt = min(jb.tm, 2 * tni + 2);
          if (tile < cnt) break;
          tile -= cnt;
        }
        tmi = tile;
      }
      it.m0 = tmi * BM;
      it.n0 = tni * BN;
      return true;
    }
    tile -= jb.ntiles;
  }
  return false;
}

// one accumulator (256 TMEM columns, this warp's 32 lanes x its 128-column half) -> out (+)= alpha * acc
__device__ __forceinline__ void drain_acc(uint32_t tmem_acc, int quad, int half, int lane, float* __restrict__ orow,
                                          bool row_ok, int ncols_left, float alpha, bool store, bool vec_ok) {
#pragma unroll 1
  for (int cc = 0; cc < 4; ++cc) {
    const int col0 = half * 128 + cc * 32;
    const int ncols = ncols_left - col0;                  // valid columns from col0 on (warp-uniform)
    if (ncols <= 0) break;
    float d[32];
    tc::tmem_ld_32x32(tmem_acc + ((uint32_t)(quad * 32) << 16) + (uint32_t)col0, d);
    tc::tmem_ld_wait();
    if (row_ok) {
      float* o = orow + col0;
      if (vec_ok && ncols >= 32) {
        float4 v[8];
        if (!store) {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = reinterpret_cast<const float4*>(o)[j];
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          v[j].x = fmaf(alpha, d[4 * j + 0], v[j].x);
          v[j].y = fmaf(alpha, d[4 * j + 1], v[j].y);
          v[j].z = fmaf(alpha, d[4 * j + 2], v[j].z);
          v[j].w = fmaf(alpha, d[4 * j + 3], v[j].w);
          reinterpret_cast<float4*>(o)[j] = v[j];
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < ncols) o[j] = store ? alpha * d[j] : fmaf(alpha, d[j], o[j]);
      }
    }
  }
}
}  // namespace

__global__ void __launch_bounds__(kThreads, 1)
gemm_f16x3_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmA2,
                  const __grid_constant__ CUtensorMap tmB1, const __grid_constant__ CUtensorMap tmB2,
                  const __grid_constant__ G3Params P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  Bars& bars = *reinterpret_cast<Bars*>(base + kStages * kStageBytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nitems = P.ntiles_total * P.ksplit;

  if (threadIdx.x == 0) {
    tc::prefetch_tmap(&tmA1);
    tc::prefetch_tmap(&tmA2);
    tc::prefetch_tmap(&tmB1);
    tc::prefetch_tmap(&tmB2);
    for (int s = 0; s < kStages; ++s) {
      tc::mbar_init(&bars.full[s], 1);
      tc::mbar_init(&bars.empty[s], 1);
    }
    tc::mbar_init(&bars.acc_full, 1);
    tc::mbar_init(&bars.acc_free, kEpiWarps);
    tc::fence_barrier_init();
  }
  if (warp == 1) {
    tc::tmem_alloc(&bars.tmem_base, kTmemCols);
    tc::tmem_relinquish();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = bars.tmem_base;

  if (warp == 0) {
    // ------------------------------- TMA producer ---------------------------------------------
    if (tc::elect_one()) {
      int stage = 0, phase = 0;
      Item it;
      for (int w = blockIdx.x; w < nitems; w += gridDim.x) {
        if (!decode_item(P, w, it)) break;
        const G3Job& jb = P.job[it.job];
        const int kb0 = it.ks * P.kb_per_split, kb1 = min(P.nkb, kb0 + P.kb_per_split);
        const int arow = jb.a_row0 + it.m0, brow = jb.b_row0 + it.n0;
        for (int kb = kb0; kb < kb1; ++kb) {
          tc::mbar_wait(&bars.empty[stage], phase ^ 1);
          uint8_t* sb = base + (size_t)stage * kStageBytes;
          tc::mbar_arrive_expect_tx(&bars.full[stage], (uint32_t)kStageBytes);
          tc::tma_load_2d(&tmA1, &bars.full[stage], sb, kb * BK, arow);
          tc::tma_load_2d(&tmB1, &bars.full[stage], sb + 2 * kABytes, kb * BK, brow);
          tc::tma_load_2d(&tmB2, &bars.full[stage], sb + 2 * kABytes + kBBytes, kb * BK, brow);
          tc::tma_load_2d(&tmA2, &bars.full[stage], sb + kABytes, kb * BK, arow);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
      // drain the asynchronous tcgen05.commit arrivals on empty[] before the CTA may exit (see grad_tcgen05.cu)
      for (int i = 0; i < kStages; ++i) {
        tc::mbar_wait(&bars.empty[stage], phase ^ 1);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer -----------------------------------------------
    if (tc::elect_one()) {
      const uint32_t idesc = tc::make_idesc_f16(BM, BN, 0, 0);
      int stage = 0, phase = 0;
      uint32_t g = 0;                                       // running chunk number (acc_full / acc_free phases)
      Item it;
      for (int w = blockIdx.x; w < nitems; w += gridDim.x) {
        if (!decode_item(P, w, it)) break;
        const int kb0 = it.ks * P.kb_per_split, kb1 = min(P.nkb, kb0 + P.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          const int rel = kb - kb0;
          const bool chunk_first = (rel % P.drain) == 0;
          const bool chunk_last = ((rel + 1) % P.drain) == 0 || kb == kb1 - 1;
          if (chunk_first) {                                 // the drain warps have read the accumulators back
            tc::mbar_wait(&bars.acc_free, (g & 1u) ^ 1u);
            tc::tc_fence_after();
          }
          tc::mbar_wait(&bars.full[stage], phase);
          tc::tc_fence_after();
          const uint32_t sb = tc::smem_u32(base + (size_t)stage * kStageBytes);
#pragma unroll
          for (int k4 = 0; k4 < BK / 16; ++k4) {
            const uint64_t a1 = tc::make_smem_desc_sw128(sb + k4 * 32, 16, 1024);
            const uint64_t a2 = tc::make_smem_desc_sw128(sb + kABytes + k4 * 32, 16, 1024);
            const uint64_t b1 = tc::make_smem_desc_sw128(sb + 2 * kABytes + k4 * 32, 16, 1024);
            const uint64_t b2 = tc::make_smem_desc_sw128(sb + 2 * kABytes + kBBytes + k4 * 32, 16, 1024);
            tc::umma_f16(tmem, a1, b1, idesc, (chunk_first && k4 == 0) ? 0u : 1u);
            tc::umma_f16(tmem + kAccX, a1, b2, idesc, (rel == 0 && k4 == 0) ? 0u : 1u);
            tc::umma_f16(tmem + kAccX, a2, b1, idesc, 1u);
          }
          tc::umma_commit(&bars.empty[stage]);
          if (chunk_last) {
            tc::umma_commit(&bars.acc_full);
            ++g;
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ------------------------------- drain / epilogue -----------------------------------------
    const int ew = warp - 2;
    const int quad = warp & 3;                              // TMEM lane quadrant this warp may read
    const int half = ew >> 2;                               // columns [128 half, 128 half + 128)
    const float alpha = P.alpha * (P.alpha_dev ? *P.alpha_dev : 1.f);
    uint32_t g = 0;
    Item it;
    for (int w = blockIdx.x; w < nitems; w += gridDim.x) {
      if (!decode_item(P, w, it)) break;
      const G3Job& jb = P.job[it.job];
      const int kb0 = it.ks * P.kb_per_split, kb1 = min(P.nkb, kb0 + P.kb_per_split);
      const int nchunks = (kb1 - kb0 + P.drain - 1) / P.drain;
      const int row = it.m0 + quad * 32 + lane;
      const bool row_ok = row < jb.m;
      float* orow = jb.out + (long long)it.ks * jb.ks_stride + (long long)row * jb.ld + it.n0;
      const int ncols_left = jb.n - it.n0;
      const bool vec_ok = ((jb.ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(jb.out) & 15) == 0) &&
                          (((long long)it.ks * jb.ks_stride & 3) == 0);
      for (int c = 0; c < nchunks; ++c) {
        tc::mbar_wait(&bars.acc_full, g & 1u);
        tc::tc_fence_after();
        drain_acc(tmem, quad, half, lane, orow, row_ok, ncols_left, alpha, c == 0 && !P.accumulate, vec_ok);
        if (c == nchunks - 1) drain_acc(tmem + kAccX, quad, half, lane, orow, row_ok, ncols_left, alpha, false, vec_ok);
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&bars.acc_free);
        ++g;
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem, kTmemCols);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
int g3_count_tiles(G3Job* jb) {
  jb->tm = (jb->m + BM - 1) / BM;
  jb->tn = (jb->n + BN - 1) / BN;
  if (!jb->tri) {
    jb->ntiles = jb->tm * jb->tn;
  } else {
    int n = 0;
    for (int t = 0; t < jb->tn; ++t) n += (jb->tm < 2 * t + 2) ? jb->tm : 2 * t + 2;
    jb->ntiles = n;
  }
  return jb->ntiles;
}

// k-split so that the grid is filled when the tile count is small (mid-size B); each split keeps >= 8 k-blocks
void g3_plan_split(int ntiles, int nkb, int units, int* ksplit, int* kb_per_split) {
  const int sms = units;
  int ks = 1;
  if (ntiles < sms) {
    ks = (sms + ntiles - 1) / ntiles;
    const int maxks = nkb / 8 > 1 ? nkb / 8 : 1;
    if (ks > maxks) ks = maxks;
  }
  int per = (nkb + ks - 1) / ks;
  *ksplit = (nkb + per - 1) / per;
  *kb_per_split = per;
}

int launch_gemm_f16x3(const __half* A1, const __half* A2, long long a_rows, long long a_pitch_elems,
                      const __half* B1, const __half* B2, long long b_rows, long long b_pitch_elems, long long kdim,
                      G3Params P, cudaStream_t st) {
  CUtensorMap tA1, tA2, tB1, tB2;
  if (int rc = encode_tmap_2d_f16(&tA1, A1, (uint64_t)kdim, (uint64_t)a_rows, (uint64_t)a_pitch_elems * 2, BK, BM)) return rc;
  if (int rc = encode_tmap_2d_f16(&tA2, A2, (uint64_t)kdim, (uint64_t)a_rows, (uint64_t)a_pitch_elems * 2, BK, BM)) return rc;
  if (int rc = encode_tmap_2d_f16(&tB1, B1, (uint64_t)kdim, (uint64_t)b_rows, (uint64_t)b_pitch_elems * 2, BK, BN)) return rc;
  if (int rc = encode_tmap_2d_f16(&tB2, B2, (uint64_t)kdim, (uint64_t)b_rows, (uint64_t)b_pitch_elems * 2, BK, BN)) return rc;
  P.nkb = (int)((kdim + BK - 1) / BK);
  P.ntiles_total = 0;
  for (int j = 0; j < P.njobs; ++j) P.ntiles_total += P.job[j].ntiles;
  if (P.ntiles_total == 0) return KCCOT_OK;
  if (P.drain < 1) P.drain = 16;
  const size_t smem = (size_t)kStages * kStageBytes + sizeof(Bars) + 1024;
  static size_t attr_set[kMaxDevices] = {};
  if (smem_attr_needed(attr_set, smem))
    KCCOT_CUDA(cudaFuncSetAttribute(gemm_f16x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long nitems = (long long)P.ntiles_total * P.ksplit;
  const int grid = (int)(nitems < num_sms() ? nitems : num_sms());
  gemm_f16x3_kernel<<<grid, kThreads, smem, st>>>(tA1, tA2, tB1, tB2, P);
  KCCOT_LAUNCH_CHECK();
  return KCCOT_OK;
}

}  // namespace kccot
