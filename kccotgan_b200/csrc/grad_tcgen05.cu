// Adjoint of the squared-distance cost on the tensor cores.
//
// With the stacked rows z = [x; y] (R = Bx + By <= 128) and a weight matrix W [R,R] built from the
// adjoints of the cost matrices, every wanted gradient row is
//     g_r = 2s * sum_c W_rc (z_r - z_c) = -2s * sum_c W'_rc z_c,   W' = W - diag(rowsum(W)),
// i.e. a skinny GEMM  G^T[col, r] = sum_c Z[c, col] * W'[r, c]  streamed once over the video columns.
//
//   A operand = the TMA-loaded video tile itself, MN-major.  tcgen05 accepts exactly one shared-memory
//               layout for MN-major 32-bit operands, SWIZZLE_128B_BASE32B (32-byte chunks XOR row%4,
//               4-row atoms): the tiles are loaded with the matching TMA mode 128B_ATOM_32B, so a
//               [rows x 32] box is a stack of canonical atoms (SBO = 512 B), and 4 boxes side by side
//               give M = 128 video columns (LBO = box bytes).
//   B operand = W' (K-major SW128), split tf32 hi/lo once per CTA and resident in shared memory.
//   3xTF32:    D += Zhi.Whi + Zlo.Whi + Zhi.Wlo, fp32 accumulate in TMEM (2 accumulator buffers).
//   Stages alternate between the x rows and the y rows of a 128-column tile (two TMA tensors).
//   Epilogue:  TMEM lane = video column, so for a fixed output row a warp writes 32 consecutive
//              floats: coalesced stores straight from registers.
#include "cost.cuh"
#include "tc_common.cuh"

namespace kccot {

namespace {
constexpr int kCols = 128;                 // video columns per work tile (UMMA M)
constexpr int kBoxCols = 32;               // fp32 columns per TMA box (128-byte swizzle row)
constexpr int kMaxN = 64;                  // output rows per launch (UMMA N)
constexpr int kConvWarps = 8;
constexpr int kConvThreads = kConvWarps * 32;
constexpr int kEpiWarps = 4;
constexpr int kThreads = 64 + kConvThreads + kEpiWarps * 32;   // 448
constexpr int kMaxStages = 6;
constexpr int kTmemCols = 128;             // 2 accumulator buffers x 64 columns

struct Bars {
  uint64_t full[kMaxStages], conv[kMaxStages], empty[kMaxStages];
  uint64_t acc_full[2], acc_empty[2];
  uint32_t tmem_base;
};

// builds W' for the mixed loss from Cbar3 [nprob,3,B,B] (xy, xx, yy; weights already applied)
__global__ void __launch_bounds__(128) build_w_mixed_kernel(const float* __restrict__ Cbar3, int B,
                                                            float* __restrict__ W) {
  const int R = 2 * B;
  const int p = blockIdx.y, r = blockIdx.x, c = threadIdx.x;
  const long long BB = (long long)B * B;
  const float* Cxy = Cbar3 + (long long)p * 3 * BB;
  const float* Cxx = Cxy + BB;
  const float* Cyy = Cxy + 2 * BB;
  float w = 0.f;
  if (c < R && c != r) {
    if (r < B) w = (c < B) ? Cxx[(long long)r * B + c] + Cxx[(long long)c * B + r] : Cxy[(long long)r * B + (c - B)];
    else w = (c < B) ? Cxy[(long long)c * B + (r - B)]
                     : Cyy[(long long)(r - B) * B + (c - B)] + Cyy[(long long)(c - B) * B + (r - B)];
  }
  __shared__ float red[4];
  float s = warp_sum(w);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  const float d = red[0] + red[1] + red[2] + red[3];
  if (c < R) W[((long long)p * R + r) * R + c] = (c == r) ? -d : w;
}

// generic pair: W'[i][Bx+j] = Cbar_ij, W'[Bx+j][i] = Cbar_ij
__global__ void __launch_bounds__(128) build_w_pair_kernel(const float* __restrict__ Cbar, int Bx, int By,
                                                           float* __restrict__ W) {
  const int R = Bx + By;
  const int p = blockIdx.y, r = blockIdx.x, c = threadIdx.x;
  const float* Cp = Cbar + (long long)p * Bx * By;
  float w = 0.f;
  if (c < R) {
    if (r < Bx && c >= Bx) w = Cp[(long long)r * By + (c - Bx)];
    else if (r >= Bx && c < Bx) w = Cp[(long long)c * By + (r - Bx)];
  }
  __shared__ float red[4];
  float s = warp_sum(w);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  const float d = red[0] + red[1] + red[2] + red[3];
  if (c < R) W[((long long)p * R + r) * R + c] = (c == r) ? -d : w;
}

__global__ void __launch_bounds__(kThreads, 1)
grad_tc_kernel(const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmy, int Bx, int By,
               long long K, const float* __restrict__ W, int row_off, int N, float neg2s, float* __restrict__ out,
               int accumulate, int nstages) {
  extern __shared__ uint8_t smem_raw[];
  // align by OFFSET so that the compiler keeps the shared address space (LDS/STS, not generic LD/ST)
  uint8_t* base = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int R = Bx + By;
  const int Npad = (N + 15) & ~15;
  const int wtiles = (R + 31) / 32;                 // 32-wide contraction tiles of W'
  const int wtile_bytes = Npad * 128;
  const int rows_max = max(Bx, By);
  const int box_bytes_max = rows_max * 128;
  const int stage_bytes = 4 * box_bytes_max;        // one half (x or y rows) of a 128-column tile
  // layout: W_hi | W_lo | stage hi[nstages] | stage lo[nstages] | barriers
  uint8_t* w_hi = base;
  uint8_t* w_lo = w_hi + wtiles * wtile_bytes;
  uint8_t* st_hi = w_lo + wtiles * wtile_bytes;
  uint8_t* st_lo = st_hi + (size_t)nstages * stage_bytes;
  Bars& bars = *reinterpret_cast<Bars*>(st_lo + (size_t)nstages * stage_bytes);

  const int p = blockIdx.y;
  const long long ntiles = (K + kCols - 1) / kCols;

  if (threadIdx.x == 0) {
    tc::prefetch_tmap(&tmx);
    tc::prefetch_tmap(&tmy);
    for (int s = 0; s < nstages; ++s) {
      tc::mbar_init(&bars.full[s], 1);
      tc::mbar_init(&bars.conv[s], kConvWarps);
      tc::mbar_init(&bars.empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      tc::mbar_init(&bars.acc_full[b], 1);
      tc::mbar_init(&bars.acc_empty[b], kEpiWarps);
    }
    tc::fence_barrier_init();
  }
  if (warp == 1) {
    tc::tmem_alloc(&bars.tmem_base, kTmemCols);
    tc::tmem_relinquish();
  }
  // W' -> shared memory, K-major 128-byte-swizzled tiles, split into tf32 hi / lo
  {
    const float* Wp = W + ((long long)p * R + row_off) * R;
    const int total = wtiles * Npad * 32;
    for (int e = threadIdx.x; e < total; e += kThreads) {
      const int cc = e & 31, r = (e >> 5) % Npad, tw = e / (32 * Npad);
      const int c = tw * 32 + cc;
      const float v = (r < N && c < R) ? Wp[(long long)r * R + c] : 0.f;
      const float h = tc::to_tf32(v);
      const float l = tc::to_tf32(v - h);
      const int off = tw * wtile_bytes + r * 128 + ((((cc >> 2) ^ (r & 7)) << 4) | ((cc & 3) << 2));
      *reinterpret_cast<float*>(w_hi + off) = h;
      *reinterpret_cast<float*>(w_lo + off) = l;
    }
  }
  tc::fence_proxy_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = bars.tmem_base;

  if (warp == 0) {
    // ------------------------------- TMA producer ---------------------------------------------
    if (tc::elect_one()) {
      int stage = 0, phase = 0;
      for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int col0 = (int)(t * kCols);
        for (int half = 0; half < 2; ++half) {
          const int rows = half ? By : Bx;
          if (rows == 0) continue;
          const CUtensorMap* tm = half ? &tmy : &tmx;
          tc::mbar_wait(&bars.empty[stage], phase ^ 1);
          tc::mbar_arrive_expect_tx(&bars.full[stage], (uint32_t)(4 * rows * 128));
          uint8_t* dst = st_hi + (size_t)stage * stage_bytes;
#pragma unroll
          for (int b = 0; b < 4; ++b) tc::tma_load_3d(tm, &bars.full[stage], dst + b * rows * 128, col0 + b * kBoxCols, 0, p);
          if (++stage == nstages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer -----------------------------------------------
    if (tc::elect_one()) {
      const uint32_t idesc = tc::make_idesc_tf32(kCols, Npad, /*A MN-major*/ 1, /*B K-major*/ 0);
      const uint32_t whi = tc::smem_u32(w_hi), wlo = tc::smem_u32(w_lo);
      int stage = 0, phase = 0;
      int ab = 0, ab_phase = 0;
      for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        tc::mbar_wait(&bars.acc_empty[ab], ab_phase ^ 1);
        tc::tc_fence_after();
        const uint32_t d_tmem = tmem + (uint32_t)(ab * kMaxN);
        bool first = true;
        for (int half = 0; half < 2; ++half) {
          const int rows = half ? By : Bx;
          if (rows == 0) continue;
          const int c_off = half ? Bx : 0;
          tc::mbar_wait(&bars.conv[stage], phase);
          tc::tc_fence_after();
          const uint32_t ahi = tc::smem_u32(st_hi + (size_t)stage * stage_bytes);
          const uint32_t alo = tc::smem_u32(st_lo + (size_t)stage * stage_bytes);
          const uint32_t lbo = (uint32_t)rows * 128;          // bytes between the 32-column boxes
          for (int kk = 0; kk < rows / 8; ++kk) {
            const int c = c_off + kk * 8;                     // contraction index of this k-step
            const uint32_t woff = (uint32_t)((c >> 5) * wtile_bytes + (c & 31) * 4);
            const uint64_t a_h = tc::make_smem_desc(ahi + kk * 1024, lbo, 512, 1);
            const uint64_t a_l = tc::make_smem_desc(alo + kk * 1024, lbo, 512, 1);
            const uint64_t b_h = tc::make_smem_desc_sw128(whi + woff, 16, 1024);
            const uint64_t b_l = tc::make_smem_desc_sw128(wlo + woff, 16, 1024);
            tc::umma_tf32(d_tmem, a_h, b_h, idesc, first ? 0u : 1u);
            tc::umma_tf32(d_tmem, a_l, b_h, idesc, 1u);
            tc::umma_tf32(d_tmem, a_h, b_l, idesc, 1u);
            first = false;
          }
          tc::umma_commit(&bars.empty[stage]);
          if (++stage == nstages) { stage = 0; phase ^= 1; }
        }
        tc::umma_commit(&bars.acc_full[ab]);
        if (++ab == 2) { ab = 0; ab_phase ^= 1; }
      }
    }
  } else if (warp < 2 + kConvWarps) {
    // ------------------------------- tf32 hi / lo split ---------------------------------------
    const int ct = threadIdx.x - 64;
    int stage = 0, phase = 0;
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
      for (int half = 0; half < 2; ++half) {
        const int rows = half ? By : Bx;
        if (rows == 0) continue;
        tc::mbar_wait(&bars.full[stage], phase);
        const uint32_t hi = tc::smem_u32(st_hi + (size_t)stage * stage_bytes);
        const uint32_t lo = tc::smem_u32(st_lo + (size_t)stage * stage_bytes);
        const int n16 = 4 * rows * 8;                         // 16-byte units in this stage
#pragma unroll 4
        for (int e = ct; e < n16; e += kConvThreads) {
          const float4 v = tc::lds128(hi + e * 16);
          float4 h, l;
          h.x = tc::to_tf32(v.x); h.y = tc::to_tf32(v.y); h.z = tc::to_tf32(v.z); h.w = tc::to_tf32(v.w);
          l.x = tc::to_tf32(v.x - h.x); l.y = tc::to_tf32(v.y - h.y);
          l.z = tc::to_tf32(v.z - h.z); l.w = tc::to_tf32(v.w - h.w);
          tc::sts128(hi + e * 16, h);
          tc::sts128(lo + e * 16, l);
        }
        tc::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&bars.conv[stage]);
        if (++stage == nstages) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ------------------------------- epilogue --------------------------------------------------
    const int quad = warp & 3;
    int ab = 0, ab_phase = 0;
    float* outp = out + (long long)p * (row_off == 0 ? Bx : By) * K;
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
      tc::mbar_wait(&bars.acc_full[ab], ab_phase);
      tc::tc_fence_after();
      const long long col = t * kCols + quad * 32 + lane;
      for (int g = 0; g < Npad; g += 32) {
        float d[32];
        tc::tmem_ld_32x32(tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(ab * kMaxN + g), d);
        tc::tmem_ld_wait();
        if (col < K) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int r = g + j;
            if (r < N) {
              float* dst = outp + (long long)r * K + col;
              const float v = neg2s * d[j];
              *dst = accumulate ? (*dst + v) : v;
            }
          }
        }
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&bars.acc_empty[ab]);
      if (++ab == 2) { ab = 0; ab_phase ^= 1; }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem, kTmemCols);
}

struct GradPlan { int nstages; size_t smem; };
GradPlan plan_grad(int Bx, int By, int N) {
  const int R = Bx + By;
  const int Npad = (N + 15) & ~15;
  const size_t wbytes = (size_t)2 * ((R + 31) / 32) * Npad * 128;
  const size_t stage = (size_t)2 * 4 * (Bx > By ? Bx : By) * 128;     // hi + lo
  const size_t budget = 227 * 1024 - 2048 - sizeof(Bars);
  int ns = (int)((budget - wbytes) / stage);
  if (ns > kMaxStages) ns = kMaxStages;
  GradPlan g;
  g.nstages = ns;
  g.smem = wbytes + ns * stage + sizeof(Bars) + 1024;
  return g;
}
}  // namespace

bool tc_grad_supported(const float* x, const float* y, int Bx, int By, long long K, const float* gx, const float* gy) {
  (void)gx; (void)gy;
  if (Bx % 8 || By % 8 || Bx + By > 128 || Bx > kMaxN || By > kMaxN || Bx < 8 || By < 8) return false;
  if (K % 4 != 0 || K < kCols) return false;
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(y) & 15)) return false;
  return plan_grad(Bx, By, Bx > By ? Bx : By).nstages >= 2;
}

// W: [nprob, R, R] already holds W' (diagonal = -rowsum)
static int launch_grad_rows(const CUtensorMap& tmx, const CUtensorMap& tmy, const float* W, int nprob, int Bx, int By,
                            long long K, float s, int row_off, int N, float* out, int accumulate, cudaStream_t st) {
  const GradPlan g = plan_grad(Bx, By, N);
  static size_t attr_smem = 0;
  if (g.smem > attr_smem) {
    KCCOT_CUDA(cudaFuncSetAttribute(grad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem));
    attr_smem = g.smem;
  }
  const long long ntiles = (K + kCols - 1) / kCols;
  int gx = (int)((num_sms() + nprob - 1) / nprob);
  if (gx > ntiles) gx = (int)ntiles;
  if (gx < 1) gx = 1;
  dim3 grid(gx, nprob);
  grad_tc_kernel<<<grid, kThreads, g.smem, st>>>(tmx, tmy, Bx, By, K, W, row_off, N, -2.f * s, out, accumulate,
                                                g.nstages);
  KCCOT_LAUNCH_CHECK();
  return KCCOT_OK;
}

int launch_grad_tc(const float* Cbar3, const float* x, const float* y, int nprob, int Bx, int By, long long K, float s,
                   float* gx, float* gy, int accumulate, float* Wws, cudaStream_t st) {
  const int R = Bx + By;
  build_w_mixed_kernel<<<dim3(R, nprob), 128, 0, st>>>(Cbar3, Bx, Wws);
  KCCOT_LAUNCH_CHECK();
  CUtensorMap tmx, tmy;
  if (int rc = encode_tmap_3d(&tmx, x, (uint64_t)K, (uint64_t)Bx, (uint64_t)nprob, (uint64_t)K * 4, (uint64_t)K * 4 * Bx,
                              kBoxCols, (uint32_t)Bx, true))
    return rc;
  if (int rc = encode_tmap_3d(&tmy, y, (uint64_t)K, (uint64_t)By, (uint64_t)nprob, (uint64_t)K * 4, (uint64_t)K * 4 * By,
                              kBoxCols, (uint32_t)By, true))
    return rc;
  if (gy)
    if (int rc = launch_grad_rows(tmx, tmy, Wws, nprob, Bx, By, K, s, Bx, By, gy, accumulate, st)) return rc;
  if (gx)
    if (int rc = launch_grad_rows(tmx, tmy, Wws, nprob, Bx, By, K, s, 0, Bx, gx, accumulate, st)) return rc;
  return KCCOT_OK;
}

int launch_grad_pair_tc(const float* Cbar, const float* x, const float* y, int nprob, int Bx, int By, long long K,
                        float s, float* gx, float* gy, int accumulate, float* Wws, cudaStream_t st) {
  const int R = Bx + By;
  build_w_pair_kernel<<<dim3(R, nprob), 128, 0, st>>>(Cbar, Bx, By, Wws);
  KCCOT_LAUNCH_CHECK();
  CUtensorMap tmx, tmy;
  if (int rc = encode_tmap_3d(&tmx, x, (uint64_t)K, (uint64_t)Bx, (uint64_t)nprob, (uint64_t)K * 4, (uint64_t)K * 4 * Bx,
                              kBoxCols, (uint32_t)Bx, true))
    return rc;
  if (int rc = encode_tmap_3d(&tmy, y, (uint64_t)K, (uint64_t)By, (uint64_t)nprob, (uint64_t)K * 4, (uint64_t)K * 4 * By,
                              kBoxCols, (uint32_t)By, true))
    return rc;
  if (gy)
    if (int rc = launch_grad_rows(tmx, tmy, Wws, nprob, Bx, By, K, s, Bx, By, gy, accumulate, st)) return rc;
  if (gx)
    if (int rc = launch_grad_rows(tmx, tmy, Wws, nprob, Bx, By, K, s, 0, Bx, gx, accumulate, st)) return rc;
  return KCCOT_OK;
}

}  // namespace kccot
