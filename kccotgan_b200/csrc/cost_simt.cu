// CUDA-core (fp32) kernels of the cost path: pairwise squared distances in the reference's direct
// (x - y)^2 form (gan_utils.py:14-17), the cost finalisation that adds the martingale terms
// (gan_utils.py:34-43, 59-72), and the adjoints.  Any shape, any alignment.  The tensor-core
// kernels in gram_tcgen05.cu replace the distance part when the shape allows; these remain the
// generic-shape path and the on-device cross-check.
#include "common.cuh"
#include <type_traits>

#include "cost.cuh"

namespace kccot {

// ---------------------------------------------------------------------------------------------
// P[p,ks,i,j] = sum_{k in slab ks} (x[p,i,k] - y[p,j,k])^2
// ---------------------------------------------------------------------------------------------
constexpr int ST = 64;   // output tile edge
constexpr int SK = 16;   // k chunk

__global__ void __launch_bounds__(256) sqdist_partial_kernel(const float* __restrict__ x,
                                                             const float* __restrict__ y, int Bx, int By,
                                                             long long K, long long kslab, int ksplit,
                                                             float* __restrict__ part) {
  __shared__ float xs[SK][ST + 4];
  __shared__ float ys[SK][ST + 4];
  const int p = blockIdx.z / ksplit, ks = blockIdx.z % ksplit;
  const long long k0 = (long long)ks * kslab;
  const long long k1 = min(K, k0 + kslab);
  const float* xp = x + (long long)p * Bx * K;
  const float* yp = y + (long long)p * By * K;
  const int i0 = blockIdx.y * ST, j0 = blockIdx.x * ST;
  const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
  const int lrow = t >> 2, lk = (t & 3) * 4;
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;

  for (long long kk = k0; kk < k1; kk += SK) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const long long k = kk + lk + e;
      const bool kin = k < k1;
      xs[lk + e][lrow] = (kin && i0 + lrow < Bx) ? xp[(long long)(i0 + lrow) * K + k] : 0.f;
      ys[lk + e][lrow] = (kin && j0 + lrow < By) ? yp[(long long)(j0 + lrow) * K + k] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&xs[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&ys[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int n = 0; n < 4; ++n) {
          const float d = av[m] - bv[n];
          acc[m][n] = fmaf(d, d, acc[m][n]);
        }
    }
    __syncthreads();
  }
  float* out = part + ((long long)p * ksplit + ks) * Bx * By;
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    const int i = i0 + ty * 4 + m;
    if (i >= Bx) continue;
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      const int j = j0 + tx * 4 + n;
      if (j < By) out[(long long)i * By + j] = acc[m][n];
    }
  }
}

int launch_sqdist_partials_simt(const float* x, const float* y, int nprob, int Bx, int By, long long K,
                                int ksplit, long long kslab, float* part, cudaStream_t st) {
  dim3 grid((By + ST - 1) / ST, (Bx + ST - 1) / ST, nprob * ksplit);
  sqdist_partial_kernel<<<grid, 256, 0, st>>>(x, y, Bx, By, K, kslab, ksplit, part);
  KCCOT_LAUNCH_CHECK();
  return KCCOT_OK;
}

void choose_ksplit_simt(int nprob, int Bx, int By, long long K, int* ksplit, long long* kslab) {
  const long long tiles = (long long)nprob * ((Bx + ST - 1) / ST) * ((By + ST - 1) / ST);
  long long want = (2LL * num_sms() + tiles - 1) / tiles;
  const long long maxsplit = max(1LL, K / 512);
  want = max(1LL, min(want, maxsplit));
  long long slab = (K + want - 1) / want;
  slab = (slab + SK - 1) / SK * SK;
  *ksplit = (int)((K + slab - 1) / slab);
  *kslab = slab;
}

// ---------------------------------------------------------------------------------------------
// C = s * sum_ks P + s * sum_pairs h_row . DeltaM_col
//
// One CTA = one 8 x 8 tile of one output block (and, for `sym` partials, its mirror tile), G k-slab
// groups of 128 threads.  Within a group 64 threads read the tile as stored and 64 read the mirror
// tile, eight consecutive floats (one 32-byte sector) per row, so every sector fetched from L2 is
// fully used although k-slabs are 64 KB apart.  `sym` partials (tensor-core path) hold
// P'_ij = n_i + n_j - 2 (H_ij + 2 X_ij) with X = hi lo^T only: the missing lo hi^T = X^T comes back as
// C_ij = (P'_ij + P'_ji) / 2.  The split-K partials are summed in fp64: a sequential fp32 sum of
// ~150-300 partials of a value near 1e4 would by itself cost ~5e-4 absolute on C, more than the whole
// fp32 budget of the path.
// ---------------------------------------------------------------------------------------------
constexpr int FT = 8;     // tile edge
constexpr int FGmax = 8;  // k-slab groups per CTA

__global__ void __launch_bounds__(128 * FGmax) cost_finalize_kernel(CostBlocks blocks, int T, int J, float s) {
  __shared__ double red[FGmax][2][FT * FT];
  pdl_wait();                    // the partial tiles come from the kernel before
  pdl_launch_dependents();
  if (blocks.zero != nullptr && blockIdx.x == 0 && blockIdx.z == 0 && threadIdx.x == 0) blocks.zero[blockIdx.y] = 0;
  const CostBlock& b = blocks.b[blockIdx.z];
  const int tiles_j = (b.By + FT - 1) / FT, tiles_i = (b.Bx + FT - 1) / FT;
  if ((int)blockIdx.x >= tiles_i * tiles_j) return;
  const int p = blockIdx.y;
  const int ti = blockIdx.x / tiles_j, tj = blockIdx.x % tiles_j;
  const int G = blockDim.x >> 7;
  const int tid = threadIdx.x, g = tid >> 7, mirror = (tid >> 6) & 1, r = (tid >> 3) & 7, c = tid & 7;
  // the output element this thread contributes to
  const int ei = ti * FT + (mirror ? c : r), ej = tj * FT + (mirror ? r : c);
  const bool live = ei < b.Bx && ej < b.By;
  double d = 0.0;
  if (live && !(b.zero_diag && ei == ej) && (!mirror || b.sym)) {
    const float* pp = b.part + (long long)p * b.prob_stride +
                      (mirror ? (long long)(b.col_off + ej) * b.ld + b.row_off + ei
                              : (long long)(b.row_off + ei) * b.ld + b.col_off + ej);
    constexpr int kBatch = 8;
    for (int ks0 = g; ks0 < b.nks; ks0 += G * kBatch) {
      float v[kBatch];
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        const int ks = ks0 + u * G;
        v[u] = (ks < b.nks) ? pp[(long long)ks * b.ks_stride] : 0.f;
      }
#pragma unroll
      for (int u = 0; u < kBatch; ++u) d += (double)v[u];
    }
    if (b.sym) d *= 0.5;
  }
  const int tj1 = (T - 1) * J;
  float a = 0.f;
  if (live) {
    for (int pair = 0; pair < 2; ++pair) {
      const float* h = pair ? b.h2 : b.h1;
      const float* M = pair ? b.M2 : b.M1;
      if (h == nullptr) continue;
      const float* hr = h + ((long long)p * b.Bx + ei) * T * J;
      const float* Mr = M + ((long long)p * b.By + ej) * T * J;
      for (int q = g * 2 + mirror; q < tj1; q += 2 * G) a = fmaf(hr[q], Mr[q + J] - Mr[q], a);
    }
  }
  red[g][mirror][mirror ? c * FT + r : r * FT + c] = d + (double)a;
  __syncthreads();
  if (tid < FT * FT) {
    const int oi = ti * FT + (tid >> 3), oj = tj * FT + (tid & 7);
    if (oi < b.Bx && oj < b.By) {
      double t = 0.0;
      for (int gg = 0; gg < G; ++gg) t += red[gg][0][tid] + red[gg][1][tid];      // fixed order: deterministic
      b.C[(long long)p * b.C_prob_stride + (long long)oi * b.By + oj] = (float)((double)s * t);
    }
  }
}

// Same result for the tensor-core partial layout part[p][16 x 16 tiles][ks][8][8]: the nks partials of an
// output tile (and of its mirror tile) are two contiguous streams of nks x 256 bytes, read with coalesced
// 16-byte loads in two batches; the martingale operands are staged in shared memory by the same threads, so
// the kernel is two L2 round trips deep instead of eight (the scattered-sector version above took 11 us in
// the evaluation chain).  512 threads: 256 on the tile as stored, 256 on the mirror tile; thread = (k-slab
// group g of 16, 16-byte slot q of 16).
constexpr int kFinMaxTJ = 256;      // (T-1)*J limit of the staged martingale operands

__global__ void __launch_bounds__(512, 2) cost_finalize_tiled_kernel(CostBlocks blocks, int T, int J, float s) {
  __shared__ double red[2][16][16][4];       // [as stored / mirror][k-slab group][slot][component]
  __shared__ double redm[8][FT * FT];        // martingale partial sums
  __shared__ float hs[FT][kFinMaxTJ], dms[FT][kFinMaxTJ];
  pdl_wait();                                // the partial tiles come from the kernel before
  pdl_launch_dependents();
  if (blocks.zero != nullptr && blockIdx.x == 0 && blockIdx.z == 0 && threadIdx.x == 0) blocks.zero[blockIdx.y] = 0;
  const CostBlock& b = blocks.b[blockIdx.z];
  const int tiles_j = b.By / FT;
  const int p = blockIdx.y;
  const int ti = blockIdx.x / tiles_j, tj = blockIdx.x % tiles_j;
  const int i0 = ti * FT, j0 = tj * FT;
  const int tid = threadIdx.x, half = tid >> 8, t = tid & 255, g = t >> 4, q = t & 15;
  const int tj1 = (T - 1) * J, TJ = T * J;
  const int tr = (b.row_off + i0) >> 3, tc = (b.col_off + j0) >> 3;
  const int tile = half ? tc * 16 + tr : tr * 16 + tc;
  const float4* pp = reinterpret_cast<const float4*>(b.part + ((long long)p * 256 + tile) * b.nks * 64) + q;
  // ---- two batches of coalesced loads; the staging loads of the martingale operands go out in between ----
  constexpr int kBatch = 5;
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  float4 v[kBatch];
#pragma unroll
  for (int u = 0; u < kBatch; ++u) {
    const int ks = g + u * 16;
    v[u] = (ks < b.nks) ? pp[ks * 16] : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  // h rows i0.., DeltaM rows j0.., columns [c0, c0 + n) -> shared memory
  auto stage = [&](const float* h, const float* M, int c0, int n) {
    for (int e = tid; e < FT * n; e += 512) {
      const int r = e / n, c = e - r * n;
      const float* Mr = M + ((long long)p * b.By + j0 + r) * TJ + c0 + c;
      hs[r][c] = h[((long long)p * b.Bx + i0 + r) * TJ + c0 + c];
      dms[r][c] = Mr[J] - Mr[0];
    }
  };
  const int n_first = min(tj1, kFinMaxTJ);
  if (b.h1) stage(b.h1, b.M1, 0, n_first);
#pragma unroll
  for (int u = 0; u < kBatch; ++u) { a0 += (double)v[u].x; a1 += (double)v[u].y; a2 += (double)v[u].z; a3 += (double)v[u].w; }
  for (int k0 = kBatch * 16; k0 < b.nks; k0 += kBatch * 16) {
#pragma unroll
    for (int u = 0; u < kBatch; ++u) {
      const int ks = k0 + g + u * 16;
      v[u] = (ks < b.nks) ? pp[ks * 16] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < kBatch; ++u) { a0 += (double)v[u].x; a1 += (double)v[u].y; a2 += (double)v[u].z; a3 += (double)v[u].w; }
  }
  red[half][g][q][0] = a0; red[half][g][q][1] = a1; red[half][g][q][2] = a2; red[half][g][q][3] = a3;
  // ---- martingale term: sum_pairs h_row . DeltaM_col out of shared memory ---------------------------
  const int e = tid & 63, part = tid >> 6;                  // 8 parts x 64 elements
  const int er = e >> 3, ec = e & 7;
  double am = 0.0;
  for (int pr = 0; pr < 2; ++pr) {
    const float* h = pr ? b.h2 : b.h1;
    const float* M = pr ? b.M2 : b.M1;
    if (h == nullptr) continue;
    for (int c0 = 0; c0 < tj1; c0 += kFinMaxTJ) {
      const int n = min(kFinMaxTJ, tj1 - c0);
      if (pr == 1 || c0 > 0) {
        __syncthreads();                                     // previous chunk consumed
        stage(h, M, c0, n);
      }
      __syncthreads();
      float a = 0.f;
      for (int c = part; c < n; c += 8) a = fmaf(hs[er][c], dms[ec][c], a);
      am += (double)a;
    }
  }
  redm[part][e] = am;
  __syncthreads();
  if (tid < FT * FT) {
    const int r = tid >> 3, c = tid & 7;
    double d = 0.0;
    if (!(b.zero_diag && i0 + r == j0 + c)) {
      const int qd = r * 2 + (c >> 2), cd = c & 3;           // (r, c) in the tile as stored
      const int qm = c * 2 + (r >> 2), cm = r & 3;           // (c, r) in the mirror tile
      for (int gg = 0; gg < 16; ++gg) d += red[0][gg][qd][cd] + red[1][gg][qm][cm];   // fixed order: deterministic
      d *= 0.5;
    }
    for (int gg = 0; gg < 8; ++gg) d += redm[gg][tid];
    b.C[(long long)p * b.C_prob_stride + (long long)(i0 + r) * b.By + j0 + c] = (float)((double)s * d);
  }
}

// Many problems with one or two k-slabs each (BASELINE config 4: 256 problems, nks = 1): the tile kernel above would
// launch 64 x nprob x 3 CTAs of 512 threads that each sum ONE partial (49 152 CTAs, 733 us at config 4).  Here one CTA
// of 256 threads finishes a whole B x B block (B <= 64) of one problem: the h rows and DeltaM rows are staged in shared
// memory once, thread (i = tid / 4, lane4 = tid % 4) owns row i and the columns lane4, lane4 + 4, ...
constexpr int kFinSmallMaxB = 64;
__global__ void __launch_bounds__(256) cost_finalize_small_kernel(CostBlocks blocks, int T, int J, float s) {
  extern __shared__ __align__(16) float fs_sh[];
  pdl_wait();                                // the partial tiles come from the kernel before
  pdl_launch_dependents();
  const CostBlock& b = blocks.b[blockIdx.y];
  const int p = blockIdx.x, tid = threadIdx.x;
  if (blocks.zero != nullptr && blockIdx.y == 0 && tid == 0) blocks.zero[p] = 0;
  // row pitch: a multiple of 4 floats (16-byte loads along the contraction) that is 12 mod 32 when (T-1)J is a
  // multiple of 8, so the eight rows of a warp's 16-byte loads fall into different banks
  const int tj1 = (T - 1) * J, TJ = T * J, tj4 = (tj1 + 3) & ~3, ldh = tj4 + 4;
  float* hs = fs_sh;                          // [Bx][ldh], zero-padded to tj4
  float* dms = fs_sh + kFinSmallMaxB * ldh;   // [By][ldh]
  const int i = tid >> 2, l4 = tid & 3;
  double acc[kFinSmallMaxB / 4];
#pragma unroll
  for (int m = 0; m < kFinSmallMaxB / 4; ++m) acc[m] = 0.0;
  // ---- distance partials: (P_ij + P'_ji) / 2 summed over the k-slabs in fp64 ----
  if (i < b.Bx) {
    const float* pb = b.part + (long long)p * 256 * b.nks * 64;
    // four outputs at a time: all their loads (2 x nks <= 8 each) go out before the first fp64 add
    constexpr int MB = 4;
#pragma unroll
    for (int m0 = 0; m0 < kFinSmallMaxB / 4; m0 += MB) {
      float vd[MB][4], vm[MB][4];
#pragma unroll
      for (int q = 0; q < MB; ++q) {
        const int j = l4 + 4 * (m0 + q);
        const bool ok = j < b.By && !(b.zero_diag && i == j);
        const int gi = b.row_off + i, gj = b.col_off + (ok ? j : 0);
        const float* pd = pb + ((long long)((gi >> 3) * 16 + (gj >> 3)) * b.nks) * 64 + (gi & 7) * 8 + (gj & 7);
        const float* pm = pb + ((long long)((gj >> 3) * 16 + (gi >> 3)) * b.nks) * 64 + (gj & 7) * 8 + (gi & 7);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const bool lk = ok && ks < b.nks;
          vd[q][ks] = lk ? pd[ks * 64] : 0.f;
          vm[q][ks] = lk ? pm[ks * 64] : 0.f;
        }
      }
#pragma unroll
      for (int q = 0; q < MB; ++q) {
        double d = 0.0;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) d += (double)vd[q][ks] + (double)vm[q][ks];
        acc[m0 + q] = 0.5 * d;
      }
    }
  }
  // ---- martingale terms: thread (i, l4) owns row i and the columns l4, l4 + 4, ...; one 16-byte load of its h row
  //      and one per column feed 4 FMAs each ----
  for (int pr = 0; pr < 2; ++pr) {
    const float* h = pr ? b.h2 : b.h1;
    const float* M = pr ? b.M2 : b.M1;
    if (h == nullptr) continue;
    __syncthreads();
    // staging with six loads in flight per thread (one load -> one store per iteration made the kernel a chain of
    // 36 memory latencies: 120 us for 256 problems)
    constexpr int SU = 6;
    for (int e0 = tid; e0 < b.Bx * tj4; e0 += 256 * SU) {
      float v[SU];
#pragma unroll
      for (int u = 0; u < SU; ++u) {
        const int e = e0 + 256 * u, r = e / tj4, c = e - r * tj4;
        v[u] = (e < b.Bx * tj4 && c < tj1) ? h[((long long)p * b.Bx + r) * TJ + c] : 0.f;
      }
#pragma unroll
      for (int u = 0; u < SU; ++u) {
        const int e = e0 + 256 * u, r = e / tj4, c = e - r * tj4;
        if (e < b.Bx * tj4) hs[r * ldh + c] = v[u];
      }
    }
    for (int e0 = tid; e0 < b.By * tj4; e0 += 256 * SU) {
      float v0[SU], v1[SU];
#pragma unroll
      for (int u = 0; u < SU; ++u) {
        const int e = e0 + 256 * u, r = e / tj4, c = e - r * tj4;
        const bool ok = e < b.By * tj4 && c < tj1;
        const float* Mr = M + ((long long)p * b.By + (ok ? r : 0)) * TJ + (ok ? c : 0);
        v0[u] = ok ? Mr[0] : 0.f;
        v1[u] = ok ? Mr[J] : 0.f;
      }
#pragma unroll
      for (int u = 0; u < SU; ++u) {
        const int e = e0 + 256 * u, r = e / tj4, c = e - r * tj4;
        if (e < b.By * tj4) dms[r * ldh + c] = v1[u] - v0[u];
      }
    }
    __syncthreads();
    if (i < b.Bx) {
      float a[kFinSmallMaxB / 4];
#pragma unroll
      for (int m = 0; m < kFinSmallMaxB / 4; ++m) a[m] = 0.f;
      const float4* hr = reinterpret_cast<const float4*>(hs + i * ldh);
      for (int c4 = 0; c4 < tj4 / 4; ++c4) {
        const float4 hv = hr[c4];
#pragma unroll
        for (int m = 0; m < kFinSmallMaxB / 4; ++m) {
          const int j = min(l4 + 4 * m, b.By - 1);
          const float4 dv = reinterpret_cast<const float4*>(dms + j * ldh)[c4];
          a[m] = fmaf(hv.x, dv.x, fmaf(hv.y, dv.y, fmaf(hv.z, dv.z, fmaf(hv.w, dv.w, a[m]))));
        }
      }
#pragma unroll
      for (int m = 0; m < kFinSmallMaxB / 4; ++m) acc[m] += (double)a[m];
    }
  }
  if (i < b.Bx) {
#pragma unroll
    for (int m = 0; m < kFinSmallMaxB / 4; ++m) {
      const int j = l4 + 4 * m;
      if (j < b.By) b.C[(long long)p * b.C_prob_stride + (long long)i * b.By + j] = (float)((double)s * acc[m]);
    }
  }
}

int launch_cost_finalize(const CostBlocks& blocks, int nblocks, int nprob, int T, int J, float s,
                         cudaStream_t st) {
  int tiles = 0, nks = 1;
  for (int i = 0; i < nblocks; ++i) {
    const CostBlock& b = blocks.b[i];
    tiles = max(tiles, ((b.Bx + FT - 1) / FT) * ((b.By + FT - 1) / FT));
    nks = max(nks, b.nks);
  }
  const int G = nks >= 64 ? 8 : nks >= 16 ? 4 : nks >= 4 ? 2 : 1;
  bool all_tiled = true;
  for (int i = 0; i < nblocks; ++i) {
    const CostBlock& b = blocks.b[i];
    if (!b.tiled || b.Bx % FT || b.By % FT || b.row_off % FT || b.col_off % FT) all_tiled = false;
  }
  bool small_ok = all_tiled && nks <= 4 && nprob >= 16 && (T - 1) * J <= 384;      // (staged rows: <= 199 KB of shared memory)
  for (int i = 0; i < nblocks; ++i)
    if (blocks.b[i].Bx > kFinSmallMaxB || blocks.b[i].By > kFinSmallMaxB) small_ok = false;
  if (small_ok) {
    const size_t smem = (size_t)2 * kFinSmallMaxB * ((((T - 1) * J + 3) & ~3) + 4) * sizeof(float);
    static size_t attr_fs[kMaxDevices] = {};
    if (smem > 48 * 1024 && smem_attr_needed(attr_fs, smem))
      KCCOT_CUDA(cudaFuncSetAttribute(cost_finalize_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    KCCOT_CUDA(launch_pdl(cost_finalize_small_kernel, dim3((unsigned)nprob, (unsigned)nblocks), dim3(256), smem, st, blocks,
                          T, J, s));
    KCCOT_LAUNCH_CHECK();
    return KCCOT_OK;
  }
  for (int p0 = 0; p0 < nprob; p0 += 65535) {     // grid.y limit
    CostBlocks bl = blocks;
    const int np = min(65535, nprob - p0);
    for (int i = 0; i < nblocks; ++i) {
      bl.b[i].part += (long long)p0 * bl.b[i].prob_stride;
      bl.b[i].C += (long long)p0 * bl.b[i].C_prob_stride;
      if (bl.b[i].h1) { bl.b[i].h1 += (long long)p0 * bl.b[i].Bx * T * J; bl.b[i].M1 += (long long)p0 * bl.b[i].By * T * J; }
      if (i == 0 && bl.zero) bl.zero += p0;
      if (bl.b[i].h2) { bl.b[i].h2 += (long long)p0 * bl.b[i].Bx * T * J; bl.b[i].M2 += (long long)p0 * bl.b[i].By * T * J; }
    }
    dim3 grid((unsigned)tiles, (unsigned)np, (unsigned)nblocks);
    if (all_tiled) {
      KCCOT_CUDA(launch_pdl(cost_finalize_tiled_kernel, grid, dim3(512), (size_t)0, st, bl, T, J, s));
    } else {
      KCCOT_CUDA(launch_pdl(cost_finalize_kernel, grid, dim3(128 * G), (size_t)0, st, bl, T, J, s));
    }
    KCCOT_LAUNCH_CHECK();
  }
  return KCCOT_OK;
}

// ---------------------------------------------------------------------------------------------
// g[p,r,k] (+)= 2s * sum_c W[p,r,c] * (a[p,r,k] - b[p,c,k]);  W[r,c] = Cbar[r*sr + c*sc]
// (gx: a=x, b=y, W=Cbar;  gy: a=y, b=x, W=Cbar^T)
// ---------------------------------------------------------------------------------------------
constexpr int GI = 32, GK = 128, GJ = 32;

__global__ void __launch_bounds__(256) cost_bwd_simt_kernel(const float* __restrict__ W, long long sr,
                                                            long long sc, long long wprob,
                                                            const float* __restrict__ a,
                                                            const float* __restrict__ b, int Ba, int Bb,
                                                            long long K, float two_s, float* __restrict__ g,
                                                            int accumulate) {
  __shared__ float ws[GI][GJ + 1];
  __shared__ float bs[GJ][GK];
  const int p = blockIdx.z;
  const float* Wp = W + (long long)p * wprob;
  const float* ap = a + (long long)p * Ba * K;
  const float* bp = b + (long long)p * Bb * K;
  float* gp = g + (long long)p * Ba * K;
  const long long kc = (long long)blockIdx.x * GK;
  const int r0 = blockIdx.y * GI;
  const int t = threadIdx.x, tk = t & 31, tr = t >> 5;   // 8 row groups x 4 rows, 32 lanes x 4 strided k
  float av[4][4], acc[4][4];
#pragma unroll
  for (int m = 0; m < 4; ++m)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int r = r0 + tr * 4 + m;
      const long long k = kc + tk + 32 * e;
      av[m][e] = (r < Ba && k < K) ? ap[(long long)r * K + k] : 0.f;
      acc[m][e] = 0.f;
    }
  for (int c0 = 0; c0 < Bb; c0 += GJ) {
    for (int q = t; q < GI * GJ; q += 256) {
      const int rr = q / GJ, cc = q % GJ;
      ws[rr][cc] = (r0 + rr < Ba && c0 + cc < Bb) ? Wp[(long long)(r0 + rr) * sr + (long long)(c0 + cc) * sc] : 0.f;
    }
    for (int q = t; q < GJ * GK; q += 256) {
      const int cc = q / GK, kk = q % GK;
      bs[cc][kk] = (c0 + cc < Bb && kc + kk < K) ? bp[(long long)(c0 + cc) * K + kc + kk] : 0.f;
    }
    __syncthreads();
#pragma unroll 4
    for (int c = 0; c < GJ; ++c) {
      float bv[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) bv[e] = bs[c][tk + 32 * e];
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const float w = ws[tr * 4 + m][c];
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[m][e] = fmaf(w, av[m][e] - bv[e], acc[m][e]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    const int r = r0 + tr * 4 + m;
    if (r >= Ba) continue;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const long long k = kc + tk + 32 * e;
      if (k >= K) continue;
      const float v = two_s * acc[m][e];
      float* dst = gp + (long long)r * K + k;
      *dst = accumulate ? (*dst + v) : v;
    }
  }
}

int launch_cost_bwd_simt(const float* W, long long sr, long long sc, long long wprob, const float* a,
                         const float* b, int nprob, int Ba, int Bb, long long K, float s, float* g,
                         int accumulate, cudaStream_t st) {
  dim3 grid((unsigned)((K + GK - 1) / GK), (Ba + GI - 1) / GI, nprob);
  cost_bwd_simt_kernel<<<grid, 256, 0, st>>>(W, sr, sc, wprob, a, b, Ba, Bb, K, 2.f * s, g, accumulate);
  KCCOT_LAUNCH_CHECK();
  return KCCOT_OK;
}

// ---------------------------------------------------------------------------------------------
// martingale adjoint: up to 4 output tensors in one launch, each the sum of up to two products
//   out[p,r,q] (+)= s * sum_k W1[r,k] F1[k,q] (+ s * sum_k W2[r,k] F2[k,q]),  q = (t, c)
//   W[r,k] = Cbar[r*ld + k] or (transposed) Cbar[k*ld + r]
//   F mode 0 (gradient of h, F from M): M[k,t+1,c] - M[k,t,c] for t < T-1, else 0
//   F mode 1 (gradient of M, F from h): h[k,t-1,c] [t>=1] - h[k,t,c] [t<=T-2]
// ---------------------------------------------------------------------------------------------
constexpr int MR = 4;          // output rows per CTA
constexpr int MKC = 64;        // contraction chunk staged in shared memory

// One CTA = one job, MR consecutive output rows, all T*J columns.  The difference operator commutes
// with the contraction: out[r, q] = s * (m1 * G[r, q + d1] - m2 * G[r, q + d2]) with the plain product
// G = W X over the RAW rows of X (M or h), so nothing but X and the MR weight rows is staged: both
// sources' blocks are fetched with one batch of independent 16-byte loads, every thread owns a few
// (row, column) entries of G, and the differences are taken once at the end through shared memory.
// History: v0 walked global memory with two dependent loads per FMA (~45 us for 80 KFLOP); v1 staged
// a materialised factor matrix F and spent 60 % of its 19 us building it (ncu source view, r42).
__global__ void __launch_bounds__(256) martingale_bwd_kernel(MartJobs jobs, int T, int J, float s) {
  extern __shared__ float4 sh4[];
  float* sh = reinterpret_cast<float*>(sh4);     // X[2][MKC*TJ] | W[2][MR*MKC];  G[MR*TJ] aliases X at the end
  const MartJob& jb = jobs.j[blockIdx.y];
  const int r0 = blockIdx.x * MR, p = blockIdx.z;
  if (r0 >= jb.nrows || jb.out == nullptr) return;
  const int TJ = T * J;
  const int nout = MR * TJ;
  float* Wbase = sh + 2 * MKC * TJ;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  float acc[8];
  int wof[8], qof[8];                              // per-output offsets: weight row, column
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    acc[u] = 0.f;
    const int o = threadIdx.x + u * 256;
    const int rr = (o < nout) ? o / TJ : 0;
    wof[u] = rr * MKC;
    qof[u] = o - rr * TJ;
  }
  for (int k0 = 0; k0 < jb.ncontr; k0 += MKC) {
    const int kc = min(MKC, jb.ncontr - k0);
    const int n = kc * TJ;
    __syncthreads();
    for (int src = 0; src < 2; ++src) {
      const float* C = src ? jb.C2 : jb.C1;
      if (C == nullptr) continue;
      C += (long long)p * jb.cprob;
      const float* X = (src ? jb.X2 : jb.X1) + ((long long)p * jb.ncontr + k0) * TJ;   // kc contiguous raw rows
      float* Xs = sh + src * MKC * TJ;
      float* Ws = Wbase + src * MR * MKC;
      if ((n & 3) == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0) {
        const float4* X4 = reinterpret_cast<const float4*>(X);
        float4* Xs4 = reinterpret_cast<float4*>(Xs);
        constexpr int kBatch = 8;
        for (int e0 = threadIdx.x; e0 < (n >> 2); e0 += 256 * kBatch) {
          float4 v[kBatch];
#pragma unroll
          for (int u = 0; u < kBatch; ++u) {
            const int e = e0 + u * 256;
            if (e < (n >> 2)) v[u] = X4[e];
          }
#pragma unroll
          for (int u = 0; u < kBatch; ++u) {
            const int e = e0 + u * 256;
            if (e < (n >> 2)) Xs4[e] = v[u];
          }
        }
      } else {
        for (int e = threadIdx.x; e < n; e += 256) Xs[e] = X[e];
      }
      for (int rr = warp; rr < MR; rr += nwarps) {
        const int r = r0 + rr;
        for (int k = lane; k < kc; k += 32) {
          float w = 0.f;
          if (r < jb.nrows) w = jb.transposed ? C[(long long)(k0 + k) * jb.ld + r] : C[(long long)r * jb.ld + k0 + k];
          Ws[rr * MKC + k] = w;
        }
      }
    }
    __syncthreads();
    for (int src = 0; src < 2; ++src) {
      if ((src ? jb.C2 : jb.C1) == nullptr) continue;
      const float* Xs = sh + src * MKC * TJ;
      const float* Ws = Wbase + src * MR * MKC;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (threadIdx.x + u * 256 < nout) {
          const float* wr = Ws + wof[u];
          const float* fc = Xs + qof[u];
          float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
          int k = 0;
          for (; k + 8 <= kc; k += 8) {                 // 10 independent shared-memory loads per chain step
            const float4 w0 = *reinterpret_cast<const float4*>(wr + k);
            const float4 w1 = *reinterpret_cast<const float4*>(wr + k + 4);
            const float f0 = fc[(k + 0) * TJ], f1 = fc[(k + 1) * TJ], f2 = fc[(k + 2) * TJ], f3 = fc[(k + 3) * TJ];
            const float f4 = fc[(k + 4) * TJ], f5 = fc[(k + 5) * TJ], f6 = fc[(k + 6) * TJ], f7 = fc[(k + 7) * TJ];
            a0 = fmaf(w0.x, f0, a0); a1 = fmaf(w0.y, f1, a1); a2 = fmaf(w0.z, f2, a2); a3 = fmaf(w0.w, f3, a3);
            a0 = fmaf(w1.x, f4, a0); a1 = fmaf(w1.y, f5, a1); a2 = fmaf(w1.z, f6, a2); a3 = fmaf(w1.w, f7, a3);
          }
          for (; k < kc; ++k) a0 = fmaf(wr[k], fc[k * TJ], a0);
          acc[u] += (a0 + a1) + (a2 + a3);
        }
      }
    }
  }
  // G -> shared memory, then the (masked) differences of mart_factor applied to G's columns
  __syncthreads();
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const int o = threadIdx.x + u * 256;
    if (o < nout) sh[o] = acc[u];
  }
  __syncthreads();
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const int o = threadIdx.x + u * 256;
    if (o < nout) {
      const int rr = wof[u] / MKC, q = qof[u];
      const int r = r0 + rr;
      if (r < jb.nrows) {
        const int t = q / J;
        const float* g = sh + rr * TJ;
        float v;
        if (jb.mode == 0) v = (t < T - 1) ? g[q + J] - g[q] : 0.f;
        else v = ((t >= 1) ? g[q - J] : 0.f) - ((t <= T - 2) ? g[q] : 0.f);
        v *= s;
        float* dst = jb.out + ((long long)p * jb.nrows + r) * TJ + q;
        *dst = jb.acc ? *dst + v : v;
      }
    }
  }
}

// Large batches (B >= 256): the same products as a tiled GEMM.  One CTA = one job, TR = 64 output rows, all
// T*J <= 256 columns; the contraction is staged in chunks of 32 (weights [64 x 32], raw rows of X [32 x T*J]);
// thread (ty, tx) owns rows 8 ty .. 8 ty + 7 and columns tx + 32 c.  martingale_bwd_kernel re-reads all of X for
// every 4 output rows (25 ms at B = 8192); this one reads it once per 64 rows.
constexpr int TR = 64, TKC = 32;
// NC = ceil(T*J / 32): column groups a thread really owns (T*J = 160 at config 5: 5 of 8; 80 at config 4: 3 of 8)
template <int NC>
__global__ void __launch_bounds__(256) martingale_bwd_tiled_kernel(MartJobs jobs, int T, int J, float s) {
  extern __shared__ float4 sh4[];
  float* sh = reinterpret_cast<float*>(sh4);
  const MartJob& jb = jobs.j[blockIdx.y];
  const int r0 = blockIdx.x * TR, p = blockIdx.z;
  if (r0 >= jb.nrows || jb.out == nullptr) return;
  const int TJ = T * J;
  float* Xs = sh;                          // [TKC][TJ]
  float* Ws = sh + TKC * TJ;               // [TR][TKC + 1]
  const int t = threadIdx.x, tx = t & 31, ty = t >> 5;
  float acc[8][NC];
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int c = 0; c < NC; ++c) acc[a][c] = 0.f;
  for (int src = 0; src < 2; ++src) {
    const float* C = src ? jb.C2 : jb.C1;
    if (C == nullptr) continue;
    C += (long long)p * jb.cprob;
    const float* X = (src ? jb.X2 : jb.X1) + (long long)p * jb.ncontr * TJ;
    // software pipeline: the global loads of chunk k0 + TKC are in flight (registers) while chunk k0 is multiplied
    constexpr int XPT = TKC * 32 * NC / 256, WPT = TR * TKC / 256;          // X values (T*J <= 32 NC) and weights per thread
    float xr[XPT], wr[WPT];
    auto fetch = [&](int k0) {
      const int kc = min(TKC, jb.ncontr - k0);
#pragma unroll
      for (int i = 0; i < XPT; ++i) {
        const int e = t + 256 * i;
        xr[i] = (e < kc * TJ) ? X[(long long)k0 * TJ + e] : 0.f;
      }
#pragma unroll
      for (int i = 0; i < WPT; ++i) {
        const int e = t + 256 * i;
        if (!jb.transposed) {
          const int rr = e / TKC, k = e % TKC;
          wr[i] = (r0 + rr < jb.nrows && k < kc) ? C[(long long)(r0 + rr) * jb.ld + k0 + k] : 0.f;
        } else {
          const int k = e / TR, rr = e % TR;
          wr[i] = (r0 + rr < jb.nrows && k < kc) ? C[(long long)(k0 + k) * jb.ld + r0 + rr] : 0.f;
        }
      }
    };
    fetch(0);
    for (int k0 = 0; k0 < jb.ncontr; k0 += TKC) {
      __syncthreads();
#pragma unroll
      for (int i = 0; i < XPT; ++i) {
        const int e = t + 256 * i;
        if (e < TKC * TJ) Xs[e] = xr[i];
      }
#pragma unroll
      for (int i = 0; i < WPT; ++i) {
        const int e = t + 256 * i;
        if (!jb.transposed) Ws[(e / TKC) * (TKC + 1) + e % TKC] = wr[i];
        else Ws[(e % TR) * (TKC + 1) + e / TR] = wr[i];
      }
      __syncthreads();
      if (k0 + TKC < jb.ncontr) fetch(k0 + TKC);
#pragma unroll 4
      for (int k = 0; k < TKC; ++k) {
        float w[8], x[NC];
#pragma unroll
        for (int a = 0; a < 8; ++a) w[a] = Ws[(ty * 8 + a) * (TKC + 1) + k];
#pragma unroll
        for (int c = 0; c < NC; ++c) x[c] = (tx + 32 * c < TJ) ? Xs[k * TJ + tx + 32 * c] : 0.f;
#pragma unroll
        for (int a = 0; a < 8; ++a)
#pragma unroll
          for (int c = 0; c < NC; ++c) acc[a][c] = fmaf(w[a], x[c], acc[a][c]);
      }
    }
  }
  // G -> shared memory [TR][TJ], then the (masked) time differences
  __syncthreads();
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int c = 0; c < NC; ++c)
      if (tx + 32 * c < TJ) sh[(ty * 8 + a) * TJ + tx + 32 * c] = acc[a][c];
  __syncthreads();
  for (int e = t; e < TR * TJ; e += 256) {
    const int rr = e / TJ, q = e - rr * TJ;
    const int r = r0 + rr;
    if (r >= jb.nrows) continue;
    const int tt = q / J;
    const float* g = sh + rr * TJ;
    float v;
    if (jb.mode == 0) v = (tt < T - 1) ? g[q + J] - g[q] : 0.f;
    else v = ((tt >= 1) ? g[q - J] : 0.f) - ((tt <= T - 2) ? g[q] : 0.f);
    v *= s;
    float* dst = jb.out + ((long long)p * jb.nrows + r) * TJ + q;
    *dst = jb.acc ? *dst + v : v;
  }
}

int launch_martingale_jobs(const MartJobs& jobs, int njobs, int nprob, int T, int J, float s, cudaStream_t st) {
  int maxrows = 0;
  for (int i = 0; i < njobs; ++i)
    if (jobs.j[i].out) maxrows = max(maxrows, jobs.j[i].nrows);
  if (maxrows == 0) return KCCOT_OK;
  const int TJ = T * J;
  // many problems per call (BASELINE config 4): the 4-rows-per-CTA kernel below would launch 16 x 4 x nprob CTAs that
  // each stage all of X (321 us at 256 problems); one 64-row CTA per (job, problem) stages it once
  if ((maxrows >= 256 || nprob >= 16) && TJ <= 256) {
    const size_t a = (size_t)(TKC * TJ + TR * (TKC + 1)) * sizeof(float), b = (size_t)TR * TJ * sizeof(float);
    const size_t smem_t = a > b ? a : b;
    const dim3 grid((maxrows + TR - 1) / TR, njobs, nprob);
    auto go = [&](auto nc_tag) -> int {
      constexpr int NC = decltype(nc_tag)::value;
      static size_t attr_t[kMaxDevices] = {};
      if (smem_t > 48 * 1024 && smem_attr_needed(attr_t, smem_t))
        KCCOT_CUDA(cudaFuncSetAttribute(martingale_bwd_tiled_kernel<NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_t));
      martingale_bwd_tiled_kernel<NC><<<grid, 256, smem_t, st>>>(jobs, T, J, s);
      KCCOT_LAUNCH_CHECK();
      return KCCOT_OK;
    };
    const int nc = (TJ + 31) / 32;
    if (nc <= 3) { if (int rc = go(std::integral_constant<int, 3>{})) return rc; }
    else if (nc <= 5) { if (int rc = go(std::integral_constant<int, 5>{})) return rc; }
    else { if (int rc = go(std::integral_constant<int, 8>{})) return rc; }
    return KCCOT_OK;
  }
  const size_t smem = (size_t)(2 * MKC * TJ + 2 * MR * MKC) * sizeof(float);
  if (MR * TJ > 8 * 256 || smem > 200 * 1024) {
    set_error("martingale adjoint: T*J = %d too large (max %d)", TJ, 8 * 256 / MR);
    return KCCOT_EINVAL;
  }
  static size_t attr[kMaxDevices] = {};
  if (smem > 48 * 1024 && smem_attr_needed(attr, 200 * 1024))
    KCCOT_CUDA(cudaFuncSetAttribute(martingale_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  dim3 grid((maxrows + MR - 1) / MR, njobs, nprob);
  martingale_bwd_kernel<<<grid, 256, smem, st>>>(jobs, T, J, s);
  KCCOT_LAUNCH_CHECK();
  return KCCOT_OK;
}

int launch_martingale_bwd(const float* Cbar, long long cprob, const float* h, const float* M, int nprob,
                          int Bx, int By, int T, int J, float s, float w, float* gh, float* gM,
                          int acc_h, int acc_M, cudaStream_t st) {
  (void)w;
  MartJobs jobs{};
  jobs.j[0] = MartJob{gh, Cbar, M, nullptr, nullptr, cprob, By, Bx, By, 0, 0, acc_h};
  jobs.j[1] = MartJob{gM, Cbar, h, nullptr, nullptr, cprob, By, By, Bx, 1, 1, acc_M};
  return launch_martingale_jobs(jobs, 2, nprob, T, J, s, st);
}

}  // namespace kccot
