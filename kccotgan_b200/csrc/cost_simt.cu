// CUDA-core (fp32) kernels of the cost path: pairwise squared distances in the reference's direct
// (x - y)^2 form (gan_utils.py:14-17), the cost finalisation that adds the martingale terms
// (gan_utils.py:34-43, 59-72), and the adjoints.  Any shape, any alignment.  The tensor-core
// kernels in gram_tcgen05.cu replace the distance part when the shape allows; these remain the
// generic-shape path and the on-device cross-check.
#include "common.cuh"
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
// ---------------------------------------------------------------------------------------------
constexpr int FL = 8;   // lanes cooperating on one output element

__global__ void __launch_bounds__(256) cost_finalize_kernel(CostBlocks blocks, int nprob, int T, int J,
                                                            float s) {
  const CostBlock& b = blocks.b[blockIdx.z];
  const long long n = (long long)nprob * b.Bx * b.By;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long idx = gid / FL;
  const int sub = (int)(gid % FL);
  const bool live = idx < n;
  const int j = live ? (int)(idx % b.By) : 0;
  const int i = live ? (int)((idx / b.By) % b.Bx) : 0;
  const int p = live ? (int)(idx / ((long long)b.Bx * b.By)) : 0;
  // the split-K partials are summed in fp64: a sequential fp32 sum of ~150-300 partials of a value
  // near 1e4 would by itself cost ~5e-4 absolute on C, more than the whole fp32 budget of the path
  double d = 0.0;
  if (live && !(b.zero_diag && i == j)) {
    const float* pp = b.part + (long long)p * b.prob_stride + (long long)(b.row_off + i) * b.ld + b.col_off + j;
    for (int ks = sub; ks < b.nks; ks += FL) d += (double)pp[(long long)ks * b.ks_stride];
  }
  const int tj = (T - 1) * J;
  float a = 0.f;
  if (live) {
    for (int pair = 0; pair < 2; ++pair) {
      const float* h = pair ? b.h2 : b.h1;
      const float* M = pair ? b.M2 : b.M1;
      if (h == nullptr) continue;
      const float* hr = h + ((long long)p * b.Bx + i) * T * J;
      const float* Mr = M + ((long long)p * b.By + j) * T * J;
      for (int q = sub; q < tj; q += FL) a = fmaf(hr[q], Mr[q + J] - Mr[q], a);
    }
  }
  d += (double)a;
#pragma unroll
  for (int o = FL / 2; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
  if (live && sub == 0) b.C[(long long)p * b.C_prob_stride + (long long)i * b.By + j] = (float)((double)s * d);
}

int launch_cost_finalize(const CostBlocks& blocks, int nblocks, int nprob, int T, int J, float s,
                         cudaStream_t st) {
  long long nmax = 0;
  for (int i = 0; i < nblocks; ++i) nmax = max(nmax, (long long)nprob * blocks.b[i].Bx * blocks.b[i].By);
  dim3 grid((unsigned)((nmax * FL + 255) / 256), 1, nblocks);
  cost_finalize_kernel<<<grid, 256, 0, st>>>(blocks, nprob, T, J, s);
  KCCOT_LAUNCH_CHECK();
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
__device__ __forceinline__ float mart_factor(const float* X, int mode, int t, int T, int J, int q) {
  if (mode == 0) return (t < T - 1) ? X[q + J] - X[q] : 0.f;
  return ((t >= 1) ? X[q - J] : 0.f) - ((t <= T - 2) ? X[q] : 0.f);
}

__global__ void __launch_bounds__(128) martingale_bwd_kernel(MartJobs jobs, int T, int J, float s) {
  extern __shared__ float wrow[];      // [2][ncontr]
  const MartJob& jb = jobs.j[blockIdx.y];
  const int r = blockIdx.x, p = blockIdx.z;
  if (r >= jb.nrows || jb.out == nullptr) return;
  const int TJ = T * J;
  for (int src = 0; src < 2; ++src) {
    const float* C = src ? jb.C2 : jb.C1;
    if (C == nullptr) continue;
    C += (long long)p * jb.cprob;
    for (int k = threadIdx.x; k < jb.ncontr; k += blockDim.x)
      wrow[src * jb.ncontr + k] = jb.transposed ? C[(long long)k * jb.ld + r] : C[(long long)r * jb.ld + k];
  }
  __syncthreads();
  float* out = jb.out + ((long long)p * jb.nrows + r) * TJ;
  for (int q = threadIdx.x; q < TJ; q += blockDim.x) {
    const int t = q / J;
    float a = 0.f;
    for (int src = 0; src < 2; ++src) {
      const float* X = src ? jb.X2 : jb.X1;
      if ((src ? jb.C2 : jb.C1) == nullptr) continue;
      X += (long long)p * jb.ncontr * TJ;
      const float* w = wrow + src * jb.ncontr;
      for (int k = 0; k < jb.ncontr; ++k) a = fmaf(w[k], mart_factor(X + (long long)k * TJ, jb.mode, t, T, J, q), a);
    }
    const float v = s * a;
    out[q] = jb.acc ? out[q] + v : v;
  }
}

int launch_martingale_jobs(const MartJobs& jobs, int njobs, int nprob, int T, int J, float s, cudaStream_t st) {
  int maxrows = 0, maxc = 0;
  for (int i = 0; i < njobs; ++i) { maxrows = max(maxrows, jobs.j[i].nrows); maxc = max(maxc, jobs.j[i].ncontr); }
  if (maxrows == 0) return KCCOT_OK;
  dim3 grid(maxrows, njobs, nprob);
  martingale_bwd_kernel<<<grid, 128, (size_t)2 * maxc * sizeof(float), st>>>(jobs, T, J, s);
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
