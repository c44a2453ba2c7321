// Persistent Sinkhorn for 64 < B <= 8192 (and for one rank's row block of a row-sharded problem):
// ALL iterations of all problems of a call run inside ONE cooperative kernel, gan_utils.py:151-164.
//
// Scaling form with the kernel absorbed at the row minima (same algebra as sinkhorn_small.cu):
//     Kt_ij = exp2(alpha_i - Chat_ij),  alpha_i = min_j Chat_ij,  Chat = (C - c0) log2(e) / eps
//     s_i = sum_j Kt_ij b_j,   uhat_i = ahat + alpha_i - log2 s_i,   a_i = 2^ahat / s_i
//     t_j = sum_i Kt_ij a_i,   vhat_j = ahat - log2 t_j,             b_j = 2^ahat / t_j
// One iteration = ONE pass over Kt: a CTA owns a block of rows; for a group of RG rows it loads the rows into
// registers (thread = fixed set of columns, 16-byte loads), reduces the RG row sums over the block, forms a_i
// and adds a_i Kt_ij into its per-thread column accumulators before the rows leave the registers.  Column
// partials of all CTAs are combined after a grid barrier (column-sliced over the CTAs, fixed order:
// deterministic), a second barrier publishes b.  No exponential in the loop, 4 bytes per matrix element per
// iteration; the rows of a CTA stay in SHARED MEMORY for the whole solve when they fit (B <= ~1500 on one GPU,
// 1024 x 8192 shards on eight), else they stream from L2 / HBM.
//
// Row shards (nranks > 1): every rank runs the same kernel on its rows; the [B] column sums are exchanged
// INSIDE the kernel: each rank stores its partial vector into every peer's mailbox through NVLink-mapped
// pointers, publishes an epoch flag with system-scope release, spins on its own flags, and sums the mailbox
// slots in rank order (identical result on all ranks).  No host-launched collective in the loop.
//
// Guard: if a scaling leaves [2^-90, 2^90] the problem switches, at the next barrier, to the reference's
// log-domain updates with max subtraction (from the last good history row), inside the same kernel.
//
// Backward (reverse mode through the executed iterations): every softmax matrix of the
// recurrences is Kt times a rank-one factor, so a reverse step is again one pass of two mat-vecs over Kt, and
// the adjoint of C is NOT accumulated step by step: the 2n rank-one factors are stored (2 x [B, 2n]) and
//     Cbar = g pi (1 - C/eps) + Kt .* (RowF ColF^T)
// is formed once at the end by a tiled kernel (B^2 x 2n FMAs).  If a factor would leave 2^+-60 the problem
// falls back to direct exponentials with a read-modify-write of Cbar per step (same kernel).
#include <cooperative_groups.h>

#include "common.cuh"
#include "sinkhorn.cuh"

namespace cg = cooperative_groups;

namespace kccot {

namespace {
constexpr int kNT = 512;
constexpr int kNW = kNT / 32;
constexpr float kLo = 8.0e-28f;    // ~2^-90
constexpr float kHi = 1.2e27f;     // ~2^90
constexpr float kExpLim = 60.f;
constexpr float kNegBig = -3.0e38f;
constexpr int kRedFloats = 2 * kNW * 4 * 2;      // two parities x warps x (up to 4 rows) x 2 values

struct PState {           // per problem, in global memory
  int trip;               // forward: scaling left the safe range -> log mode; backward: direct mode
  int done;
  int nits;
  int error;              // exchange timeout
  float err;              // sum |u - u_prev| of the current iteration (log2 units)
  float s0, s1;
  int pad;
};

__device__ __forceinline__ float ldcg_f(const float* p) { return __ldcg(p); }
__device__ __forceinline__ float4 ldcg_f4(const float4* p) { return __ldcg(p); }
__device__ __forceinline__ int ld_volatile_i(const int* p) { return *reinterpret_cast<const volatile int*>(p); }

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// sum (or max) of RG per-thread values over the block; every thread gets the result; one barrier
template <int RG, bool MAX>
__device__ __forceinline__ void block_reduce_rg(float (&v)[RG], float* red, int& parity) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int r = 0; r < RG; ++r) v[r] = MAX ? warp_max(v[r]) : warp_sum(v[r]);
  float* buf = red + parity * (kNW * 4);
  if (lane == 0) {
#pragma unroll
    for (int r = 0; r < RG; ++r) buf[warp * 4 + r] = v[r];
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < RG; ++r) {
    float a = MAX ? kNegBig : 0.f;
#pragma unroll
    for (int w = 0; w < kNW; ++w) a = MAX ? fmaxf(a, buf[w * 4 + r]) : a + buf[w * 4 + r];
    v[r] = a;
  }
  parity ^= 1;
}

// ---------------------------------------------------------------------------------------------
// shared geometry of the persistent kernels
// ---------------------------------------------------------------------------------------------
struct Geo {
  int p, cl, Gp;            // problem of this CTA, index within the problem's CTAs, CTAs of the problem
  int r_begin, r_end;       // local rows [r_begin, r_end) of this CTA
  int nv4;                  // float4 granules of the exchanged vector: B / 4 + 1 (last granule: scalar slots)
  int g_begin, g_end;       // granules this CTA combines
};
__device__ __forceinline__ Geo make_geo(int np, int Brows, int B) {
  Geo g;
  g.p = blockIdx.x % np;
  g.cl = blockIdx.x / np;
  g.Gp = ((int)gridDim.x - g.p + np - 1) / np;
  const int rpc = (Brows + g.Gp - 1) / g.Gp;
  g.r_begin = min(Brows, g.cl * rpc);
  g.r_end = min(Brows, g.r_begin + rpc);
  g.nv4 = B / 4 + 1;
  const int slice = (g.nv4 + g.Gp - 1) / g.Gp;
  g.g_begin = min(g.nv4, g.cl * slice);
  g.g_end = min(g.nv4, g.g_begin + slice);
  return g;
}

// Cross-rank sum (or (max, sum-exp) combine) of the column vectors.  `nvec` = 1: plain sums; 2: pairs.
// Called by ALL CTAs of the grid the same number of times (it contains grid barriers).
struct Xchg {
  int nranks, rank;
  float* mbox[kMaxShardRanks];              // every rank's mailbox [2][np][nranks][2][nv4 * 4] (peer-mapped)
  unsigned long long* flags[kMaxShardRanks];   // every rank's epoch flags [nranks][kFlagStride]; epochs grow monotonically over
  unsigned long long epoch0;                   // launches (the caller passes a fresh base per launch), never reset
};

__device__ __forceinline__ size_t mbox_off(int parity, int np, int p, int nranks, int r, int vec, int nv4) {
  return ((((size_t)parity * np + p) * nranks + r) * 2 + vec) * ((size_t)nv4 * 4);
}

// phase A: push this rank's combined local granule to every rank's mailbox
__device__ __forceinline__ void xchg_push(const Xchg& X, unsigned long long epoch, int np, int p, int nv4, int gi, int vec, float4 v) {
  for (int r = 0; r < X.nranks; ++r) {
    float4* dst = reinterpret_cast<float4*>(X.mbox[r] + mbox_off((int)(epoch & 1), np, p, X.nranks, X.rank, vec, nv4)) + gi;
    *dst = v;
  }
}
// between phase A and B: a handshake PER CTA.  Every rank runs the same grid with the same column slices, CTA c pushes
// slice c to every peer and afterwards reads only slice c of what the peers pushed, so CTA c needs nothing but "CTA c
// of every peer has pushed": it publishes flag (rank, c) on every peer and spins on its own flags (r, c).  (The first
// version funnelled the handshake through CTA 0 between two grid barriers: four grid barriers per iteration instead
// of the two a single rank needs.)
constexpr int kFlagStride = KCCOT_SHARD_FLAGS_PER_RANK;     // flags per source rank >= CTAs of the cooperative grid
__device__ __forceinline__ void xchg_sync(const Xchg& X, unsigned long long epoch, PState* st) {
  __threadfence_system();                  // this thread's pushes are visible system-wide ...
  __syncthreads();                         // ... and so are those of the whole CTA, before the flag goes out
  if ((int)threadIdx.x < X.nranks && (int)threadIdx.x != X.rank) {
    const int r = threadIdx.x;
    st_release_sys(X.flags[r] + (size_t)X.rank * kFlagStride + blockIdx.x, epoch);
    const unsigned long long t0 = globaltimer_ns();
    while (ld_acquire_sys(X.flags[X.rank] + (size_t)r * kFlagStride + blockIdx.x) < epoch) {
      if (globaltimer_ns() - t0 > 4000000000ull) {       // 4 s: a peer died; give up instead of hanging the GPU
        st->error = 1;
        break;
      }
    }
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// absorb: alpha_i = min_j Chat_ij ; Kt_ij = exp2(alpha_i - Chat_ij).  One CTA per row.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) absorb_rows_kernel(const float* __restrict__ C, int Brows, int B, long long cstride,
                                                          float kscale, const float* __restrict__ shift, int shift_stride,
                                                          float* __restrict__ Kt, long long kstride,
                                                          float* __restrict__ alpha, int astride) {
  __shared__ float red[8];
  const int p = blockIdx.y, i = blockIdx.x;
  const float* row = C + (long long)p * cstride + (long long)i * B;
  const float c0 = shift[(long long)p * shift_stride];
  float mn = 3.0e38f;
  for (int j = threadIdx.x; j < B; j += 256) mn = fminf(mn, row[j]);
  mn = warp_min(mn);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mn;
  __syncthreads();
  mn = red[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) mn = fminf(mn, red[w]);
  const float al = (mn - c0) * kscale;
  if (threadIdx.x == 0) alpha[(long long)p * astride + i] = al;
  float* out = Kt + (long long)p * kstride + (long long)i * B;
  for (int j = threadIdx.x; j < B; j += 256) out[j] = fast_exp2(al - (row[j] - c0) * kscale);
}

__device__ __forceinline__ void atomic_min_float(float* addr, float v) {
  if (v >= 0.f) atomicMin(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMax(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}
__global__ void persist_init_kernel(PState* st, float* shift, int shift_stride, int np) {
  const int p = threadIdx.x;
  if (p < np) {
    st[p] = PState{0, 0, 0, 0, 0.f, 0.f, 0.f, 0};
    if (shift) shift[(long long)p * shift_stride] = __int_as_float(0x7f800000);
  }
}
__global__ void __launch_bounds__(256) persist_min_kernel(const float* __restrict__ C, long long n, long long cstride,
                                                          float* shift, int shift_stride) {
  const int p = blockIdx.y;
  const float* Cp = C + (long long)p * cstride;
  float m = 3.0e38f;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) m = fminf(m, Cp[i]);
  m = warp_min(m);
  if ((threadIdx.x & 31) == 0) atomic_min_float(shift + (long long)p * shift_stride, m);
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
struct PFwd {
  const float* C;        // [np][Brows][B]
  const float* Kt;       // [np][Brows][B]
  const float* alpha;    // [np][Brows]
  const float* shift;    // [np] (stride shift_stride floats)
  float *u_hist, *v_hist;   // [np][L+1][B]; row 0 zeroed by the host
  int32_t* nits;         // [np]
  float* cost;           // [np]  (single rank) or partial (s0, s1) pairs [np][2] when nranks > 1
  float *part_a, *part_b;   // [np][Gmax][nv4 * 4]
  float* bvec;           // [np][B]
  PState* state;         // [np]
  int np, Brows, B, row0, L, Lmin, exit_on_index, resident, shift_stride, Gmax;
  int partial_cost;      // write the (s0, s1) sums of this rank's rows instead of the finished cost
  float kscale, ahat, thresh;
  Xchg X;
};

template <int CPT, int RG>
__global__ void __launch_bounds__(kNT, 1) sk_persist_fwd_kernel(const __grid_constant__ PFwd P) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ float4 smem4[];
  float* red = reinterpret_cast<float*>(smem4);
  float* srows = red + kRedFloats;
  const int tid = threadIdx.x;
  const int B = P.B, Brows = P.Brows;
  const Geo G = make_geo(P.np, Brows, B);
  const int p = G.p;
  const long long hs = (long long)(P.L + 1) * B;
  const float* Cp = P.C + (long long)p * Brows * B;
  const float* Ktp = P.Kt + (long long)p * Brows * B;
  const float* alp = P.alpha + (long long)p * Brows;
  float* uh = P.u_hist + p * hs;
  float* vh = P.v_hist + p * hs;
  float* pa = P.part_a + (long long)p * P.Gmax * G.nv4 * 4;
  float* pb = P.part_b + (long long)p * P.Gmax * G.nv4 * 4;
  float* bv = P.bvec + (long long)p * B;
  PState* st = P.state + p;
  const float c0 = P.shift[(long long)p * P.shift_stride];
  const float kscale = P.kscale, ahat = P.ahat, two_ahat = exp2f(P.ahat);
  constexpr int NG = CPT / 4;                       // float4 granules per thread
  bool colok[NG];
#pragma unroll
  for (int m = 0; m < NG; ++m) colok[m] = 4 * (tid + kNT * m) < B;

  if (P.resident) {
    for (int r = G.r_begin; r < G.r_end; ++r) {
      const float4* src = reinterpret_cast<const float4*>(Ktp + (long long)r * B);
      float4* dst = reinterpret_cast<float4*>(srows + (long long)(r - G.r_begin) * B);
      for (int q = tid; q < B / 4; q += kNT) dst[q] = src[q];
    }
  }
  __syncthreads();

  float b[CPT], tcol[CPT], cmx[CPT];
#pragma unroll
  for (int c = 0; c < CPT; ++c) { b[c] = 1.f; tcol[c] = 0.f; cmx[c] = kNegBig; }
  int it = 0, mode = 0, parity = 0;
  bool done = false;
  unsigned long long epoch = P.X.epoch0;

  for (;;) {
    // ------------------------------------------------------------------ row pass
    if (!done) {
      float errp = 0.f;
#pragma unroll
      for (int c = 0; c < CPT; ++c) { tcol[c] = 0.f; cmx[c] = kNegBig; }
      for (int g0 = G.r_begin; g0 < G.r_end; g0 += RG) {
        float kr[RG][CPT];
#pragma unroll
        for (int r = 0; r < RG; ++r) {
          const int row = g0 + r;
          const bool rok = row < G.r_end;
          const float* base = (mode == 0) ? (P.resident ? srows + (long long)(row - G.r_begin) * B : Ktp + (long long)row * B)
                                          : Cp + (long long)row * B;
          const float4* src = reinterpret_cast<const float4*>(base);
#pragma unroll
          for (int m = 0; m < NG; ++m) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (rok && colok[m]) v = src[tid + kNT * m];
            kr[r][4 * m + 0] = v.x; kr[r][4 * m + 1] = v.y; kr[r][4 * m + 2] = v.z; kr[r][4 * m + 3] = v.w;
          }
        }
        if (mode == 0) {
          float part[RG];
#pragma unroll
          for (int r = 0; r < RG; ++r) {
            float a0 = 0.f, a1 = 0.f;
#pragma unroll
            for (int c = 0; c < CPT; c += 2) { a0 = fmaf(kr[r][c], b[c], a0); a1 = fmaf(kr[r][c + 1], b[c + 1], a1); }
            part[r] = a0 + a1;
          }
          block_reduce_rg<RG, false>(part, red, parity);
          float a[RG];
#pragma unroll
          for (int r = 0; r < RG; ++r) {
            const int row = g0 + r;
            const float s = part[r];
            a[r] = (row < G.r_end) ? two_ahat / s : 0.f;
            if (tid == r && row < G.r_end) {
              const float unew = ahat + alp[row] - log2f(s);
              const long long gi = (long long)P.row0 + row;
              errp += fabsf(unew - uh[(long long)it * B + gi]);
              uh[(long long)(it + 1) * B + gi] = unew;
              if (!(s > kLo && s < kHi)) st->trip = 1;
            }
          }
#pragma unroll
          for (int r = 0; r < RG; ++r)
#pragma unroll
            for (int c = 0; c < CPT; ++c) tcol[c] = fmaf(kr[r][c], a[r], tcol[c]);
        } else {
          // log-domain: b[] holds vhat_j, kr <- x = vhat_j - Chat_ij
          float mx[RG];
#pragma unroll
          for (int r = 0; r < RG; ++r) {
            float m0 = kNegBig;
#pragma unroll
            for (int c = 0; c < CPT; ++c) {
              const bool ok = colok[c >> 2] && (g0 + r < G.r_end);
              kr[r][c] = ok ? b[c] - (kr[r][c] - c0) * kscale : kNegBig;
              m0 = fmaxf(m0, kr[r][c]);
            }
            mx[r] = m0;
          }
          block_reduce_rg<RG, true>(mx, red, parity);
          float sm[RG];
#pragma unroll
          for (int r = 0; r < RG; ++r) {
            float a0 = 0.f;
#pragma unroll
            for (int c = 0; c < CPT; ++c) a0 += (kr[r][c] > kNegBig) ? fast_exp2(kr[r][c] - mx[r]) : 0.f;
            sm[r] = a0;
          }
          block_reduce_rg<RG, false>(sm, red, parity);
#pragma unroll
          for (int r = 0; r < RG; ++r) {
            const int row = g0 + r;
            if (row >= G.r_end) continue;
            const float unew = ahat - (mx[r] + log2f(sm[r]));
            if (tid == r) {
              const long long gi = (long long)P.row0 + row;
              errp += fabsf(unew - uh[(long long)it * B + gi]);
              uh[(long long)(it + 1) * B + gi] = unew;
            }
            // column side: y = unew - Chat_ij = unew - (vhat_j - x)
#pragma unroll
            for (int c = 0; c < CPT; ++c) {
              if (kr[r][c] > kNegBig) {
                const float y = unew - (b[c] - kr[r][c]);
                if (y > cmx[c]) { tcol[c] = tcol[c] * fast_exp2(cmx[c] - y) + 1.f; cmx[c] = y; }
                else tcol[c] += fast_exp2(y - cmx[c]);
              }
            }
          }
        }
      }
      // column partials of this CTA (+ the err partial in the scalar granule)
      float4* pa4 = reinterpret_cast<float4*>(pa + (long long)G.cl * G.nv4 * 4);
      float4* pb4 = reinterpret_cast<float4*>(pb + (long long)G.cl * G.nv4 * 4);
#pragma unroll
      for (int m = 0; m < NG; ++m) {
        if (colok[m]) {
          if (mode == 0) {
            pa4[tid + kNT * m] = make_float4(tcol[4 * m], tcol[4 * m + 1], tcol[4 * m + 2], tcol[4 * m + 3]);
          } else {
            pa4[tid + kNT * m] = make_float4(cmx[4 * m], cmx[4 * m + 1], cmx[4 * m + 2], cmx[4 * m + 3]);
            pb4[tid + kNT * m] = make_float4(tcol[4 * m], tcol[4 * m + 1], tcol[4 * m + 2], tcol[4 * m + 3]);
          }
        }
      }
      // err partial: threads 0..RG-1 hold pieces
      float e4[1] = {errp};
      block_reduce_rg<1, false>(e4, red, parity);
      if (tid == 0) {
        pa4[G.nv4 - 1] = make_float4(e4[0], 0.f, 0.f, 0.f);
        if (mode != 0) pb4[G.nv4 - 1] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    grid.sync();
    // ------------------------------------------------------------------ exit test (flags written before the barrier)
    {
      int alld = 1;
      for (int q = 0; q < P.np; ++q) alld &= ld_volatile_i(&P.state[q].done);
      if (alld) break;
    }
    // ------------------------------------------------------------------ column combine
    const bool multi = P.X.nranks > 1;
    ++epoch;
    const bool tripped_u = !done && mode == 0 && ld_volatile_i(&st->trip) != 0;   // uniform over the problem's CTAs
    auto finalize = [&](int gi, float4 A, float4 Bv) {
      if (gi == G.nv4 - 1) {                         // scalar granule: total err, number of tripped ranks
        st->err = A.x;
        if (A.y > 0.f) st->trip = 1;
        return;
      }
      float4 vv, bb;
      float* av = reinterpret_cast<float*>(&A);
      float* sv = reinterpret_cast<float*>(&Bv);
      float* vo = reinterpret_cast<float*>(&vv);
      float* bo = reinterpret_cast<float*>(&bb);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (mode == 0) {
          const float t = av[e];
          vo[e] = ahat - log2f(t);
          bo[e] = two_ahat / t;
          if (!(t > kLo && t < kHi)) st->trip = 1;
        } else {
          vo[e] = ahat - (av[e] + log2f(sv[e]));
          bo[e] = vo[e];
        }
      }
      reinterpret_cast<float4*>(vh + (long long)(it + 1) * B)[gi] = vv;
      reinterpret_cast<float4*>(bv)[gi] = bb;
    };
    if (!done) {
      for (int gi = G.g_begin + tid; gi < G.g_end; gi += kNT) {
        float4 A, Bv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gi == G.nv4 - 1) {                       // (err, tripped) travel with the vector: every rank must switch mode together
          A = make_float4(0.f, tripped_u ? 1.f : 0.f, 0.f, 0.f);
          for (int c = 0; c < G.Gp; ++c) A.x += ldcg_f(pa + (long long)c * G.nv4 * 4 + (long long)gi * 4);
        } else if (tripped_u) {
          continue;
        } else if (mode == 0) {
          A = make_float4(0.f, 0.f, 0.f, 0.f);
          for (int c = 0; c < G.Gp; ++c) {
            const float4 v = ldcg_f4(reinterpret_cast<const float4*>(pa + (long long)c * G.nv4 * 4) + gi);
            A.x += v.x; A.y += v.y; A.z += v.z; A.w += v.w;
          }
        } else {
          A = make_float4(kNegBig, kNegBig, kNegBig, kNegBig);
          for (int c = 0; c < G.Gp; ++c) {
            const float4 v = ldcg_f4(reinterpret_cast<const float4*>(pa + (long long)c * G.nv4 * 4) + gi);
            A.x = fmaxf(A.x, v.x); A.y = fmaxf(A.y, v.y); A.z = fmaxf(A.z, v.z); A.w = fmaxf(A.w, v.w);
          }
          for (int c = 0; c < G.Gp; ++c) {
            const float4 m = ldcg_f4(reinterpret_cast<const float4*>(pa + (long long)c * G.nv4 * 4) + gi);
            const float4 s = ldcg_f4(reinterpret_cast<const float4*>(pb + (long long)c * G.nv4 * 4) + gi);
            Bv.x += s.x * fast_exp2(m.x - A.x); Bv.y += s.y * fast_exp2(m.y - A.y);
            Bv.z += s.z * fast_exp2(m.z - A.z); Bv.w += s.w * fast_exp2(m.w - A.w);
          }
        }
        if (multi) {
          xchg_push(P.X, epoch, P.np, p, G.nv4, gi, 0, A);
          if (mode != 0) xchg_push(P.X, epoch, P.np, p, G.nv4, gi, 1, Bv);
        } else {
          finalize(gi, A, Bv);
        }
      }
    }
    if (multi) {
      xchg_sync(P.X, epoch, st);
      if (!done) {
        const float* mb = P.X.mbox[P.X.rank];
        for (int gi = G.g_begin + tid; gi < G.g_end; gi += kNT) {
          if (tripped_u && gi != G.nv4 - 1) continue;
          float4 A, Bv = make_float4(0.f, 0.f, 0.f, 0.f);
          if (mode == 0 || gi == G.nv4 - 1) {
            A = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int r = 0; r < P.X.nranks; ++r) {
              const float4 v = ldcg_f4(reinterpret_cast<const float4*>(mb + mbox_off((int)(epoch & 1), P.np, p, P.X.nranks, r, 0, G.nv4)) + gi);
              A.x += v.x; A.y += v.y; A.z += v.z; A.w += v.w;
            }
          } else {
            A = make_float4(kNegBig, kNegBig, kNegBig, kNegBig);
            for (int r = 0; r < P.X.nranks; ++r) {
              const float4 v = ldcg_f4(reinterpret_cast<const float4*>(mb + mbox_off((int)(epoch & 1), P.np, p, P.X.nranks, r, 0, G.nv4)) + gi);
              A.x = fmaxf(A.x, v.x); A.y = fmaxf(A.y, v.y); A.z = fmaxf(A.z, v.z); A.w = fmaxf(A.w, v.w);
            }
            for (int r = 0; r < P.X.nranks; ++r) {
              const float4 m = ldcg_f4(reinterpret_cast<const float4*>(mb + mbox_off((int)(epoch & 1), P.np, p, P.X.nranks, r, 0, G.nv4)) + gi);
              const float4 s = ldcg_f4(reinterpret_cast<const float4*>(mb + mbox_off((int)(epoch & 1), P.np, p, P.X.nranks, r, 1, G.nv4)) + gi);
              Bv.x += s.x * fast_exp2(m.x - A.x); Bv.y += s.y * fast_exp2(m.y - A.y);
              Bv.z += s.z * fast_exp2(m.z - A.z); Bv.w += s.w * fast_exp2(m.w - A.w);
            }
          }
          finalize(gi, A, Bv);
        }
      }
    }
    grid.sync();
    // ------------------------------------------------------------------ decide
    if (!done) {
      if (mode == 0 && ld_volatile_i(&st->trip) != 0) {
        // redo iteration `it` in the log domain from the last good potentials (history row `it`)
        mode = 1;
#pragma unroll
        for (int m = 0; m < NG; ++m) {
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (colok[m]) v = ldcg_f4(reinterpret_cast<const float4*>(vh + (long long)it * B) + tid + kNT * m);
          b[4 * m] = v.x; b[4 * m + 1] = v.y; b[4 * m + 2] = v.z; b[4 * m + 3] = v.w;
        }
      } else {
#pragma unroll
        for (int m = 0; m < NG; ++m) {
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (colok[m]) v = ldcg_f4(reinterpret_cast<const float4*>(bv) + tid + kNT * m);
          b[4 * m] = v.x; b[4 * m + 1] = v.y; b[4 * m + 2] = v.z; b[4 * m + 3] = v.w;
        }
        const bool may_stop = (P.exit_on_index ? (it >= P.Lmin) : (it + 1 >= P.Lmin)) && (it + 1 < P.L);
        ++it;
        if (may_stop) {
          const float err = *reinterpret_cast<const volatile float*>(&st->err) / kscale;
          if (P.thresh > err) done = true;            // gan_utils.py:157-160 / :114-117
        }
        if (it >= P.L) done = true;
        if (done && G.cl == 0 && tid == 0) { st->nits = it; st->done = 1; }
      }
    }
  }

  // ------------------------------------------------------------------ sharp cost sum(pi * C) over this CTA's rows
  const int nits = ld_volatile_i(&st->nits);
  {
    float v[CPT];
#pragma unroll
    for (int m = 0; m < NG; ++m) {
      float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
      if (colok[m]) q = ldcg_f4(reinterpret_cast<const float4*>(vh + (long long)nits * B) + tid + kNT * m);
      v[4 * m] = q.x; v[4 * m + 1] = q.y; v[4 * m + 2] = q.z; v[4 * m + 3] = q.w;
    }
    float s0 = 0.f, s1 = 0.f;
    for (int row = G.r_begin; row < G.r_end; ++row) {
      const float ui = ldcg_f(uh + (long long)nits * B + P.row0 + row);
      const float4* src = reinterpret_cast<const float4*>(Cp + (long long)row * B);
#pragma unroll
      for (int m = 0; m < NG; ++m) {
        if (colok[m]) {
          const float4 q = src[tid + kNT * m];
          const float cq[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float ch = (cq[e] - c0) * kscale;
            const float pi = fast_exp2(ui + v[4 * m + e] - ch);
            s0 += pi;
            s1 = fmaf(pi, ch, s1);
          }
        }
      }
    }
    float r2[2] = {s0, s1};
    block_reduce_rg<2, false>(r2, red, parity);
    if (tid == 0) {
      float* slot = pa + (long long)G.cl * G.nv4 * 4;
      slot[0] = r2[0];
      slot[1] = r2[1];
    }
  }
  grid.sync();
  if (G.cl == 0 && tid == 0) {
    float s0 = 0.f, s1 = 0.f;
    for (int c = 0; c < G.Gp; ++c) {
      s0 += ldcg_f(pa + (long long)c * G.nv4 * 4);
      s1 += ldcg_f(pa + (long long)c * G.nv4 * 4 + 1);
    }
    if (P.partial_cost) {                 // partial sums of this rank's rows; the caller all-reduces them
      P.cost[2 * p] = s0;
      P.cost[2 * p + 1] = s1;
    } else {
      P.cost[p] = s1 / kscale + c0 * s0;
    }
    P.nits[p] = nits;
  }
}

// ---------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------
struct PBwd {
  const float* C;
  const float* Kt;
  const float* alpha;
  const float* shift;
  const float *u_hist, *v_hist;
  const int32_t* nits;
  const float* gcost;       // [np] device (or nullptr: gcost_host)
  float gcost_host;
  float* Cbar;              // [np][Brows][B], zeroed by the host
  float *rowF, *colF;       // [np][Brows][2L], [np][B][2L]
  float* ubar;              // [np][Brows] seed carry
  float* vbar;              // [np][B]
  float* part_a;            // [np][Gmax][nv4 * 4]
  PState* state;
  int np, Brows, B, row0, L, resident, shift_stride, Gmax;
  float kscale, ahat, inv_eps;
  Xchg X;
};

template <int CPT, int RG>
__global__ void __launch_bounds__(kNT, 1) sk_persist_bwd_kernel(const __grid_constant__ PBwd P) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ float4 smem4[];
  float* red = reinterpret_cast<float*>(smem4);
  float* srows = red + kRedFloats;
  const int tid = threadIdx.x;
  const int B = P.B, Brows = P.Brows;
  const Geo G = make_geo(P.np, Brows, B);
  const int p = G.p;
  const long long hs = (long long)(P.L + 1) * B;
  const float* Cp = P.C + (long long)p * Brows * B;
  const float* Ktp = P.Kt + (long long)p * Brows * B;
  const float* alp = P.alpha + (long long)p * Brows;
  const float* uh = P.u_hist + p * hs;
  const float* vh = P.v_hist + p * hs;
  float* Cb = P.Cbar + (long long)p * Brows * B;
  float* rowF = P.rowF + (long long)p * Brows * 2 * P.L;
  float* colF = P.colF + (long long)p * B * 2 * P.L;
  float* ubar_g = P.ubar + (long long)p * Brows;
  float* vbar_g = P.vbar + (long long)p * B;
  float* pa = P.part_a + (long long)p * P.Gmax * G.nv4 * 4;
  PState* st = P.state + p;
  const float c0 = P.shift[(long long)p * P.shift_stride];
  const float kscale = P.kscale, ahat = P.ahat, inv_eps = P.inv_eps;
  const float Bf = exp2f(-P.ahat);
  const float g = P.gcost ? P.gcost[p] : P.gcost_host;
  const int nits = P.nits[p];
  const int F2 = 2 * P.L;
  constexpr int NG = CPT / 4;
  bool colok[NG];
#pragma unroll
  for (int m = 0; m < NG; ++m) colok[m] = 4 * (tid + kNT * m) < B;
  int parity = 0;
  unsigned long long epoch = P.X.epoch0;
  const bool multi = P.X.nranks > 1;

  if (P.resident) {
    for (int r = G.r_begin; r < G.r_end; ++r) {
      const float4* src = reinterpret_cast<const float4*>(Ktp + (long long)r * B);
      float4* dst = reinterpret_cast<float4*>(srows + (long long)(r - G.r_begin) * B);
      for (int q = tid; q < B / 4; q += kNT) dst[q] = src[q];
    }
  }
  // factor range check over the executed history: |uhat^k_i - alpha_i| and |vhat^k_j| must stay below kExpLim
  {
    bool bad = false;
    for (int row = G.r_begin; row < G.r_end; ++row) {
      const float al = alp[row];
      for (int k = 1 + tid; k <= nits; k += kNT) bad |= !(fabsf(uh[(long long)k * B + P.row0 + row] - al) < kExpLim);
    }
    for (int gi = G.g_begin; gi < min(G.g_end, B / 4); ++gi)
      for (int k = tid; k <= nits; k += kNT) {
        const float4 v = reinterpret_cast<const float4*>(vh + (long long)k * B)[gi];
        bad |= !(fabsf(v.x) < kExpLim && fabsf(v.y) < kExpLim && fabsf(v.z) < kExpLim && fabsf(v.w) < kExpLim);
      }
    if (bad) st->trip = 1;
  }
  __syncthreads();
  grid.sync();
  if (multi) {                                         // a problem goes direct on every rank or on none
    ++epoch;
    const bool owner = (G.g_end == G.nv4) && (G.g_begin < G.g_end) && tid == 0;
    if (owner) xchg_push(P.X, epoch, P.np, p, G.nv4, G.nv4 - 1, 0, make_float4((float)ld_volatile_i(&st->trip), 0.f, 0.f, 0.f));
    xchg_sync(P.X, epoch, st);
    if (owner) {
      float tsum = 0.f;
      for (int r = 0; r < P.X.nranks; ++r)
        tsum += ldcg_f(P.X.mbox[P.X.rank] + mbox_off((int)(epoch & 1), P.np, p, P.X.nranks, r, 0, G.nv4) + (size_t)(G.nv4 - 1) * 4);
      if (tsum > 0.f) st->trip = 1;
    }
    grid.sync();
  }
  const int direct = ld_volatile_i(&st->trip);        // uniform over the problem's CTAs (and ranks)

  auto load_vec = [&](const float* src, float (&dst)[CPT]) {
#pragma unroll
    for (int m = 0; m < NG; ++m) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (colok[m]) v = ldcg_f4(reinterpret_cast<const float4*>(src) + tid + kNT * m);
      dst[4 * m] = v.x; dst[4 * m + 1] = v.y; dst[4 * m + 2] = v.z; dst[4 * m + 3] = v.w;
    }
  };
  auto load_rows = [&](const float* basep, bool use_resident, int g0, float (&kr)[RG][CPT]) {
#pragma unroll
    for (int r = 0; r < RG; ++r) {
      const int row = g0 + r;
      const bool rok = row < G.r_end;
      const float* base = use_resident ? srows + (long long)(row - G.r_begin) * B : basep + (long long)row * B;
      const float4* src = reinterpret_cast<const float4*>(base);
#pragma unroll
      for (int m = 0; m < NG; ++m) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (rok && colok[m]) v = src[tid + kNT * m];
        kr[r][4 * m + 0] = v.x; kr[r][4 * m + 1] = v.y; kr[r][4 * m + 2] = v.z; kr[r][4 * m + 3] = v.w;
      }
    }
  };
  // combine the column partials of the problem's CTAs (and ranks); `fin(gi, T4)` consumes the totals
  auto combine = [&](auto fin) {
    ++epoch;
    for (int gi = G.g_begin + tid; gi < min(G.g_end, B / 4); gi += kNT) {
      float4 A = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int c = 0; c < G.Gp; ++c) {
        const float4 v = ldcg_f4(reinterpret_cast<const float4*>(pa + (long long)c * G.nv4 * 4) + gi);
        A.x += v.x; A.y += v.y; A.z += v.z; A.w += v.w;
      }
      if (multi) xchg_push(P.X, epoch, P.np, p, G.nv4, gi, 0, A);
      else fin(gi, A);
    }
    if (multi) {
      xchg_sync(P.X, epoch, st);
      const float* mb = P.X.mbox[P.X.rank];
      for (int gi = G.g_begin + tid; gi < min(G.g_end, B / 4); gi += kNT) {
        float4 A = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = 0; r < P.X.nranks; ++r) {
          const float4 v = ldcg_f4(reinterpret_cast<const float4*>(mb + mbox_off((int)(epoch & 1), P.np, p, P.X.nranks, r, 0, G.nv4)) + gi);
          A.x += v.x; A.y += v.y; A.z += v.z; A.w += v.w;
        }
        fin(gi, A);
      }
    }
  };
  auto write_partials = [&](const float (&tcol)[CPT]) {
    float4* pa4 = reinterpret_cast<float4*>(pa + (long long)G.cl * G.nv4 * 4);
#pragma unroll
    for (int m = 0; m < NG; ++m)
      if (colok[m]) pa4[tid + kNT * m] = make_float4(tcol[4 * m], tcol[4 * m + 1], tcol[4 * m + 2], tcol[4 * m + 3]);
  };

  // ------------------------------------------------------------------ seed: pi at the final potentials
  float vn[CPT], tcol[CPT];
  load_vec(vh + (long long)nits * B, vn);
#pragma unroll
  for (int c = 0; c < CPT; ++c) tcol[c] = 0.f;
  for (int g0 = G.r_begin; g0 < G.r_end; g0 += RG) {
    float kr[RG][CPT];
    load_rows(Cp, false, g0, kr);
    float part[RG];
#pragma unroll
    for (int r = 0; r < RG; ++r) {
      const int row = g0 + r;
      const float ui = (row < G.r_end) ? uh[(long long)nits * B + P.row0 + row] : 0.f;
      float a0 = 0.f;
#pragma unroll
      for (int c = 0; c < CPT; ++c) {
        const bool ok = colok[c >> 2] && row < G.r_end;
        const float d = kr[r][c] - c0;
        const float w = ok ? g * fast_exp2(ui + vn[c] - d * kscale) * (d * inv_eps) : 0.f;   // g pi C/eps (shifted cost)
        a0 += w;
        tcol[c] += w;
      }
      part[r] = a0;
    }
    block_reduce_rg<RG, false>(part, red, parity);
#pragma unroll
    for (int r = 0; r < RG; ++r)
      if (tid == r && g0 + r < G.r_end) ubar_g[g0 + r] = part[r];
  }
  write_partials(tcol);
  grid.sync();
  combine([&](int gi, float4 T) { reinterpret_cast<float4*>(vbar_g)[gi] = T; });     // vbar = + column sums
  grid.sync();

  // ------------------------------------------------------------------ reverse steps k = nits .. 1
  for (int k = nits; k >= 1; --k) {
    float pv[CPT];                      // scaling: p_j = b^k_j vbar_j ; direct: vbar_j
    {
      float vb[CPT], vk[CPT];
      load_vec(vbar_g, vb);
      load_vec(vh + (long long)k * B, vk);
#pragma unroll
      for (int c = 0; c < CPT; ++c) pv[c] = direct ? vb[c] : fast_exp2(vk[c]) * vb[c];
      if (direct) {
#pragma unroll
        for (int c = 0; c < CPT; ++c) vn[c] = vk[c];       // vhat^k
      }
    }
    float vkm1[CPT];
    if (direct) load_vec(vh + (long long)(k - 1) * B, vkm1);
#pragma unroll
    for (int c = 0; c < CPT; ++c) tcol[c] = 0.f;
    for (int g0 = G.r_begin; g0 < G.r_end; g0 += RG) {
      float kr[RG][CPT];
      load_rows(direct ? Cp : Ktp, !direct && P.resident, g0, kr);
      float part[RG], ui[RG];
#pragma unroll
      for (int r = 0; r < RG; ++r) {
        const int row = g0 + r;
        ui[r] = (row < G.r_end) ? uh[(long long)k * B + P.row0 + row] : 0.f;
        float a0 = 0.f, a1 = 0.f;
        if (!direct) {
#pragma unroll
          for (int c = 0; c < CPT; c += 2) { a0 = fmaf(kr[r][c], pv[c], a0); a1 = fmaf(kr[r][c + 1], pv[c + 1], a1); }
        } else {
          // kr <- Chat ; w = exp2(u - ahat + v^k - Chat) vbar ; Cbar += w
          float* crow = Cb + (long long)row * B;
#pragma unroll
          for (int c = 0; c < CPT; ++c) {
            const bool ok = colok[c >> 2] && row < G.r_end;
            kr[r][c] = (kr[r][c] - c0) * kscale;
            const float w = ok ? fast_exp2(ui[r] - ahat + vn[c] - kr[r][c]) * pv[c] : 0.f;
            if (ok) {
              const int col = 4 * (tid + kNT * (c >> 2)) + (c & 3);
              crow[col] += w;
            }
            a0 += w;
          }
        }
        part[r] = a0 + a1;
      }
      block_reduce_rg<RG, false>(part, red, parity);
      float q[RG];
#pragma unroll
      for (int r = 0; r < RG; ++r) {
        const int row = g0 + r;
        const bool rok = row < G.r_end;
        const float carry = (k == nits && rok) ? ubar_g[row] : 0.f;
        if (!direct) {
          const float Ba = rok ? Bf * fast_exp2(ui[r] - alp[min(row, Brows - 1)]) : 0.f;
          const float ub = carry - Ba * part[r];
          q[r] = Ba * ub;
          if (tid == r && rok) {
            rowF[(long long)row * F2 + 2 * (k - 1)] = Ba;
            rowF[(long long)row * F2 + 2 * (k - 1) + 1] = q[r];
          }
        } else {
          q[r] = carry - part[r];           // ubar_i
        }
      }
      if (!direct) {
#pragma unroll
        for (int r = 0; r < RG; ++r)
#pragma unroll
          for (int c = 0; c < CPT; ++c) tcol[c] = fmaf(kr[r][c], q[r], tcol[c]);
      } else {
#pragma unroll
        for (int r = 0; r < RG; ++r) {
          const int row = g0 + r;
          if (row >= G.r_end) continue;
          float* crow = Cb + (long long)row * B;
#pragma unroll
          for (int c = 0; c < CPT; ++c) {
            if (colok[c >> 2]) {
              const float w = fast_exp2(ui[r] - ahat + vkm1[c] - kr[r][c]) * q[r];
              const int col = 4 * (tid + kNT * (c >> 2)) + (c & 3);
              crow[col] += w;
              tcol[c] += w;
            }
          }
        }
      }
    }
    write_partials(tcol);
    grid.sync();
    combine([&](int gi, float4 T) {
      const float4 vk = reinterpret_cast<const float4*>(vh + (long long)k * B)[gi];
      const float4 vm = reinterpret_cast<const float4*>(vh + (long long)(k - 1) * B)[gi];
      const float4 vb = ldcg_f4(reinterpret_cast<const float4*>(vbar_g) + gi);
      const float tv[4] = {T.x, T.y, T.z, T.w}, vkv[4] = {vk.x, vk.y, vk.z, vk.w}, vmv[4] = {vm.x, vm.y, vm.z, vm.w},
                  vbv[4] = {vb.x, vb.y, vb.z, vb.w};
      float out[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (!direct) {
          const float bm = fast_exp2(vmv[e]);
          const long long j = (long long)gi * 4 + e;
          colF[j * F2 + 2 * (k - 1)] = fast_exp2(vkv[e]) * vbv[e];       // the p_j used in this step
          colF[j * F2 + 2 * (k - 1) + 1] = bm;
          out[e] = -bm * tv[e];
        } else {
          out[e] = -tv[e];
        }
      }
      reinterpret_cast<float4*>(vbar_g)[gi] = make_float4(out[0], out[1], out[2], out[3]);
    });
    grid.sync();
  }
}

// ---------------------------------------------------------------------------------------------
// Cbar = [Cbar +] g pi (1 - C/eps) + Kt .* (RowF ColF^T)      (64 x 64 tiles, 256 threads, 4 x 4 per thread)
// ---------------------------------------------------------------------------------------------
constexpr int FC = 32;
__global__ void __launch_bounds__(256) sk_cbar_final_kernel(PBwd P) {
  __shared__ float rs[64][FC + 1], cs[64][FC + 1];
  const int p = blockIdx.z;
  const int B = P.B, Brows = P.Brows;
  const int i0 = blockIdx.y * 64, j0 = blockIdx.x * 64;
  const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
  const long long hs = (long long)(P.L + 1) * B;
  const float* Cp = P.C + (long long)p * Brows * B;
  const float* Ktp = P.Kt + (long long)p * Brows * B;
  const float* uh = P.u_hist + p * hs;
  const float* vh = P.v_hist + p * hs;
  float* Cb = P.Cbar + (long long)p * Brows * B;
  const float* rowF = P.rowF + (long long)p * Brows * 2 * P.L;
  const float* colF = P.colF + (long long)p * B * 2 * P.L;
  const float c0 = P.shift[(long long)p * P.shift_stride];
  const float g = P.gcost ? P.gcost[p] : P.gcost_host;
  const int nits = P.nits[p];
  const int direct = P.state[p].trip;
  const int F2 = 2 * P.L, nf = 2 * nits;
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[a][c] = 0.f;
  if (!direct) {
    for (int f0 = 0; f0 < nf; f0 += FC) {
      __syncthreads();
      for (int e = t; e < 64 * FC; e += 256) {
        const int rr = e / FC, f = f0 + e % FC;
        rs[rr][e % FC] = (f < nf && i0 + rr < Brows) ? rowF[(long long)(i0 + rr) * F2 + f] : 0.f;
        cs[rr][e % FC] = (f < nf && j0 + rr < B) ? colF[(long long)(j0 + rr) * F2 + f] : 0.f;
      }
      __syncthreads();
#pragma unroll 8
      for (int f = 0; f < FC; ++f) {
        float rv[4], cv[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) { rv[a] = rs[ty * 4 + a][f]; cv[a] = cs[tx * 4 + a][f]; }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[a][c] = fmaf(rv[a], cv[c], acc[a][c]);
      }
    }
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int i = i0 + ty * 4 + a;
    if (i >= Brows) continue;
    const float ui = uh[(long long)nits * B + P.row0 + i];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int j = j0 + tx * 4 + c;
      if (j >= B) continue;
      const long long idx = (long long)i * B + j;
      const float d = Cp[idx] - c0;
      const float pi = g * fast_exp2(ui + vh[(long long)nits * B + j] - d * P.kscale);
      float out = pi * (1.f - d * P.inv_eps);          // shifted cost: sum(pi) == 1 has zero gradient
      if (!direct) out = fmaf(Ktp[idx], acc[a][c], out);
      else out += Cb[idx];
      Cb[idx] = out;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
struct PersistLayout {
  size_t off_state, off_shift, off_Kt, off_alpha, off_pa, off_pb, off_bvec, off_rowF, off_colF, off_ubar, off_vbar, total;
  int Gmax, nv4;
};
PersistLayout persist_layout(int np, int Brows, int B, int L) {
  PersistLayout l{};
  l.Gmax = num_sms();
  l.nv4 = B / 4 + 1;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes, 256); return r; };
  l.off_state = take((size_t)np * sizeof(PState));
  l.off_shift = take((size_t)np * 4);
  l.off_Kt = take((size_t)np * Brows * B * 4);
  l.off_alpha = take((size_t)np * Brows * 4);
  l.off_pa = take((size_t)np * l.Gmax * l.nv4 * 16);
  l.off_pb = take((size_t)np * l.Gmax * l.nv4 * 16);
  l.off_bvec = take((size_t)np * B * 4);
  l.off_rowF = take((size_t)np * Brows * 2 * L * 4);
  l.off_colF = take((size_t)np * B * 2 * L * 4);
  l.off_ubar = take((size_t)np * Brows * 4);
  l.off_vbar = take((size_t)np * B * 4);
  l.total = o;
  return l;
}

int cpt_for(int B) {
  const int per = (B + 4 * kNT - 1) / (4 * kNT) * 4;
  return per <= 4 ? 4 : per <= 8 ? 8 : per <= 16 ? 16 : 0;
}
size_t resident_smem(int np, int Brows, int B, int grid) {
  const int Gp_min = grid / np;                  // the problem with the fewest CTAs
  if (Gp_min < 1) return 0;
  const int rpc = (Brows + Gp_min - 1) / Gp_min;
  const size_t need = (size_t)rpc * B * 4;
  return need <= 200 * 1024 ? need : 0;
}

template <typename Params>
int launch_coop(void (*kernel)(Params), const Params& P, int grid, size_t smem, cudaStream_t st) {
  void* args[] = {const_cast<Params*>(&P)};
  KCCOT_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(kernel), dim3(grid), dim3(kNT), args, smem, st));
  count_launch();
  return KCCOT_OK;
}

template <int CPT, int RG>
int launch_fwd_t(const PFwd& P, int grid, size_t smem, cudaStream_t st) {
  static size_t attr[kMaxDevices] = {};
  if (smem_attr_needed(attr, smem))
    KCCOT_CUDA(cudaFuncSetAttribute(sk_persist_fwd_kernel<CPT, RG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  return launch_coop(sk_persist_fwd_kernel<CPT, RG>, P, grid, smem, st);
}
template <int CPT, int RG>
int launch_bwd_t(const PBwd& P, int grid, size_t smem, cudaStream_t st) {
  static size_t attr[kMaxDevices] = {};
  if (smem_attr_needed(attr, smem))
    KCCOT_CUDA(cudaFuncSetAttribute(sk_persist_bwd_kernel<CPT, RG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  return launch_coop(sk_persist_bwd_kernel<CPT, RG>, P, grid, smem, st);
}
}  // namespace

bool persist_supported(int Brows, int B, int L) {
  return B > kSmallSinkhornMaxB && B % 4 == 0 && cpt_for(B) != 0 && Brows >= 1 && L >= 0;
}
size_t persist_workspace_bytes(int np, int Brows, int B, int L) { return persist_layout(np, Brows, B, L).total; }

// Shards pass the communicator; single GPU: comm == nullptr.  `shift_dev`: optional [np] device floats (global
// minimum of C over all ranks); nullptr -> computed here over the given rows.
int persist_sinkhorn_fwd(const float* C, int np, int Brows, int B, int row0, float eps, int L, int Lmin, float thresh,
                         int exit_on_index, float* u_hist, float* v_hist, int32_t* nits, float* cost, void* ws,
                         const ShardComm* comm, const float* shift_dev, cudaStream_t st) {
  const PersistLayout l = persist_layout(np, Brows, B, L);
  char* w = (char*)ws;
  PState* state = (PState*)(w + l.off_state);
  float* shift = (float*)(w + l.off_shift);
  float* Kt = (float*)(w + l.off_Kt);
  float* alpha = (float*)(w + l.off_alpha);
  const float kscale = kLog2e / eps, ahat = -log2f((float)B);
  const long long n = (long long)Brows * B;
  persist_init_kernel<<<1, 32, 0, st>>>(state, shift_dev ? nullptr : shift, 1, np);
  KCCOT_LAUNCH_CHECK();
  if (shift_dev) {
    KCCOT_CUDA(cudaMemcpyAsync(shift, shift_dev, (size_t)np * 4, cudaMemcpyDeviceToDevice, st));
  } else {
    persist_min_kernel<<<dim3((unsigned)min((long long)2 * num_sms(), (n + 255) / 256), np), 256, 0, st>>>(C, n, n, shift, 1);
    KCCOT_LAUNCH_CHECK();
  }
  absorb_rows_kernel<<<dim3(Brows, np), 256, 0, st>>>(C, Brows, B, n, kscale, shift, 1, Kt, n, alpha, Brows);
  KCCOT_LAUNCH_CHECK();
  const long long hs = (long long)(L + 1) * B;
  for (int p = 0; p < np; ++p) {                       // history row 0: u = v = 0
    KCCOT_CUDA(cudaMemsetAsync(u_hist + p * hs, 0, (size_t)B * 4, st));
    KCCOT_CUDA(cudaMemsetAsync(v_hist + p * hs, 0, (size_t)B * 4, st));
  }
  PFwd P{};
  P.C = C; P.Kt = Kt; P.alpha = alpha; P.shift = shift;
  P.u_hist = u_hist; P.v_hist = v_hist; P.nits = nits; P.cost = cost;
  P.part_a = (float*)(w + l.off_pa); P.part_b = (float*)(w + l.off_pb); P.bvec = (float*)(w + l.off_bvec);
  P.state = state;
  P.np = np; P.Brows = Brows; P.B = B; P.row0 = row0; P.L = L; P.Lmin = Lmin; P.exit_on_index = exit_on_index;
  P.shift_stride = 1; P.Gmax = l.Gmax;
  P.kscale = kscale; P.ahat = ahat; P.thresh = thresh;
  P.X.nranks = 1; P.X.rank = 0; P.X.epoch0 = 0;
  P.partial_cost = comm ? 1 : 0;
  if (comm) {
    P.X.nranks = comm->nranks; P.X.rank = comm->rank; P.X.epoch0 = comm->epoch0;
    for (int r = 0; r < comm->nranks; ++r) { P.X.mbox[r] = comm->mbox[r]; P.X.flags[r] = comm->flags[r]; }
  }
  const int grid = min(min(num_sms(), kFlagStride), np * Brows);     // (one exchange flag per CTA: <= kFlagStride)
  const size_t res = resident_smem(np, Brows, B, grid);
  P.resident = res ? 1 : 0;
  const size_t smem = kRedFloats * 4 + res;
  const int cpt = cpt_for(B);
  if (L == 0) {                                         // no iterations: the kernel's loop expects L >= 1
    set_error("persistent Sinkhorn needs L >= 1");
    return KCCOT_EINVAL;
  }
  if (cpt == 4) return launch_fwd_t<4, 4>(P, grid, smem, st);
  if (cpt == 8) return launch_fwd_t<8, 4>(P, grid, smem, st);
  return launch_fwd_t<16, 2>(P, grid, smem, st);
}

int persist_sinkhorn_bwd(const float* C, int np, int Brows, int B, int row0, float eps, int L, const float* u_hist,
                         const float* v_hist, const int32_t* nits, const float* gcost, float gcost_host, float* Cbar,
                         void* ws, const ShardComm* comm, const float* shift_dev, cudaStream_t st) {
  const PersistLayout l = persist_layout(np, Brows, B, L);
  char* w = (char*)ws;
  PState* state = (PState*)(w + l.off_state);
  float* shift = (float*)(w + l.off_shift);
  float* Kt = (float*)(w + l.off_Kt);
  float* alpha = (float*)(w + l.off_alpha);
  const float kscale = kLog2e / eps, ahat = -log2f((float)B);
  const long long n = (long long)Brows * B;
  // the saved potentials contain the forward's shift (the global minimum): recompute the same value
  persist_init_kernel<<<1, 32, 0, st>>>(state, shift_dev ? nullptr : shift, 1, np);
  KCCOT_LAUNCH_CHECK();
  if (shift_dev) {
    KCCOT_CUDA(cudaMemcpyAsync(shift, shift_dev, (size_t)np * 4, cudaMemcpyDeviceToDevice, st));
  } else {
    persist_min_kernel<<<dim3((unsigned)min((long long)2 * num_sms(), (n + 255) / 256), np), 256, 0, st>>>(C, n, n, shift, 1);
    KCCOT_LAUNCH_CHECK();
  }
  absorb_rows_kernel<<<dim3(Brows, np), 256, 0, st>>>(C, Brows, B, n, kscale, shift, 1, Kt, n, alpha, Brows);
  KCCOT_LAUNCH_CHECK();
  KCCOT_CUDA(cudaMemsetAsync(Cbar, 0, (size_t)np * n * 4, st));
  PBwd P{};
  P.C = C; P.Kt = Kt; P.alpha = alpha; P.shift = shift;
  P.u_hist = u_hist; P.v_hist = v_hist; P.nits = nits; P.gcost = gcost; P.gcost_host = gcost_host; P.Cbar = Cbar;
  P.rowF = (float*)(w + l.off_rowF); P.colF = (float*)(w + l.off_colF);
  P.ubar = (float*)(w + l.off_ubar); P.vbar = (float*)(w + l.off_vbar);
  P.part_a = (float*)(w + l.off_pa);
  P.state = state;
  P.np = np; P.Brows = Brows; P.B = B; P.row0 = row0; P.L = L; P.shift_stride = 1; P.Gmax = l.Gmax;
  P.kscale = kscale; P.ahat = ahat; P.inv_eps = 1.f / eps;
  P.X.nranks = 1; P.X.rank = 0; P.X.epoch0 = 0;
  if (comm) {
    P.X.nranks = comm->nranks; P.X.rank = comm->rank; P.X.epoch0 = comm->epoch0;
    for (int r = 0; r < comm->nranks; ++r) { P.X.mbox[r] = comm->mbox[r]; P.X.flags[r] = comm->flags[r]; }
  }
  const int grid = min(min(num_sms(), kFlagStride), np * Brows);     // (one exchange flag per CTA: <= kFlagStride)
  const size_t res = resident_smem(np, Brows, B, grid);
  P.resident = res ? 1 : 0;
  const size_t smem = kRedFloats * 4 + res;
  const int cpt = cpt_for(B);
  int rc;
  if (cpt == 4) rc = launch_bwd_t<4, 4>(P, grid, smem, st);
  else if (cpt == 8) rc = launch_bwd_t<8, 4>(P, grid, smem, st);
  else rc = launch_bwd_t<16, 2>(P, grid, smem, st);
  if (rc) return rc;
  sk_cbar_final_kernel<<<dim3((B + 63) / 64, (Brows + 63) / 64, np), 256, 0, st>>>(P);
  KCCOT_LAUNCH_CHECK();
  return KCCOT_OK;
}

}  // namespace kccot

namespace kccot {
size_t persist_mailbox_floats(int np, int nranks, int B) {
  return (size_t)2 * np * nranks * 2 * ((size_t)(B / 4 + 1) * 4);
}
}  // namespace kccot
