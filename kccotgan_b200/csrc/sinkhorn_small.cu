// Log-domain Sinkhorn for B <= 64: one CTA per problem, the cost matrix held in REGISTERS in both
// row-major and column-major slices for all L iterations (gan_utils.py:151-164), and the fully
// unrolled reverse pass with the adjoint of C accumulated in registers.
//
// 4 threads share a row (or a column): thread (i, q) owns C[i][q*EPT .. q*EPT+EPT) and
// C[q*EPT .. q*EPT+EPT)[i].  A half-iteration is: one FADD + one ex2 per element, a 4-lane shuffle
// reduction, one __syncthreads.  Internally everything is in log2 units and the cost is shifted by
// its minimum:  Chat = (C - c0) * log2(e)/eps,  uhat = (u - c0) * log2(e)/eps,  vhat = v * log2(e)/eps,
// so that the exponent arguments stay O(spread/eps) instead of O(|C|/eps).
#include "common.cuh"
#include "sinkhorn.cuh"

namespace kccot {

namespace {
constexpr float kBig = 1e30f;   // padding cost: exp2(anything - kBig) == 0, and 0 * kBig == 0

__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}
__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
  return v;
}

template <int NT>
__device__ __forceinline__ float block_reduce(float v, float* red, bool is_min) {
  v = is_min ? warp_min(v) : warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = is_min ? kBig : 0.f;
  const int nw = (blockDim.x + 31) >> 5;
  for (int w = 0; w < nw; ++w) r = is_min ? fminf(r, red[w]) : r + red[w];
  return r;
}

// loads both register slices of C, shifted by its minimum and scaled to log2 units
template <int EPT>
__device__ __forceinline__ float load_slices(const float* __restrict__ Cn, int B, float kscale, float* red,
                                             float (&Cr)[EPT], float (&Cc)[EPT]) {
  const int i = threadIdx.x >> 2, q = threadIdx.x & 3;
  float mn = kBig;
#pragma unroll
  for (int e = 0; e < EPT; ++e) {
    const int j = q * EPT + e;
    Cr[e] = (i < B && j < B) ? Cn[(long long)i * B + j] : kBig;
    Cc[e] = (i < B && j < B) ? Cn[(long long)j * B + i] : kBig;
    mn = fminf(mn, Cr[e]);
  }
  const float c0 = block_reduce<0>(mn, red, true);
#pragma unroll
  for (int e = 0; e < EPT; ++e) {
    Cr[e] = (Cr[e] < kBig) ? (Cr[e] - c0) * kscale : kBig;
    Cc[e] = (Cc[e] < kBig) ? (Cc[e] - c0) * kscale : kBig;
  }
  return c0;
}
}  // namespace

template <int EPT>
__global__ void __launch_bounds__(4 * 4 * EPT) sinkhorn_fwd_small_kernel(
    const float* __restrict__ C, int B, float eps, int L, int Lmin, float thresh, int exit_on_index,
    float* __restrict__ u_hist, float* __restrict__ v_hist, int32_t* __restrict__ nits_out,
    float* __restrict__ cost_out) {
  constexpr int BM = 4 * EPT;
  __shared__ float us[BM], vs[BM], red[32];
  __shared__ int stop_flag;
  const int n = blockIdx.x;
  const int tid = threadIdx.x, i = tid >> 2, q = tid & 3;
  const float kscale = kLog2e / eps;
  const float ahat = -log2f((float)B);
  float Cr[EPT], Cc[EPT];
  const float c0 = load_slices<EPT>(C + (long long)n * B * B, B, kscale, red, Cr, Cc);
  float* uh = u_hist + (long long)n * (L + 1) * B;
  float* vh = v_hist + (long long)n * (L + 1) * B;
  for (int t = tid; t < BM; t += blockDim.x) { us[t] = 0.f; vs[t] = 0.f; }
  if (tid < B) { uh[tid] = 0.f; vh[tid] = 0.f; }
  if (tid == 0) stop_flag = 0;
  __syncthreads();

  int nits = 0;
  for (int it = 0; it < L; ++it) {
    // ---- u update: rows ---------------------------------------------------------------------
    float t[EPT];
    float m = -kBig;
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
      t[e] = vs[q * EPT + e] - Cr[e];
      m = fmaxf(m, t[e]);
    }
    m = quad_max(m);
    float ssum = 0.f;
#pragma unroll
    for (int e = 0; e < EPT; ++e) ssum += fast_exp2(t[e] - m);
    ssum = quad_sum(ssum);
    const float unew = ahat - (m + fast_log2(ssum));
    float du = 0.f;
    if (q == 0 && i < B) {
      du = fabsf(unew - us[i]);
      us[i] = unew;
      uh[(long long)(it + 1) * B + i] = unew;
    }
    __syncthreads();
    // ---- v update: columns ------------------------------------------------------------------
    m = -kBig;
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
      t[e] = us[q * EPT + e] - Cc[e];
      m = fmaxf(m, t[e]);
    }
    m = quad_max(m);
    ssum = 0.f;
#pragma unroll
    for (int e = 0; e < EPT; ++e) ssum += fast_exp2(t[e] - m);
    ssum = quad_sum(ssum);
    const float vnew = ahat - (m + fast_log2(ssum));
    if (q == 0 && i < B) {
      vs[i] = vnew;
      vh[(long long)(it + 1) * B + i] = vnew;
    }
    __syncthreads();
    nits = it + 1;
    // ---- stopping rule (gan_utils.py:157-160 / :114-117); can only fire once the minimum count
    // is reached, so the reduction is skipped before that ---------------------------------------
    const bool may_stop = exit_on_index ? (it >= Lmin) : (nits >= Lmin);
    if (may_stop && nits < L) {
      const float err = block_reduce<0>(du, red, false) / kscale;
      if (tid == 0) stop_flag = (thresh > err) ? 1 : 0;
      __syncthreads();
      if (stop_flag) break;
    }
  }
  // ---- sharp cost sum(pi * C) ---------------------------------------------------------------
  float s1 = 0.f, s0 = 0.f;
  {
    const float ui = us[min(i, BM - 1)];
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
      const float pi = fast_exp2(ui + vs[q * EPT + e] - Cr[e]);
      s0 += pi;
      s1 = fmaf(pi, Cr[e], s1);
    }
    if (i >= B) { s0 = 0.f; s1 = 0.f; }
  }
  s1 = block_reduce<0>(s1, red, false);
  s0 = block_reduce<0>(s0, red, false);
  if (tid == 0) {
    cost_out[n] = s1 / kscale + c0 * s0;
    nits_out[n] = nits;
  }
}

template <int EPT>
__global__ void __launch_bounds__(4 * 4 * EPT) sinkhorn_bwd_small_kernel(
    const float* __restrict__ C, int B, float eps, int L, const float* __restrict__ u_hist,
    const float* __restrict__ v_hist, const int32_t* __restrict__ nits_in, const float* __restrict__ gcost,
    float* __restrict__ Cbar) {
  constexpr int BM = 4 * EPT;
  __shared__ float Us[2][BM], Vs[2][BM], ub[BM], vb[BM], red[32];
  const int n = blockIdx.x;
  const int tid = threadIdx.x, i = tid >> 2, q = tid & 3;
  const float kscale = kLog2e / eps;
  const float ahat = -log2f((float)B);
  float Cr[EPT], Cc[EPT], Gr[EPT], Gc[EPT];
  load_slices<EPT>(C + (long long)n * B * B, B, kscale, red, Cr, Cc);
  const float* uh = u_hist + (long long)n * (L + 1) * B;
  const float* vh = v_hist + (long long)n * (L + 1) * B;
  const int nits = nits_in[n];
  const float g = gcost[n];

  for (int t = tid; t < 2 * BM; t += blockDim.x) { (&Us[0][0])[t] = 0.f; (&Vs[0][0])[t] = 0.f; }
  for (int t = tid; t < BM; t += blockDim.x) { ub[t] = 0.f; vb[t] = 0.f; }
  __syncthreads();
  if (tid < B) {
    Us[nits & 1][tid] = uh[(long long)nits * B + tid];
    Vs[nits & 1][tid] = vh[(long long)nits * B + tid];
    if (nits >= 1) Vs[(nits - 1) & 1][tid] = vh[(long long)(nits - 1) * B + tid];
  }
  __syncthreads();
  // ---- adjoint seeds.  cost = sum(pi*C) = sum(pi*(C - c0)) + c0*sum(pi), and sum(pi) == 1 identically
  // in the inputs (the last v-update normalises every column of pi to 1/B), so the c0 term has zero
  // gradient: seed with C' = C - c0.  Cbar = pi (1 - C'/eps), ubar = rowsum(pi C')/eps, vbar =
  // colsum(pi C')/eps.  This removes the O(|C|/eps) cancellation the reference's fp32 gradient suffers.
  {
    const float* U = Us[nits & 1];
    const float* V = Vs[nits & 1];
    const float ui = U[min(i, BM - 1)], vi = V[min(i, BM - 1)];
    float ru = 0.f, rv = 0.f;
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
      const float ce_r = Cr[e] * kLn2;                       // (C - c0) / eps (row slice)
      const float pr = fast_exp2(ui + V[q * EPT + e] - Cr[e]);
      Gr[e] = pr * (1.f - ce_r);
      ru = fmaf(pr, ce_r, ru);
      const float ce_c = Cc[e] * kLn2;
      const float pc = fast_exp2(U[q * EPT + e] + vi - Cc[e]);
      rv = fmaf(pc, ce_c, rv);
      Gc[e] = 0.f;
    }
    ru = quad_sum(ru);
    rv = quad_sum(rv);
    if (q == 0 && i < B) { ub[i] = ru; vb[i] = rv; }
  }
  __syncthreads();

  for (int k = nits; k >= 1; --k) {
    const int b = k & 1;
    // prefetch the next step's potentials (L2) while this step computes
    float pu = 0.f, pv = 0.f;
    if (tid < B && k >= 2) {
      pu = uh[(long long)(k - 1) * B + tid];
      pv = vh[(long long)(k - 2) * B + tid];
    }
    // ---- through v^k = a - eps*LSE_i((u^k_i - C_ij)/eps):  Pv_ij = exp((u^k_i + v^k_j - a - C_ij)/eps)
    {
      const float ui = Us[b][min(i, BM - 1)] - ahat;
      float acc = 0.f;
#pragma unroll
      for (int e = 0; e < EPT; ++e) {
        const float w = fast_exp2(ui + Vs[b][q * EPT + e] - Cr[e]) * vb[q * EPT + e];
        Gr[e] += w;
        acc += w;
      }
      acc = quad_sum(acc);
      if (q == 0 && i < B) ub[i] = ((k == nits) ? ub[i] : 0.f) - acc;
    }
    __syncthreads();
    // ---- through u^k = a - eps*LSE_j((v^{k-1}_j - C_ij)/eps):  Pu_ij = exp((u^k_i + v^{k-1}_j - a - C_ij)/eps)
    {
      const float vj = Vs[b ^ 1][min(i, BM - 1)] - ahat;
      float acc = 0.f;
#pragma unroll
      for (int e = 0; e < EPT; ++e) {
        const float w = fast_exp2(Us[b][q * EPT + e] + vj - Cc[e]) * ub[q * EPT + e];
        Gc[e] += w;
        acc += w;
      }
      acc = quad_sum(acc);
      if (q == 0 && i < B) vb[i] = -acc;
      if (tid < B && k >= 2) {
        Us[b ^ 1][tid] = pu;     // u^{k-1}
        Vs[b][tid] = pv;         // v^{k-2}   (v^k is dead after the row phase above)
      }
    }
    __syncthreads();
  }
  // ---- Cbar = g * (Gr + Gc^T) ---------------------------------------------------------------
  float* out = Cbar + (long long)n * B * B;
  if (i < B) {
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
      const int j = q * EPT + e;
      if (j < B) out[(long long)i * B + j] = g * Gr[e];
    }
  }
  __syncthreads();
  if (i < B) {
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
      const int r = q * EPT + e;
      if (r < B) out[(long long)r * B + i] += g * Gc[e];
    }
  }
}

int launch_sinkhorn_fwd_small(const float* C, int nsolve, int B, float eps, int L, int Lmin, float thresh,
                              int exit_on_index, float* u_hist, float* v_hist, int32_t* nits, float* cost,
                              cudaStream_t st) {
  const int threads = ((4 * B + 31) / 32) * 32;
  if (B <= 32)
    sinkhorn_fwd_small_kernel<8><<<nsolve, threads, 0, st>>>(C, B, eps, L, Lmin, thresh, exit_on_index, u_hist,
                                                             v_hist, nits, cost);
  else
    sinkhorn_fwd_small_kernel<16><<<nsolve, threads, 0, st>>>(C, B, eps, L, Lmin, thresh, exit_on_index, u_hist,
                                                              v_hist, nits, cost);
  KCCOT_LAUNCH_CHECK();
  return KCCOT_OK;
}

int launch_sinkhorn_bwd_small(const float* C, int nsolve, int B, float eps, int L, const float* u_hist,
                              const float* v_hist, const int32_t* nits, const float* gcost, float* Cbar,
                              cudaStream_t st) {
  const int threads = ((4 * B + 31) / 32) * 32;
  if (B <= 32)
    sinkhorn_bwd_small_kernel<8><<<nsolve, threads, 0, st>>>(C, B, eps, L, u_hist, v_hist, nits, gcost, Cbar);
  else
    sinkhorn_bwd_small_kernel<16><<<nsolve, threads, 0, st>>>(C, B, eps, L, u_hist, v_hist, nits, gcost, Cbar);
  KCCOT_LAUNCH_CHECK();
  return KCCOT_OK;
}

}  // namespace kccot
