// Gaussian kernel smoothing of [B,H,T,W,C] videos (data_utils.py:503-521 temporal_convolution,
// :552-582 gaussian_convolution3D): separable REFLECT-padded 7-tap passes along T (mode 1) or
// H, T, W (mode 3), then division by the global maximum of the filtered tensor.
//
// A pass views the tensor as [outer, n, inner] and applies out[o,p,i] = sum_q A[p,q] in[o,q,i], with A the
// [n,n] matrix of the REFLECT-padded filter: A[p][reflect(p + k - R)] += w[k].  The 2R+1 weights travel as KERNEL
// ARGUMENTS (no filter matrix in device memory, no host-to-device copy when sigma is annealed every step,
// data_utils.py:584-586); each kernel derives the band coefficients it needs.  A (and its transpose, the
// adjoint) is banded with radius R, so every output row needs the rows p-R..p+R and 2R+1 coefficients
// coef[p][d].  R = 1..6 (kernel sizes 2..13; kernel_train.py:216 uses 6 -> R = 3, the class default 8 -> R = 4).
// Two kernels, neither with index arithmetic in its inner loop:
//   axis_col_kernel  (inner >= 64 or n*inner > 1024): one thread per (o, 4 consecutive inner elements) walks
//                    the n rows with a 7-row register window — every input element is loaded once, 16 bytes
//                    at a time, coalesced along `inner`;
//   axis_tile_kernel (small inner, e.g. the W pass with inner = C): a CTA copies a contiguous block of
//                    o_chunk * n * inner floats to shared memory, thread t = (p, i) keeps its 7 coefficients
//                    and row offsets in registers and produces out[o][p][i] for every o of the block.
// (The first version used one generic tile kernel with three integer divisions per element: 190 us for a
//  37.7 MB tensor, 20x off the HBM roofline.)
#include "common.cuh"

namespace kccot {
namespace {
constexpr int FT = 256;
constexpr int kTileElems = 8192;     // 32 KB of shared memory per tile
constexpr int kMaxRadius = 6;
constexpr int kMaxAxis = 1024;       // coefficient table: n * (2R+1) floats of dynamic shared memory

struct Taps {                        // by value in the kernel parameters
  float w[2 * kMaxRadius + 1];
  int r;
};
__device__ __forceinline__ int reflect(int q, int n) { return q < 0 ? -q : (q >= n ? 2 * (n - 1) - q : q); }
// A[p][q] of the REFLECT-padded filter (or its transpose)
template <int R>
__device__ __forceinline__ float band_coef(const Taps& tp, int n, int p, int q, bool transposed) {
  const int row = transposed ? q : p, col = transposed ? p : q;
  float a = 0.f;
#pragma unroll
  for (int k = 0; k < 2 * R + 1; ++k) a += (reflect(row + k - R, n) == col) ? tp.w[k] : 0.f;
  return a;
}

__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
  if (v >= 0.f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

struct BwdIn {            // input transform of the first adjoint pass: gz = gout/m + tie * (-S/m) / count
  const float* out;       // forward output (== 1.0f exactly where the max was attained)
  const float* maxval;
  const float* sums;      // [0] = sum(gout*out), [1] = number of tied maxima
};

template <int V> struct Vec;
template <> struct Vec<4> {
  float4 v;
  __device__ __forceinline__ void zero() { v = make_float4(0.f, 0.f, 0.f, 0.f); }
  __device__ __forceinline__ void load(const float* p) { v = *reinterpret_cast<const float4*>(p); }
  __device__ __forceinline__ void store(float* p) const { *reinterpret_cast<float4*>(p) = v; }
  __device__ __forceinline__ void fma(float c, const Vec& o) {
    v.x = fmaf(c, o.v.x, v.x); v.y = fmaf(c, o.v.y, v.y); v.z = fmaf(c, o.v.z, v.z); v.w = fmaf(c, o.v.w, v.w);
  }
  __device__ __forceinline__ void div(float d) { v.x /= d; v.y /= d; v.z /= d; v.w /= d; }
  __device__ __forceinline__ float hmax() const { return fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)); }
  __device__ __forceinline__ void bwd(float sc, float tie, const float* o) {
    const float4 t = *reinterpret_cast<const float4*>(o);
    v.x = v.x * sc + ((t.x == 1.0f) ? tie : 0.f); v.y = v.y * sc + ((t.y == 1.0f) ? tie : 0.f);
    v.z = v.z * sc + ((t.z == 1.0f) ? tie : 0.f); v.w = v.w * sc + ((t.w == 1.0f) ? tie : 0.f);
  }
};
template <> struct Vec<1> {
  float v;
  __device__ __forceinline__ void zero() { v = 0.f; }
  __device__ __forceinline__ void load(const float* p) { v = *p; }
  __device__ __forceinline__ void store(float* p) const { *p = v; }
  __device__ __forceinline__ void fma(float c, const Vec& o) { v = fmaf(c, o.v, v); }
  __device__ __forceinline__ void div(float d) { v /= d; }
  __device__ __forceinline__ float hmax() const { return v; }
  __device__ __forceinline__ void bwd(float sc, float tie, const float* o) { v = v * sc + ((*o == 1.0f) ? tie : 0.f); }
};

// MODE bit 0: write `out`; bit 1: reduce the global max into *gmax; bit 2: divide by *divisor;
// bit 3: apply the BwdIn transform on load; bit 4: use A transposed
template <int MODE, int V, int R>
__global__ void __launch_bounds__(FT) axis_col_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                      long long outer, int n, long long inner,
                                                      const Taps tp, float* gmax,
                                                      const float* divisor, BwdIn bw) {
  constexpr int kR = R, kTaps = 2 * R + 1;
  extern __shared__ float coef[];                          // coef[p][d] = A[p][p + d - R] (0 outside the matrix)
  for (int e = threadIdx.x; e < n * kTaps; e += FT) {
    const int p = e / kTaps, q = p + (e % kTaps) - kR;
    coef[e] = (q >= 0 && q < n) ? band_coef<R>(tp, n, p, q, (MODE & 16) != 0) : 0.f;
  }
  __syncthreads();
  const long long iv = inner / V;
  const long long gid = (long long)blockIdx.x * FT + threadIdx.x;
  float lmax = -3.0e38f;
  if (gid < outer * iv) {
    const long long o = gid / iv;
    const long long base = o * n * inner + (gid - o * iv) * V;
    float gscale = 0.f, gtie = 0.f, inv = 1.f;
    if (MODE & 8) {
      const float m = *bw.maxval;
      gscale = 1.f / m;
      gtie = -(bw.sums[0] / m) / bw.sums[1];
    }
    if (MODE & 4) inv = *divisor;
    auto fetch = [&](int q) {
      Vec<V> r;
      if (q < n) {
        r.load(in + base + (long long)q * inner);
        if (MODE & 8) r.bwd(gscale, gtie, bw.out + base + (long long)q * inner);
      } else {
        r.zero();
      }
      return r;
    };
    Vec<V> win[kTaps];                                     // rows p-3 .. p+3
#pragma unroll
    for (int d = 0; d < kR; ++d) win[d].zero();
#pragma unroll
    for (int d = kR; d < kTaps; ++d) win[d] = fetch(d - kR);
    for (int p = 0; p < n; ++p) {
      const Vec<V> next = fetch(p + kR + 1);              // issued before the arithmetic of this row
      Vec<V> acc;
      acc.zero();
#pragma unroll
      for (int d = 0; d < kTaps; ++d) acc.fma(coef[p * kTaps + d], win[d]);
      if (MODE & 4) acc.div(inv);
      if (MODE & 2) lmax = fmaxf(lmax, acc.hmax());
      if (MODE & 1) acc.store(out + base + (long long)p * inner);
#pragma unroll
      for (int d = 0; d < kTaps - 1; ++d) win[d] = win[d + 1];
      win[kTaps - 1] = next;
    }
  }
  if (MODE & 2) {
    lmax = warp_max(lmax);
    if ((threadIdx.x & 31) == 0) atomic_max_float(gmax, lmax);
  }
}

template <int MODE, int R>
__global__ void __launch_bounds__(1024) axis_tile_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                         long long outer, int n, int inner, int o_chunk,
                                                         const Taps tp, float* gmax,
                                                         const float* divisor, BwdIn bw) {
  constexpr int kR = R, kTaps = 2 * R + 1;
  extern __shared__ __align__(16) float tile[];            // [o_chunk][n][inner], contiguous like global memory
  const int ni = n * inner;
  const long long o0 = (long long)blockIdx.x * o_chunk;
  const int no = (int)min((long long)o_chunk, outer - o0);
  const long long g0 = o0 * ni;
  const int telems = no * ni;
  float gscale = 0.f, gtie = 0.f, inv = 1.f;
  if (MODE & 8) {
    const float m = *bw.maxval;
    gscale = 1.f / m;
    gtie = -(bw.sums[0] / m) / bw.sums[1];
  }
  if (MODE & 4) inv = *divisor;
  if ((ni & 3) == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0 &&
      (!(MODE & 8) || (reinterpret_cast<uintptr_t>(bw.out) & 15) == 0)) {
    for (int e = threadIdx.x * 4; e < telems; e += blockDim.x * 4) {
      Vec<4> r;
      r.load(in + g0 + e);
      if (MODE & 8) r.bwd(gscale, gtie, bw.out + g0 + e);
      r.store(tile + e);
    }
  } else {
    for (int e = threadIdx.x; e < telems; e += blockDim.x) {
      Vec<1> r;
      r.load(in + g0 + e);
      if (MODE & 8) r.bwd(gscale, gtie, bw.out + g0 + e);
      tile[e] = r.v;
    }
  }
  // thread t = (p, i): its 7 coefficients and (clamped) row offsets stay in registers
  const int t = threadIdx.x;
  float c[kTaps];
  int off[kTaps];
  {
    const int p = min(t, ni - 1) / inner;
#pragma unroll
    for (int d = 0; d < kTaps; ++d) {
      const int q = p + d - kR;
      const bool ok = q >= 0 && q < n;
      c[d] = ok ? band_coef<R>(tp, n, p, q, (MODE & 16) != 0) : 0.f;
      off[d] = ok ? (d - kR) * inner : 0;
    }
  }
  __syncthreads();
  float lmax = -3.0e38f;
  if (t < ni) {
    for (int oo = 0; oo < no; ++oo) {
      const float* row = tile + oo * ni + t;
      float acc = 0.f;
#pragma unroll
      for (int d = 0; d < kTaps; ++d) acc = fmaf(c[d], row[off[d]], acc);
      if (MODE & 4) acc = acc / inv;
      if (MODE & 2) lmax = fmaxf(lmax, acc);
      if (MODE & 1) out[g0 + (long long)oo * ni + t] = acc;
    }
  }
  if (MODE & 2) {
    lmax = warp_max(lmax);
    if ((threadIdx.x & 31) == 0) atomic_max_float(gmax, lmax);
  }
}

__global__ void init_max_kernel(float* m) { *m = __int_as_float(0xff800000); }

// sums[0] = sum gout * out, sums[1] = number of elements at the maximum (out == 1).  16-byte loads, four in flight per
// thread (the scalar grid-stride version ran at 2.5 TB/s: 29.9 us for the 75.5 MB of config 3).
__global__ void __launch_bounds__(FT) tie_sums_kernel(const float* __restrict__ gout, const float* __restrict__ out,
                                                      long long n, float* sums) {
  float s = 0.f, c = 0.f;
  const long long stride = (long long)gridDim.x * FT;
  const long long tid = (long long)blockIdx.x * FT + threadIdx.x;
  const bool vec = ((reinterpret_cast<uintptr_t>(gout) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
  const long long n4 = vec ? n / 4 : 0;
  const float4* g4 = reinterpret_cast<const float4*>(gout);
  const float4* o4 = reinterpret_cast<const float4*>(out);
  auto acc4 = [&](const float4& g, const float4& o) {
    s = fmaf(g.x, o.x, s); s = fmaf(g.y, o.y, s); s = fmaf(g.z, o.z, s); s = fmaf(g.w, o.w, s);
    c += (o.x == 1.0f ? 1.f : 0.f) + (o.y == 1.0f ? 1.f : 0.f) + (o.z == 1.0f ? 1.f : 0.f) + (o.w == 1.0f ? 1.f : 0.f);
  };
  long long i = tid;
  for (; i + 3 * stride < n4; i += 4 * stride) {
    float4 g[4], o[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) { g[u] = __ldg(g4 + i + u * stride); o[u] = __ldg(o4 + i + u * stride); }
#pragma unroll
    for (int u = 0; u < 4; ++u) acc4(g[u], o[u]);
  }
  for (; i < n4; i += stride) acc4(__ldg(g4 + i), __ldg(o4 + i));
  for (long long k = 4 * n4 + tid; k < n; k += stride) {
    const float o = out[k];
    s = fmaf(gout[k], o, s);
    c += (o == 1.0f) ? 1.f : 0.f;
  }
  // one atomic pair per CTA: same-address atomics retire one at a time (19 000 of them were the kernel: 40 us)
  __shared__ float red[2][FT / 32];
  s = warp_sum(s);
  c = warp_sum(c);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s; red[1][threadIdx.x >> 5] = c; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float ss = 0.f, cc = 0.f;
#pragma unroll
    for (int w = 0; w < FT / 32; ++w) { ss += red[0][w]; cc += red[1][w]; }
    atomicAdd(&sums[0], ss);
    atomicAdd(&sums[1], cc);
  }
}

struct AxisPlan { long long outer, inner; int n, o_chunk; bool tiled; };
AxisPlan plan_axis(long long outer, int n, long long inner) {
  AxisPlan p;
  p.outer = outer; p.inner = inner; p.n = n;
  p.tiled = (inner < 64) && ((long long)n * inner <= 1024);
  p.o_chunk = (int)max(1LL, (long long)kTileElems / ((long long)n * inner));
  return p;
}

template <int MODE, int R>
int run_axis_r(const float* in, float* out, const AxisPlan& p, const Taps& tp, float* gmax, const float* divisor,
               BwdIn bw, cudaStream_t st) {
  if (p.tiled) {
    const int ni = p.n * (int)p.inner;
    const int threads = (ni + 31) / 32 * 32;
    const unsigned grid = (unsigned)((p.outer + p.o_chunk - 1) / p.o_chunk);
    const size_t smem = (size_t)p.o_chunk * ni * sizeof(float);
    axis_tile_kernel<MODE, R><<<grid, threads, smem, st>>>(in, out, p.outer, p.n, (int)p.inner, p.o_chunk, tp, gmax, divisor, bw);
  } else {
    const bool v4 = (p.inner % 4 == 0) && (reinterpret_cast<uintptr_t>(in) & 15) == 0 &&
                    (!(MODE & 1) || (reinterpret_cast<uintptr_t>(out) & 15) == 0) &&
                    (!(MODE & 8) || (reinterpret_cast<uintptr_t>(bw.out) & 15) == 0);
    const long long total = p.outer * (p.inner / (v4 ? 4 : 1));
    const unsigned grid = (unsigned)((total + FT - 1) / FT);
    const size_t smem = (size_t)p.n * (2 * R + 1) * sizeof(float);       // <= 1024 * 13 * 4 = 52 KB
    if (smem > 48 * 1024) {
      if (v4) KCCOT_CUDA(cudaFuncSetAttribute(axis_col_kernel<MODE, 4, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      else KCCOT_CUDA(cudaFuncSetAttribute(axis_col_kernel<MODE, 1, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    if (v4) axis_col_kernel<MODE, 4, R><<<grid, FT, smem, st>>>(in, out, p.outer, p.n, p.inner, tp, gmax, divisor, bw);
    else axis_col_kernel<MODE, 1, R><<<grid, FT, smem, st>>>(in, out, p.outer, p.n, p.inner, tp, gmax, divisor, bw);
  }
  KCCOT_LAUNCH_CHECK();
  return KCCOT_OK;
}
template <int MODE>
int run_axis(const float* in, float* out, const AxisPlan& p, const Taps& tp, float* gmax, const float* divisor,
             BwdIn bw, cudaStream_t st) {
  switch (tp.r) {
    case 1: return run_axis_r<MODE, 1>(in, out, p, tp, gmax, divisor, bw, st);
    case 2: return run_axis_r<MODE, 2>(in, out, p, tp, gmax, divisor, bw, st);
    case 3: return run_axis_r<MODE, 3>(in, out, p, tp, gmax, divisor, bw, st);
    case 4: return run_axis_r<MODE, 4>(in, out, p, tp, gmax, divisor, bw, st);
    case 5: return run_axis_r<MODE, 5>(in, out, p, tp, gmax, divisor, bw, st);
    case 6: return run_axis_r<MODE, 6>(in, out, p, tp, gmax, divisor, bw, st);
  }
  set_error("smoothing radius %d not supported (1..%d)", tp.r, kMaxRadius);
  return KCCOT_EUNSUPPORTED;
}
int make_taps(Taps* tp, const float* w, int radius, const char* what) {
  KCCOT_CHECK_ARG(w != nullptr, "%s: null weights", what);
  if (radius < 1 || radius > kMaxRadius) {
    set_error("%s: smoothing radius %d not supported (kernel sizes 2..%d, radius 1..%d)", what, radius, 2 * kMaxRadius + 1,
              kMaxRadius);
    return KCCOT_EUNSUPPORTED;
  }
  tp->r = radius;
  for (int k = 0; k < 2 * kMaxRadius + 1; ++k) tp->w[k] = k < 2 * radius + 1 ? w[k] : 0.f;
  return KCCOT_OK;
}
}  // namespace
}  // namespace kccot

using namespace kccot;

extern "C" {

size_t kccot_smooth_workspace_bytes(int mode, int B, int H, int T, int W, int C) {
  const size_t n = (size_t)B * H * T * W * C * sizeof(float);
  return align_up(256 + (mode == 3 ? 2 * align_up(n, 256) : 0), 256);
}

int kccot_smooth_fwd(int mode, const float* x, int B, int H, int T, int W, int C, const float* taps_t, int radius_t,
                     const float* taps_s, int radius_s, float* out, float* maxval, void* ws, size_t ws_bytes, void* stream) {
  KCCOT_CHECK_ARG(mode == 1 || mode == 3, "smoothing mode must be 1 (temporal) or 3 (3-D); the reference's '2d' "
                                          "branch raises (data_utils.py:537-538)");
  KCCOT_CHECK_ARG(x && out && maxval, "null pointer");
  KCCOT_CHECK_ARG(B >= 1 && H >= 1 && T >= 1 && W >= 1 && C >= 1, "bad sizes");
  KCCOT_CHECK_ARG(T <= kMaxAxis && H <= kMaxAxis && W <= kMaxAxis, "axis longer than %d not supported (H=%d T=%d W=%d)", kMaxAxis,
                  H, T, W);
  cudaStream_t st = (cudaStream_t)stream;
  Taps tp;
  if (int rc = mode == 1 ? make_taps(&tp, taps_t, radius_t, "temporal_convolution") : make_taps(&tp, taps_s, radius_s,
                                                                                             "gaussian_convolution3D"))
    return rc;
  KCCOT_CHECK_ARG(T > tp.r, "REFLECT padding by %d needs T > %d (T=%d)", tp.r, tp.r, T);
  init_max_kernel<<<1, 1, 0, st>>>(maxval);
  KCCOT_LAUNCH_CHECK();
  const AxisPlan pt = plan_axis((long long)B * H, T, (long long)W * C);
  BwdIn none{};
  if (mode == 1) {
    if (int rc = run_axis<2>(x, nullptr, pt, tp, maxval, nullptr, none, st)) return rc;
    return run_axis<1 | 4>(x, out, pt, tp, nullptr, maxval, none, st);
  }
  KCCOT_CHECK_ARG(H > tp.r && W > tp.r, "3-D smoothing needs H, W > %d (H=%d W=%d)", tp.r, H, W);
  const size_t n = align_up((size_t)B * H * T * W * C * sizeof(float), 256);
  KCCOT_CHECK_ARG(ws && ws_bytes >= 256 + 2 * n, "workspace too small");
  float* t1 = (float*)((char*)ws + 256);
  float* t2 = (float*)((char*)ws + 256 + n);
  const AxisPlan ph = plan_axis(B, H, (long long)T * W * C);
  const AxisPlan pw = plan_axis((long long)B * H * T, W, C);
  if (int rc = run_axis<1>(x, t1, ph, tp, nullptr, nullptr, none, st)) return rc;
  if (int rc = run_axis<1>(t1, t2, pt, tp, nullptr, nullptr, none, st)) return rc;
  if (int rc = run_axis<2>(t2, nullptr, pw, tp, maxval, nullptr, none, st)) return rc;
  return run_axis<1 | 4>(t2, out, pw, tp, nullptr, maxval, none, st);
}

int kccot_smooth_bwd(int mode, const float* gout, const float* out, const float* maxval, int B, int H, int T, int W,
                     int C, const float* taps_t, int radius_t, const float* taps_s, int radius_s, float* gx, void* ws,
                     size_t ws_bytes, void* stream) {
  KCCOT_CHECK_ARG(mode == 1 || mode == 3, "smoothing mode must be 1 or 3");
  KCCOT_CHECK_ARG(gout && out && maxval && gx && ws && ws_bytes >= 256, "null pointer / workspace");
  cudaStream_t st = (cudaStream_t)stream;
  Taps tp;
  if (int rc = mode == 1 ? make_taps(&tp, taps_t, radius_t, "temporal_convolution") : make_taps(&tp, taps_s, radius_s,
                                                                                             "gaussian_convolution3D"))
    return rc;
  const long long nel = (long long)B * H * T * W * C;
  float* sums = (float*)ws;
  KCCOT_CUDA(cudaMemsetAsync(sums, 0, 2 * sizeof(float), st));
  tie_sums_kernel<<<(unsigned)min((long long)4 * num_sms(), (nel / 4 + FT - 1) / FT + 1), FT, 0, st>>>(gout, out, nel, sums);
  KCCOT_LAUNCH_CHECK();
  BwdIn bw{out, maxval, sums};
  BwdIn none{};
  const AxisPlan pt = plan_axis((long long)B * H, T, (long long)W * C);
  if (mode == 1) return run_axis<1 | 8 | 16>(gout, gx, pt, tp, nullptr, nullptr, bw, st);
  const size_t n = align_up((size_t)nel * sizeof(float), 256);
  KCCOT_CHECK_ARG(ws_bytes >= 256 + 2 * n, "workspace too small");
  float* t1 = (float*)((char*)ws + 256);
  float* t2 = (float*)((char*)ws + 256 + n);
  const AxisPlan ph = plan_axis(B, H, (long long)T * W * C);
  const AxisPlan pw = plan_axis((long long)B * H * T, W, C);
  if (int rc = run_axis<1 | 8 | 16>(gout, t1, pw, tp, nullptr, nullptr, bw, st)) return rc;
  if (int rc = run_axis<1 | 16>(t1, t2, pt, tp, nullptr, nullptr, none, st)) return rc;
  return run_axis<1 | 16>(t2, gx, ph, tp, nullptr, nullptr, none, st);
}

}  // extern "C"
