// Gaussian kernel smoothing of [B,H,T,W,C] videos (data_utils.py:503-521 temporal_convolution,
// :552-582 gaussian_convolution3D): separable REFLECT-padded 7-tap passes along T (mode 1) or
// H, T, W (mode 3), then division by the global maximum of the filtered tensor.
//
// One generic banded "filter along an axis" kernel: the tensor is viewed as [outer, n, inner] and
// out[o,p,i] = sum_q A[p,q] in[o,q,i], with A the dense [n,n] matrix of the padded filter (band
// radius r), built on the host.  Tiles of [o_chunk][n][inner_chunk] go through shared memory so
// that every global access is a contiguous run.  The same kernel, with A transposed, is the adjoint.
#include "common.cuh"

namespace kccot {
namespace {
constexpr int FT = 256;
constexpr int kTileElems = 8192;     // 32 KB of shared memory per tile

__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
  if (v >= 0.f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

struct BwdIn {            // input transform of the first adjoint pass: gz = gout/m + tie * (-S/m) / count
  const float* out;       // forward output (== 1.0f exactly where the max was attained)
  const float* maxval;
  const float* sums;      // [0] = sum(gout*out), [1] = number of tied maxima
};

// MODE bit 0: write `out`; bit 1: reduce the global max into *gmax; bit 2: divide by *divisor;
// bit 3: apply the BwdIn transform on load; bit 4: use A transposed
template <int MODE>
__global__ void __launch_bounds__(FT) axis_filter_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                         long long outer, int n, long long inner, int o_chunk,
                                                         int i_chunk, const float* __restrict__ A, int radius,
                                                         float* gmax, const float* divisor, BwdIn bw) {
  extern __shared__ float sh[];        // A[n*n] | tile[o_chunk*n*i_chunk]
  float* As = sh;
  float* tile = sh + n * n;
  for (int e = threadIdx.x; e < n * n; e += FT) {
    const int p = e / n, q = e % n;
    As[e] = (MODE & 16) ? A[q * n + p] : A[e];
  }
  const long long n_ichunks = (inner + i_chunk - 1) / i_chunk;
  const long long blk = blockIdx.x;
  const long long o0 = (blk / n_ichunks) * o_chunk;
  const long long i0 = (blk % n_ichunks) * i_chunk;
  const int no = (int)min((long long)o_chunk, outer - o0);
  const int ni = (int)min((long long)i_chunk, inner - i0);
  const int telems = no * n * ni;
  float gscale = 0.f, gtie = 0.f;
  if (MODE & 8) {
    const float m = *bw.maxval;
    gscale = 1.f / m;
    gtie = -(bw.sums[0] / m) / bw.sums[1];
  }
  for (int e = threadIdx.x; e < telems; e += FT) {
    const int ii = e % ni, q = (e / ni) % n, oo = e / (ni * n);
    const long long g = ((o0 + oo) * n + q) * inner + i0 + ii;
    float v = in[g];
    if (MODE & 8) v = v * gscale + ((bw.out[g] == 1.0f) ? gtie : 0.f);
    tile[e] = v;
  }
  __syncthreads();
  float inv = 1.f;
  if (MODE & 4) inv = *divisor;
  float lmax = -3.0e38f;
  for (int e = threadIdx.x; e < telems; e += FT) {
    const int ii = e % ni, p = (e / ni) % n, oo = e / (ni * n);
    const int qlo = max(0, p - radius), qhi = min(n - 1, p + radius);
    const float* col = tile + (oo * n) * ni + ii;
    float acc = 0.f;
    for (int q = qlo; q <= qhi; ++q) acc = fmaf(As[p * n + q], col[q * ni], acc);
    if (MODE & 4) acc = acc / inv;
    if (MODE & 2) lmax = fmaxf(lmax, acc);
    if (MODE & 1) out[((o0 + oo) * n + p) * inner + i0 + ii] = acc;
  }
  if (MODE & 2) {
    lmax = warp_max(lmax);
    if ((threadIdx.x & 31) == 0) atomic_max_float(gmax, lmax);
  }
}

__global__ void init_max_kernel(float* m) { *m = __int_as_float(0xff800000); }

__global__ void __launch_bounds__(FT) tie_sums_kernel(const float* __restrict__ gout, const float* __restrict__ out,
                                                      long long n, float* sums) {
  float s = 0.f, c = 0.f;
  for (long long i = (long long)blockIdx.x * FT + threadIdx.x; i < n; i += (long long)gridDim.x * FT) {
    const float o = out[i];
    s = fmaf(gout[i], o, s);
    c += (o == 1.0f) ? 1.f : 0.f;
  }
  s = warp_sum(s);
  c = warp_sum(c);
  if ((threadIdx.x & 31) == 0) { atomicAdd(&sums[0], s); atomicAdd(&sums[1], c); }
}

struct AxisPlan { long long outer, inner; int n, o_chunk, i_chunk; unsigned grid; size_t smem; };
AxisPlan plan_axis(long long outer, int n, long long inner) {
  AxisPlan p;
  p.outer = outer; p.inner = inner; p.n = n;
  if (inner >= 128) { p.i_chunk = 128; p.o_chunk = 1; }
  else { p.i_chunk = (int)inner; p.o_chunk = (int)max(1LL, (long long)kTileElems / ((long long)n * inner)); }
  while ((long long)p.o_chunk * n * p.i_chunk > kTileElems && p.i_chunk > 1) p.i_chunk /= 2;
  const long long nic = (inner + p.i_chunk - 1) / p.i_chunk;
  const long long noc = (outer + p.o_chunk - 1) / p.o_chunk;
  p.grid = (unsigned)(nic * noc);
  p.smem = ((size_t)n * n + (size_t)p.o_chunk * n * p.i_chunk) * sizeof(float);
  return p;
}

template <int MODE>
int run_axis(const float* in, float* out, const AxisPlan& p, const float* A, int radius, float* gmax,
             const float* divisor, BwdIn bw, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    KCCOT_CUDA(cudaFuncSetAttribute(axis_filter_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    attr = true;
  }
  axis_filter_kernel<MODE><<<p.grid, FT, p.smem, st>>>(in, out, p.outer, p.n, p.inner, p.o_chunk, p.i_chunk, A, radius,
                                                      gmax, divisor, bw);
  KCCOT_LAUNCH_CHECK();
  return KCCOT_OK;
}
constexpr int kRadius = 3;    // temporal_kernel_size = spatial_kernel_size = 6 (kernel_train.py:216) -> radius 3
}  // namespace
}  // namespace kccot

using namespace kccot;

extern "C" {

size_t kccot_smooth_workspace_bytes(int mode, int B, int H, int T, int W, int C) {
  const size_t n = (size_t)B * H * T * W * C * sizeof(float);
  return align_up(256 + (mode == 3 ? 2 * align_up(n, 256) : 0), 256);
}

int kccot_smooth_fwd(int mode, const float* x, int B, int H, int T, int W, int C, const float* filt_h,
                     const float* filt_t, const float* filt_w, float* out, float* maxval, void* ws, size_t ws_bytes,
                     void* stream) {
  KCCOT_CHECK_ARG(mode == 1 || mode == 3, "smoothing mode must be 1 (temporal) or 3 (3-D); the reference's '2d' "
                                          "branch raises (data_utils.py:537-538)");
  KCCOT_CHECK_ARG(x && out && maxval && filt_t, "null pointer");
  KCCOT_CHECK_ARG(B >= 1 && H >= 1 && W >= 1 && C >= 1 && T > kRadius, "REFLECT padding needs T > 3 (T=%d)", T);
  KCCOT_CHECK_ARG(T <= 64 && H <= 64 && W <= 64, "axis longer than 64 not supported (H=%d T=%d W=%d)", H, T, W);
  cudaStream_t st = (cudaStream_t)stream;
  init_max_kernel<<<1, 1, 0, st>>>(maxval);
  KCCOT_LAUNCH_CHECK();
  const AxisPlan pt = plan_axis((long long)B * H, T, (long long)W * C);
  BwdIn none{};
  if (mode == 1) {
    if (int rc = run_axis<2>(x, nullptr, pt, filt_t, kRadius, maxval, nullptr, none, st)) return rc;
    return run_axis<1 | 4>(x, out, pt, filt_t, kRadius, nullptr, maxval, none, st);
  }
  KCCOT_CHECK_ARG(filt_h && filt_w && H > kRadius && W > kRadius, "3-D smoothing needs H, W > 3 and all three filters");
  const size_t n = align_up((size_t)B * H * T * W * C * sizeof(float), 256);
  KCCOT_CHECK_ARG(ws && ws_bytes >= 256 + 2 * n, "workspace too small");
  float* t1 = (float*)((char*)ws + 256);
  float* t2 = (float*)((char*)ws + 256 + n);
  const AxisPlan ph = plan_axis(B, H, (long long)T * W * C);
  const AxisPlan pw = plan_axis((long long)B * H * T, W, C);
  if (int rc = run_axis<1>(x, t1, ph, filt_h, kRadius, nullptr, nullptr, none, st)) return rc;
  if (int rc = run_axis<1>(t1, t2, pt, filt_t, kRadius, nullptr, nullptr, none, st)) return rc;
  if (int rc = run_axis<2>(t2, nullptr, pw, filt_w, kRadius, maxval, nullptr, none, st)) return rc;
  return run_axis<1 | 4>(t2, out, pw, filt_w, kRadius, nullptr, maxval, none, st);
}

int kccot_smooth_bwd(int mode, const float* gout, const float* out, const float* maxval, int B, int H, int T, int W,
                     int C, const float* filt_h, const float* filt_t, const float* filt_w, float* gx, void* ws,
                     size_t ws_bytes, void* stream) {
  KCCOT_CHECK_ARG(mode == 1 || mode == 3, "smoothing mode must be 1 or 3");
  KCCOT_CHECK_ARG(gout && out && maxval && gx && filt_t && ws && ws_bytes >= 256, "null pointer / workspace");
  cudaStream_t st = (cudaStream_t)stream;
  const long long nel = (long long)B * H * T * W * C;
  float* sums = (float*)ws;
  KCCOT_CUDA(cudaMemsetAsync(sums, 0, 2 * sizeof(float), st));
  tie_sums_kernel<<<(unsigned)min((long long)4 * num_sms(), (nel + FT - 1) / FT), FT, 0, st>>>(gout, out, nel, sums);
  KCCOT_LAUNCH_CHECK();
  BwdIn bw{out, maxval, sums};
  BwdIn none{};
  const AxisPlan pt = plan_axis((long long)B * H, T, (long long)W * C);
  if (mode == 1) return run_axis<1 | 8 | 16>(gout, gx, pt, filt_t, kRadius, nullptr, nullptr, bw, st);
  const size_t n = align_up((size_t)nel * sizeof(float), 256);
  KCCOT_CHECK_ARG(filt_h && filt_w && ws_bytes >= 256 + 2 * n, "workspace too small");
  float* t1 = (float*)((char*)ws + 256);
  float* t2 = (float*)((char*)ws + 256 + n);
  const AxisPlan ph = plan_axis(B, H, (long long)T * W * C);
  const AxisPlan pw = plan_axis((long long)B * H * T, W, C);
  if (int rc = run_axis<1 | 8 | 16>(gout, t1, pw, filt_w, kRadius, nullptr, nullptr, bw, st)) return rc;
  if (int rc = run_axis<1 | 16>(t1, t2, pt, filt_t, kRadius, nullptr, nullptr, none, st)) return rc;
  return run_axis<1 | 16>(t2, gx, ph, filt_h, kRadius, nullptr, nullptr, none, st);
}

}  // extern "C"
