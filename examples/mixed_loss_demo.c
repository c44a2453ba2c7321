/* The C ABI without PyTorch: plain C + the CUDA runtime for memory.
 *
 *   gcc -O2 -std=c99 -Iinclude -I/usr/local/cuda/include examples/mixed_loss_demo.c \
 *       -Lkccotgan_b200 -lkccot -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/kccotgan_b200 -o mixed_loss_demo
 *   ./mixed_loss_demo [B T H W C]
 *
 * Fills real / fake / h / M with a fixed linear congruential sequence, evaluates
 * compute_sinkhorn_loss (gan_utils.py:204-227) forward and backward through kccot_mixed_loss_fwd / _bwd and
 * prints the loss, its three terms and a checksum of every gradient (tests/test_gpu_parity.py repeats the
 * same sequence in numpy and checks the numbers against the Python host path and the fp64 oracle).
 * Exit code 2 if there is no usable device: the library has no CPU fallback. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <cuda_runtime_api.h>

#include "kccot.h"

static uint32_t lcg_state = 12345u;
static float lcg_uniform(void) { /* 24 random bits -> [0, 1) */
  lcg_state = lcg_state * 1664525u + 1013904223u;
  return (float)(lcg_state >> 8) * (1.0f / 16777216.0f);
}

#define CK(expr)                                                                          \
  do {                                                                                    \
    cudaError_t e_ = (expr);                                                              \
    if (e_ != cudaSuccess) {                                                              \
      fprintf(stderr, "%s: %s\n", #expr, cudaGetErrorString(e_));                         \
      return 2;                                                                           \
    }                                                                                     \
  } while (0)
#define KC(expr)                                                                          \
  do {                                                                                    \
    int rc_ = (expr);                                                                     \
    if (rc_ != KCCOT_OK) {                                                                \
      fprintf(stderr, "%s -> %d: %s\n", #expr, rc_, kccot_last_error());                  \
      return rc_ == KCCOT_EINVAL ? 3 : 2;                                                 \
    }                                                                                     \
  } while (0)

static float* upload(size_t n, float lo, float hi) {
  float* h = (float*)malloc(n * sizeof(float));
  float* d = NULL;
  for (size_t i = 0; i < n; ++i) h[i] = lo + (hi - lo) * lcg_uniform();
  if (cudaMalloc((void**)&d, n * sizeof(float)) != cudaSuccess) { free(h); return NULL; }
  cudaMemcpy(d, h, n * sizeof(float), cudaMemcpyHostToDevice);
  free(h);
  return d;
}

static double checksum(const float* d, size_t n) { /* sum_i g_i * (1 + (i mod 7)) on the host */
  float* h = (float*)malloc(n * sizeof(float));
  double s = 0.0;
  cudaMemcpy(h, d, n * sizeof(float), cudaMemcpyDeviceToHost);
  for (size_t i = 0; i < n; ++i) s += (double)h[i] * (double)(1 + (i % 7));
  free(h);
  return s;
}

int main(int argc, char** argv) {
  int B = 16, T = 6, H = 8, W = 8, C = 3;
  const int J = 8, L = 100;
  const float s = 1.0f / 15.0f, eps = 1.0f;
  if (argc == 6) { B = atoi(argv[1]); T = atoi(argv[2]); H = atoi(argv[3]); W = atoi(argv[4]); C = atoi(argv[5]); }
  if (kccot_device_check() != KCCOT_OK) {
    fprintf(stderr, "kccot_device_check: %s\n", kccot_last_error());
    return 2;
  }
  const long long K = (long long)T * H * W * C;
  const size_t nv = (size_t)B * K, nh = (size_t)B * T * J;
  float* real = upload(nv, 0.f, 1.f);
  float* fake = upload(nv, 0.f, 1.f);
  float* hm[4]; /* h_fake, m_real, h_real, m_fake in (0.1, 0.9) */
  for (int i = 0; i < 4; ++i) hm[i] = upload(nh, 0.1f, 0.9f);
  if (!real || !fake || !hm[0] || !hm[1] || !hm[2] || !hm[3]) { fprintf(stderr, "cudaMalloc failed\n"); return 2; }

  void *saved = NULL, *ws = NULL;
  float *out = NULL, *gl = NULL, *g[6];
  const size_t saved_bytes = kccot_mixed_loss_saved_bytes(1, B, L);
  const size_t ws_bytes = kccot_mixed_loss_workspace_bytes(1, B, K, L);
  CK(cudaMalloc(&saved, saved_bytes));
  CK(cudaMalloc(&ws, ws_bytes));
  CK(cudaMalloc((void**)&out, 4 * sizeof(float)));
  CK(cudaMalloc((void**)&gl, sizeof(float)));
  for (int i = 0; i < 6; ++i) CK(cudaMalloc((void**)&g[i], (i < 2 ? nv : nh) * sizeof(float)));
  const float one = 1.0f;
  CK(cudaMemcpy(gl, &one, sizeof(float), cudaMemcpyHostToDevice));

  KC(kccot_mixed_loss_fwd(real, fake, 1, B, K, hm[0], hm[1], hm[2], hm[3], T, J, s, eps, L, saved, out, out + 1, ws,
                          ws_bytes, KCCOT_PATH_AUTO, NULL));
  KC(kccot_mixed_loss_bwd(gl, real, fake, 1, B, K, hm[0], hm[1], hm[2], hm[3], T, J, s, eps, L, saved, g[0], g[1], g[2],
                          g[3], g[4], g[5], ws, ws_bytes, KCCOT_PATH_AUTO, NULL));
  CK(cudaDeviceSynchronize());
  float res[4];
  CK(cudaMemcpy(res, out, sizeof(res), cudaMemcpyDeviceToHost));
  printf("loss %.9g xy %.9g xx %.9g yy %.9g\n", res[0], res[1], res[2], res[3]);
  const char* names[6] = {"f_real", "f_fake", "h_fake", "m_real", "h_real", "m_fake"};
  for (int i = 0; i < 6; ++i) printf("grad %s %.9g\n", names[i], checksum(g[i], i < 2 ? nv : nh));
  printf("kernels launched %llu\n", kccot_launch_count());
  return 0;
}
