/*
 * kccot.h — C ABI of libkccot.so: the B200 (sm_100a) causal-OT loss path of KCCOT-GAN.
 *
 * The reference (neuripss2020/kccotgan) has no FFI: its loss path is a set of plain Python
 * functions over TensorFlow eager tensors (gan_utils.py, data_utils.py:478-586).  This header is
 * the boundary a maintainer binds instead: every entry point below replaces the arithmetic of the
 * reference function cited next to it.  The Python mirror of the reference interface
 * (kccotgan_b200/gan_utils.py, data_utils.py) calls these through ctypes; INTEGRATION.md shows
 * the stub for the reference's own TensorFlow training loop.
 *
 * Conventions
 *   - All pointers are DEVICE pointers to contiguous fp32 unless stated; the caller owns every
 *     buffer (inputs, outputs, saved state, workspace).  The library allocates nothing per call
 *     and keeps no pointer after return.
 *   - `stream` is a cudaStream_t passed as void*.  Calls only enqueue work; none synchronises.
 *   - Return 0 on success, a negative KCCOT_E* code otherwise; kccot_last_error() gives the text
 *     (thread-local).  There is no CPU fallback: without a CUDA device every compute call fails.
 *   - Videos are [B, ...] tensors flattened to rows of length K (any order of the trailing axes:
 *     the cost sums over all of them, gan_utils.py:14-17,216-220).  h / M are [B, T, J].
 *   - "nprob" batches independent problems along a leading axis of every tensor.
 */
#ifndef KCCOT_H_
#define KCCOT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KCCOT_VERSION 202
#define KCCOT_SHARD_FLAGS_PER_RANK 256

/* error codes */
#define KCCOT_OK 0
#define KCCOT_EINVAL (-1)   /* bad shape / null pointer / misaligned buffer  -> Python ValueError */
#define KCCOT_ECUDA (-2)    /* CUDA runtime / driver error                   -> Python RuntimeError */
#define KCCOT_EWORKSPACE (-3)
#define KCCOT_EUNSUPPORTED (-4)

/* kernel-path selection flags (shape-based AUTO is what the product uses; the explicit values
 * exist so that tests can check the tensor-core path against the CUDA-core path) */
#define KCCOT_PATH_AUTO 0
#define KCCOT_PATH_SIMT 1      /* fp32 CUDA-core kernels, direct (x-y)^2 form, any shape */
#define KCCOT_PATH_TCGEN05 2   /* TMA + tcgen05 3xTF32 kernels (needs K % 4 == 0, 16-B aligned rows) */
#define KCCOT_FLAG_ACCUMULATE 16  /* gradient outputs: add into the buffer instead of overwriting */

int kccot_version(void);
const char* kccot_last_error(void);
/* 0 if the current device is sm_100 (B200); KCCOT_EUNSUPPORTED otherwise. */
int kccot_device_check(void);
/* number of kernels this library has launched so far in this process (bench.py's gpu_launches) */
unsigned long long kccot_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * Cost matrices — gan_utils.py:6-18 (cost_xy), :21-43 (modified_cost), :46-72 (bi_causal_...)
 *   C[p,i,j] = s * sum_k (x[p,i,k] - y[p,j,k])^2
 *            + s * sum_{t<T-1,c} h1[p,i,t,c] * (M1[p,j,t+1,c] - M1[p,j,t,c])      (if h1 != NULL)
 *            + s * sum_{t<T-1,c} h2[p,i,t,c] * (M2[p,j,t+1,c] - M2[p,j,t,c])      (if h2 != NULL)
 * x [nprob,Bx,K], y [nprob,By,K], h* [nprob,Bx,T,J], M* [nprob,By,T,J], C [nprob,Bx,By].
 * x == y (same pointer) marks a self-cost: the diagonal is exactly 0 as in the reference.
 * ------------------------------------------------------------------------------------------ */
size_t kccot_cost_workspace_bytes(int nprob, int Bx, int By, long long K);
int kccot_cost_fwd(const float* x, const float* y, int nprob, int Bx, int By, long long K,
                   const float* h1, const float* M1, const float* h2, const float* M2, int T, int J,
                   float s, float* C, void* ws, size_t ws_bytes, int flags, void* stream);

/* The three cost matrices of the mixed loss in one pass over the videos —
 * gan_utils.py:221-223: xy = (real,fake,h_fake,m_real), xx = (real,real,h_real,m_real),
 * yy = (fake,fake,h_fake,m_fake).  C3 [nprob,3,B,B] in the order xy, xx, yy. */
size_t kccot_mixed_cost_workspace_bytes(int nprob, int B, long long K);
int kccot_mixed_cost_fwd(const float* real, const float* fake, int nprob, int B, long long K,
                         const float* h_fake, const float* m_real, const float* h_real,
                         const float* m_fake, int T, int J, float s, float* C3, void* ws,
                         size_t ws_bytes, int flags, void* stream);

/* First half of kccot_mixed_cost_fwd on its own: only the split-K partial squared distances of the
 * stacked rows [real; fake] are written to the workspace (the HBM-bound pass over the videos).
 * Exposed so that the streaming kernel can be timed and profiled in isolation. */
int kccot_mixed_sqdist_partials(const float* real, const float* fake, int nprob, int B, long long K,
                                void* ws, size_t ws_bytes, int flags, void* stream);

/* Adjoint of the squared-distance part: gx[p,i,:] = 2s * sum_j Cbar[p,i,j] (x_i - y_j),
 * gy[p,j,:] = 2s * sum_i Cbar[p,i,j] (y_j - x_i).  gx / gy may be NULL (not needed). */
size_t kccot_cost_bwd_workspace_bytes(int nprob, int Bx, int By, long long K);
int kccot_cost_bwd(const float* Cbar, const float* x, const float* y, int nprob, int Bx, int By,
                   long long K, float s, float* gx, float* gy, void* ws, size_t ws_bytes, int flags,
                   void* stream);
/* Adjoint of one martingale term: gh[:, :-1] = s*Cbar@DeltaM, gh[:, -1] = 0; gM from s*Cbar^T@h. */
int kccot_martingale_bwd(const float* Cbar, const float* h, const float* M, int nprob, int Bx, int By,
                         int T, int J, float s, float* gh, float* gM, int flags, void* stream);
/* Adjoint of kccot_mixed_cost_fwd.  Cbar3 [nprob,3,B,B] already carries the 2,-1,-1 weights and
 * the upstream gradient.  g_real may be NULL (data needs no gradient, kernel_train.py:289). */
size_t kccot_mixed_cost_bwd_workspace_bytes(int nprob, int B, long long K);
int kccot_mixed_cost_bwd(const float* Cbar3, const float* real, const float* fake, int nprob, int B,
                         long long K, const float* h_fake, const float* m_real, const float* h_real,
                         const float* m_fake, int T, int J, float s, float* g_real, float* g_fake,
                         float* gh_fake, float* gm_real, float* gh_real, float* gm_fake, void* ws,
                         size_t ws_bytes, int flags, void* stream);

/* ------------------------------------------------------------------------------------------
 * Log-domain Sinkhorn — gan_utils.py:138-165 (compute_sinkhorn), :86-121 (benchmark_sinkhorn)
 * nsolve independent B x B problems.  L iterations of the u / v updates with uniform marginals;
 * after iteration n (1-based) the solve stops if thresh > sum|u - u_prev| and
 * (exit_on_index ? n-1 >= Lmin : n >= Lmin)   [:159 vs :116].  The test runs on the device.
 * Saved for the backward (internal units: potentials * log2(e)/eps, cost shifted by its minimum):
 *   u_hist, v_hist [nsolve, L+1, B];  nits [nsolve] int32;  cost [nsolve] = sum(pi * C).
 * ------------------------------------------------------------------------------------------ */
size_t kccot_sinkhorn_workspace_bytes(int nsolve, int B, int L);
int kccot_sinkhorn_fwd(const float* C, int nsolve, int B, float eps, int L, int Lmin, float thresh,
                       int exit_on_index, float* u_hist, float* v_hist, int32_t* nits, float* cost,
                       void* ws, size_t ws_bytes, void* stream);
/* Reverse mode through the executed iterations (the reference has no stop_gradient; TF's tape
 * unrolls the loop).  Cbar[n] = gcost[n] * d cost[n] / d C[n].  gcost [nsolve] on the device. */
int kccot_sinkhorn_bwd(const float* C, int nsolve, int B, float eps, int L, const float* u_hist,
                       const float* v_hist, const int32_t* nits, const float* gcost, float* Cbar,
                       void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Row-sharded Sinkhorn, host-driven variant (round 1; kept for callers that bring their own exchange): one call
 * per half-iteration pair, the CALLER moves the column statistics between ranks.  The fused variant below
 * (kccot_shard_sinkhorn_*) supersedes it.  This rank owns rows
 * [row0, row0+Brows) of the B x B cost (Crows [Brows,B]).  The u-update is local; the v-update needs a
 * column log-sum-exp over all rows: each rank produces colstat [2,B] = (max, sum exp2(. - max)) over its
 * rows, the CALLER all-gathers them (NCCL over NVLink) and every rank combines.  The reverse pass
 * exchanges plain [B] sums.  Potentials are in the library's internal log2 units; `ws` (same buffer for
 * all calls of one solve) starts with the cost shift: the caller all-reduces (MIN) the float at ws+0
 * after kccot_shard_begin.  One kernel boundary per half-iteration pair is the synchronisation point.
 * ------------------------------------------------------------------------------------------ */
size_t kccot_shard_workspace_bytes(int Brows, int B);
int kccot_shard_begin(const float* Crows, int Brows, int B, void* ws, size_t ws_bytes, void* stream);
int kccot_shard_fwd_rows(const float* Crows, int Brows, int B, float eps, const float* v, float* u_rows,
                         float* colstat, void* ws, size_t ws_bytes, void* stream);
int kccot_shard_fwd_combine(const float* colstat_all, int nranks, int B, float* v, void* ws, void* stream);
int kccot_shard_cost_partial(const float* Crows, int Brows, int B, float eps, const float* u_rows,
                             const float* v, float* partial, void* ws, void* stream);
int kccot_shard_bwd_seed(const float* Crows, int Brows, int B, float eps, const float* u_rows, const float* v,
                         float g, float* Cbar_rows, float* ubar_rows, float* colsum, void* ws, void* stream);
int kccot_shard_bwd_rows(const float* Crows, int Brows, int B, float eps, const float* u_k_rows,
                         const float* v_k, const float* v_km1, const float* vbar, int first,
                         float* ubar_rows, float* Cbar_rows, float* colsum, void* ws, void* stream);

/* ------------------------------------------------------------------------------------------
 * Row-sharded mixed loss, fused exchange (BASELINE config 5 on N GPUs).  Videos, h and M are replicated on
 * every rank (or all-gathered once by the caller); this rank owns the samples [row0, row0 + Brows): it
 * computes the cost ROWS of the xy, xx, yy blocks (C3rows [3,Brows,B]), runs its row block of the three
 * Sinkhorn solves inside ONE persistent kernel that exchanges the [B] column sums with the other ranks through
 * peer-mapped mailboxes (no host-launched collective per iteration), and returns the gradient rows of its
 * own samples.  Collectives left to the caller, once per evaluation: MIN of `shift` [3] after
 * kccot_shard_local_min, SUM of cost_partial [3,2] (cost_n = partial[n,1] * eps / log2(e) + shift[n] *
 * partial[n,0]), all-to-all of the Cbar_xy / Cbar_yy row panels into column panels XYcol, YYcol [B,Brows],
 * SUM of the two M-gradient partials.
 *   mbox_ptrs[r] / flag_ptrs[r]: device pointers (valid on THIS device) to rank r's mailbox
 *   (kccot_shard_mailbox_bytes) and flag array (nranks x KCCOT_SHARD_FLAGS_PER_RANK x uint64, zeroed once: one flag
 *   per source rank and CTA of the persistent kernel).  epoch_base: any number larger
 *   than every epoch used so far, identical on all ranks (e.g. launch counter << 24).
 * ------------------------------------------------------------------------------------------ */
size_t kccot_shard_cost_workspace_bytes(int B, long long K, int Brows);
int kccot_shard_cost_fwd(const float* real, const float* fake, int B, long long K, int row0, int Brows,
                         const float* h_fake, const float* m_real, const float* h_real,
                         const float* m_fake, int T, int J, float s, float* C3rows, void* ws,
                         size_t ws_bytes, void* stream);
/* `ws` must be the workspace of the matching kccot_shard_cost_fwd call (it carries the transposed split). */
int kccot_shard_cost_bwd(const float* Cbar3rows, const float* XYcol, const float* YYcol, int B, long long K,
                         int row0, int Brows, const float* h_fake, const float* m_real,
                         const float* h_real, const float* m_fake, int T, int J, float s,
                         float* g_fake_rows, float* gh_fake_rows, float* gm_real_part,
                         float* gh_real_rows, float* gm_fake_part, void* ws, size_t ws_bytes, void* stream);
size_t kccot_shard_sinkhorn_workspace_bytes(int np, int Brows, int B, int L);
size_t kccot_shard_mailbox_bytes(int np, int nranks, int B);
int kccot_shard_local_min(const float* Crows, int np, int Brows, int B, float* shift_out, void* stream);
int kccot_shard_sinkhorn_fwd(const float* Crows, int np, int Brows, int B, int row0, float eps, int L,
                             int Lmin, float thresh, int exit_on_index, float* u_hist, float* v_hist,
                             int32_t* nits, float* cost_partial, const float* shift, int nranks, int rank,
                             void* const* mbox_ptrs, void* const* flag_ptrs,
                             unsigned long long epoch_base, void* ws, size_t ws_bytes, void* stream);
int kccot_shard_sinkhorn_bwd(const float* Crows, int np, int Brows, int B, int row0, float eps, int L,
                             const float* u_hist, const float* v_hist, const int32_t* nits,
                             const float* gcost, float* Cbar_rows, const float* shift, int nranks, int rank,
                             void* const* mbox_ptrs, void* const* flag_ptrs,
                             unsigned long long epoch_base, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Fused mixed Sinkhorn loss — gan_utils.py:204-227 (compute_sinkhorn_loss) in ONE call per direction:
 *   fwd: stacked squared distances -> three cost matrices -> three Sinkhorn solves -> 2*xy - xx - yy
 *   bwd: three reverse solves -> adjoint GEMMs -> gradients of real / fake / h_fake / m_real / h_real / m_fake
 * `saved` (kccot_mixed_loss_saved_bytes) holds what the backward needs: C3 | u_hist | v_hist | nits | cost.
 * loss [nprob]; terms [nprob,3] = (xy, xx, yy) (may be NULL).  gloss [nprob] on the device.
 * Any gradient pointer may be NULL.  eps / L are explicit here: the reference's Python entry point
 * always passes 1.0 / 100 (its eps/L arguments are swallowed).
 * ------------------------------------------------------------------------------------------ */
size_t kccot_mixed_loss_saved_bytes(int nprob, int B, int L);
size_t kccot_mixed_loss_workspace_bytes(int nprob, int B, long long K, int L);
int kccot_mixed_loss_fwd(const float* real, const float* fake, int nprob, int B, long long K,
                         const float* h_fake, const float* m_real, const float* h_real,
                         const float* m_fake, int T, int J, float s, float eps, int L, void* saved,
                         float* loss, float* terms, void* ws, size_t ws_bytes, int flags, void* stream);
int kccot_mixed_loss_bwd(const float* gloss, const float* real, const float* fake, int nprob, int B,
                         long long K, const float* h_fake, const float* m_real, const float* h_real,
                         const float* m_fake, int T, int J, float s, float eps, int L, const void* saved,
                         float* g_real, float* g_fake, float* gh_fake, float* gm_real, float* gh_real,
                         float* gm_fake, void* ws, size_t ws_bytes, int flags, void* stream);

/* Shared context frames — kernel_train.py:225-226 / :267-268 build  real = concat(real_in, real_pred) and
 * fake = concat(real_in, fake_pred)  along the time axis, so with `--kernel none` the context frames of the two
 * videos are the same numbers.  The caller states that with (ctx_period, ctx_len): every column c of the flattened
 * [B, K] rows with (c % ctx_period) < ctx_len holds identical values in real and fake.  For [B,H,T,W,C] videos
 * ctx_period = T*W*C and ctx_len = ctx_frames*W*C.
 *   fwd: fake's context columns are never read (the real rows stand in for them): results are bit-identical to
 *        kccot_mixed_loss_fwd on inputs that keep the promise.
 *   bwd: g_fake is the gradient with respect to the PREDICTED columns only; its context columns (constants of the
 *        caller: copies of the data) are not read and not computed; they are set to ZERO (left untouched under
 *        KCCOT_FLAG_ACCUMULATE).
 * The tcgen05 path skips the columns when ctx_period and ctx_len are multiples of 32 and K is a multiple of
 * ctx_period; every other path computes the plain result first, so the outcome is the same on all paths.
 * Not valid after temporal / 3-D smoothing (which leaks predicted frames into context frames). */
int kccot_mixed_loss_fwd_ctx(const float* real, const float* fake, int nprob, int B, long long K,
                             const float* h_fake, const float* m_real, const float* h_real,
                             const float* m_fake, int T, int J, float s, float eps, int L, void* saved,
                             float* loss, float* terms, void* ws, size_t ws_bytes, int flags, void* stream,
                             long long ctx_period, long long ctx_len);
int kccot_mixed_loss_bwd_ctx(const float* gloss, const float* real, const float* fake, int nprob, int B,
                             long long K, const float* h_fake, const float* m_real, const float* h_real,
                             const float* m_fake, int T, int J, float s, float eps, int L, const void* saved,
                             float* g_real, float* g_fake, float* gh_fake, float* gm_real, float* gh_real,
                             float* gm_fake, void* ws, size_t ws_bytes, int flags, void* stream,
                             long long ctx_period, long long ctx_len);

/* ------------------------------------------------------------------------------------------
 * Martingale penalty p_M — gan_utils.py:179-201.  M [B,T,J]; pm [1]; gM [B,T,J] = gpm * dpm/dM.
 * ------------------------------------------------------------------------------------------ */
int kccot_pm_fwd(const float* M, int B, int T, int J, float reg_lam, float s, float* pm,
                 float* stats /* [2*J + (T-1)*J] saved: mean_j, std_j, A_tj */, void* stream);
int kccot_pm_bwd(const float* M, int B, int T, int J, float reg_lam, float s, const float* stats,
                 const float* gpm /* device scalar */, float* gM, void* stream);

/* ------------------------------------------------------------------------------------------
 * Gaussian kernel smoothing — data_utils.py:503-521 (temporal_convolution, mode 1) and
 * :552-582 (gaussian_convolution3D, mode 3).  x, out [B,H,T,W,C].  out = conv(x) / max(conv(x)) with REFLECT
 * padding; maxval [1] is saved for the backward.
 * taps_t / taps_s are HOST arrays of 2*radius+1 normalised weights (gaussian_kernel1d, data_utils.py:483-491);
 * they travel as kernel arguments: no filter matrix in device memory and no copy when sigma is annealed every
 * step (:584-586).  Mode 1 filters T with (taps_t, radius_t); mode 3 filters H, T and W with (taps_s, radius_s)
 * (the reference's 3-D kernel uses the spatial size for all three axes, :553,562-564).  radius 1..6, axes up to
 * 1024 and longer than the radius.
 * ------------------------------------------------------------------------------------------ */
size_t kccot_smooth_workspace_bytes(int mode, int B, int H, int T, int W, int C);
int kccot_smooth_fwd(int mode, const float* x, int B, int H, int T, int W, int C, const float* taps_t,
                     int radius_t, const float* taps_s, int radius_s, float* out, float* maxval, void* ws,
                     size_t ws_bytes, void* stream);
int kccot_smooth_bwd(int mode, const float* gout, const float* out, const float* maxval, int B, int H,
                     int T, int W, int C, const float* taps_t, int radius_t, const float* taps_s,
                     int radius_s, float* gx, void* ws, size_t ws_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* KCCOT_H_ */
