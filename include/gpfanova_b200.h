/*
 * gpfanova_b200 -- C ABI of the B200-native Gibbs-sampling hot path of ptonner/gpfanova.
 *
 * The reference (pure Python) has no FFI layer; its hot path sits behind a Python class
 * protocol.  This header is the boundary a maintainer would bind with
 * ctypes; every entry point names the reference code it replaces
 * (paths relative to the reference repository root).
 *
 * Conventions
 *   - all arithmetic is IEEE double; matrices are row-major; "log10" hyper-parameters are
 *     the values the reference keeps in parameter_cache (kernel.pyx:35-37 applies 10**).
 *   - pointers named d_* are DEVICE pointers (e.g. torch tensor.data_ptr()); h_* are HOST.
 *   - every call returns 0 on success, >0 on error (gpf_last_error gives the text);
 *     per-chain numerical status (jitter retries / not-PD) is reported through
 *     gpf_get_status so the host can raise numpy.linalg.LinAlgError like linalg.py:54.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *   - one gpf_ctx per GPU; a ctx is not thread-safe; different ctxs are independent.
 *   - there is NO CPU fallback: without a CUDA device every compute call fails with
 *     GPF_ERR_CUDA.
 */
#ifndef GPFANOVA_B200_H
#define GPFANOVA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GPF_OK 0
#define GPF_ERR_ARG 1
#define GPF_ERR_CUDA 2
#define GPF_ERR_UNSUPPORTED 3
#define GPF_ERR_NOT_PD 4

/* per-matrix status words written by the factorisation kernels */
#define GPF_STATUS_OK 0            /* >0: number of jitter retries that were needed (linalg.py:44-51) */
#define GPF_STATUS_NONPOS_DIAG -1  /* linalg.py:42-43  "not pd: non-positive diagonal elements"      */
#define GPF_STATUS_NOT_PD -2       /* linalg.py:54     "not positive definite, even with jitter."    */

#define GPF_MAX_N 512              /* time points per function (tile engine limit, 16 tiles of 32)   */

typedef struct gpf_ctx gpf_ctx;

int gpf_version(void);
const char* gpf_build_info(void);

/* ------------------------------------------------------------------------------------------
 * Stateless batched operators (kernel objects and linalg of the reference)
 * ---------------------------------------------------------------------------------------- */

/* RBF.dist + RBF._K, kernel/rbf.pyx:8-19, with Kernel.build_params' 10** (kernel/kernel.pyx:16-44).
 * d_x: n x p.  d_log10_sigma, d_log10_ls: batch values.  d_K: batch x n x n (full, symmetric).
 * K[b] = 10^s_b * exp(-0.5 * r2),  r2 = clip(-2 X X^T + |X_i|^2 + |X_j|^2, 0),  X = x / 10^l_b. */
int gpf_rbf_K(const double* d_x, int n, int p, const double* d_log10_sigma, const double* d_log10_ls,
              int batch, double* d_K, void* stream);

/* linalg.jitchol, linalg.py:35-54: lower Cholesky with the jitter-retry ladder, per matrix.
 * d_A: batch x n x n (lower triangle read).  d_L: batch x n x n (lower, upper zeroed).
 * d_logdet (may be NULL): log det(A + jitter I) = 2 sum log L_ii.  d_status: batch ints (GPF_STATUS_*). */
int gpf_jitchol(const double* d_A, int n, int batch, int maxtries, double* d_L, double* d_logdet,
                int* d_status, void* stream);

/* Kernel.K_inv, kernel/kernel.pyx:51-64: (A + offset*I)^-1 = L^-T L^-1, L = jitchol(A + offset I).
 * d_A, d_Ainv: batch x n x n full matrices. */
int gpf_kinv(const double* d_A, int n, int batch, double offset, double* d_Ainv, int* d_status, void* stream);

/* RBF.dist, kernel/rbf.pyx:8-14: scaled distances sqrt(clip(-2 X X^T + |X_i|^2 + |X_j|^2, 0)), X = x / 10^l_b.
 * d_out: batch x n x n. */
int gpf_rbf_dist(const double* d_x, int n, int p, const double* d_log10_ls, int batch, double* d_out, void* stream);

/* RBF._dK, kernel/rbf.pyx:21-38 (used by FunctionDerivative, sample/function.py:43-66), with the reference's own
 * scaling: diff = x_i - x_j on the first input dimension;  cross: (-1/l) diff K;  else (1/l)(1 - diff^2 / l) K.
 * d_out: batch x n x n. */
int gpf_rbf_dK(const double* d_x, int n, int p, const double* d_log10_sigma, const double* d_log10_ls, int cross, int batch,
               double* d_out, void* stream);

/* sum over nf functions of log N(F_f; 0, K + (add_rel mean(K) + add_abs) I) for a batch of (sigma, lengthscale) pairs, each
 * with its own functions d_F (batch x nf x n); Cholesky with linalg.jitchol's ladder (maxtries), log-determinant from the
 * pivots, quadratic forms from the rows riding along.  Replaces the multivariate-normal log-pdf of prior.py:86-99
 * (add_rel = 0: cov = L L^T of jitchol(K)) and of base.py:222,233-238 (add_rel = 1e-6).
 * d_ll: batch; d_logdet (may be NULL): batch; d_status: batch (GPF_STATUS_*). */
int gpf_gp_loglik(const double* d_x, int n, int p, const double* d_log10_sigma, const double* d_log10_ls, int batch, const double* d_F,
                  int nf, double add_rel, double add_abs, int maxtries, double* d_ll, double* d_logdet, int* d_status, void* stream);

/* nsamp draws L z from N(0, K + add I), L = jitchol(K + add I), for every (sigma, lengthscale) pair of the batch:
 * Base.samplePrior's function draws (base.py:249-254) and Prior.sample (prior.py:67-84), batched.
 * z: d_z_inject (batch x nsamp x n) or, if NULL, Philox4x32-10 keyed by seed with counter (pair, draw, stream_id, batch item).
 * d_out: batch x nsamp x n. */
int gpf_gp_sample(const double* d_x, int n, int p, const double* d_log10_sigma, const double* d_log10_ls, int batch, int nsamp,
                  const double* d_z_inject, unsigned long long seed, unsigned stream_id, double add_rel, double add_abs, int maxtries,
                  double* d_out, int* d_status, void* stream);

/* FunctionDerivative, sample/function.py:43-66, for a batch of (sigma, lengthscale) pairs with nf <= 32 functions each
 * (d_F: batch x nf x n):  mu_f = K_ba K_a^-1 beta_f,  cov = K_b - K_ba K_a^-1 K_ba^T,  draw beta'_f = mu_f + L_cov z_f,
 * K_a^-1 = (K + offset I)^-1 (kernel.pyx:51-64), K_b / K_ba = RBF._dK without / with cross (rbf.pyx:21-38), L_cov =
 * jitchol(cov).  z: d_z_inject (batch x nf x n) or Philox (counter: pair, function, stream_id, batch item).
 * d_dF: batch x nf x n;  d_mu (batch x nf x n) and d_cov (batch x n x n) may be NULL;  d_status: batch. */
int gpf_function_derivative(const double* d_x, int n, int p, const double* d_log10_sigma, const double* d_log10_ls, int batch,
                            const double* d_F, int nf, const double* d_z_inject, unsigned long long seed, unsigned stream_id,
                            double offset, int maxtries, double* d_dF, double* d_mu, double* d_cov, int* d_status, void* stream);

/* y[b] (n x m) = F[b]^T X^T + sd[b] * standard normal noise (base.py:256-257) for a batch of function sets F (batch x f x n);
 * X: m x f design matrix; noise from Philox keyed by seed (counter: pair, 0xffff, stream_id, batch item). */
int gpf_compose_y(const double* d_F, const double* d_X, const double* d_sd, int batch, int n, int m, int f, unsigned long long seed,
                  unsigned stream_id, double* d_y, void* stream);

/* FANOVA's Transform samplers (fanova.py:58-81 -> effectSample :106-115, finitePopulationVarianceSample :117-124) for `rows`
 * stored states at once.  d_F: rows x f x n.  Block b maps its contrast functions in0[b] .. in0[b]+nin[b] to nout[b] effect
 * rows through h_coef (the blocks' nout x nin contrast matrices, row-major, concatenated) and, if h_fpv[b], appends the
 * finite-population variance row diag(a P a^T)/(nout-1).  d_out: rows x Q x n, Q = sum_b (nout[b] + fpv[b]), blocks in order. */
int gpf_fanova_transform(const double* d_F, long long rows, int f, int n, int nblk, const int* h_in0, const int* h_nin, const int* h_nout,
                         const int* h_fpv, const double* h_coef, double* d_out, void* stream);

/* Diagnostics over the stored history d_hist (S rows x C chains x cols, the layout gpf_sweep writes) for the columns
 * col0 .. col0+ncol: split-R^ over the C chains (split != 0: every chain cut in halves), pooled mean and the posterior
 * variance estimate.  d_rhat, d_mean, d_var: ncol each. */
int gpf_chain_diagnostics(const double* d_hist, int S, int C, long long cols, int col0, int ncol, int split, double* d_rhat,
                          double* d_mean, double* d_var, void* stream);

/* ascending sort of every row of d_v (rows x S, S <= 8192), in place: the sorted log densities of ScalarInterval
 * (interval.py:32-35; Chen & Shao HPD threshold = element int(S alpha)). */
int gpf_row_sort(double* d_v, int rows, int S, void* stream);

/* FunctionInterval (interval.py:62-131) for a batch of sample sets d_samples (batch x S x n): mean, population std, and
 * the epsilon found by the reference's doubling + bisection on the fraction of samples inside mean -+ 2 std -+ epsilon.
 * d_mean, d_std: batch x n; d_eps: batch; d_iters: batch (evaluations of the coverage). */
int gpf_function_interval(const double* d_samples, int batch, int S, int n, double alpha, double start, double tol, int maxiter,
                          double* d_mean, double* d_std, double* d_eps, int* d_iters, void* stream);

/* ------------------------------------------------------------------------------------------
 * Sampler context: C chains of one model  y (n x m) = F (n x f) X^T (f x m) + noise
 * replaces Base.__init__ state + SamplerContainer.parameter_cache (base.py:16-86, container.py:8-20)
 * ---------------------------------------------------------------------------------------- */
typedef struct gpf_model_desc {
  int n, p, m, f, P;          /* time points, input dims, observations, functions, prior groups   */
  int chains;                 /* C independent chains held by this ctx                              */
  int device;                 /* CUDA device ordinal                                                */
  int chain_offset;           /* global index of this ctx's first chain (Philox key; multi-GPU)     */
  uint64_t seed;              /* Philox4x32-10 key                                                  */
  const double* h_x;          /* n x p                                                              */
  const double* h_y;          /* n x m, NaN = missing (function.py:23-25, base.py:208-209)          */
  const double* h_X;          /* m x f design matrix (fanova.py:283-310)                            */
  const int* h_group_of_fn;   /* f entries: prior group of each function (fanova.py:266-281)        */
  const double* h_slice_w;    /* 1+2P slice widths   (base.py:52-56,75-82), order = hyper layout    */
  const int* h_slice_m;       /* 1+2P slice step-out limits                                         */
  double offset;              /* K_inv jitter, kernel.pyx:7 (1e-9)                                  */
  double prior_jitter;        /* prior_likelihood jitter factor, base.py:222 (1e-6)                 */
  double prior_lb, prior_ub;  /* uniform prior bounds, base.py:57,78,83 (-2, 2)                     */
} gpf_model_desc;

int gpf_create(gpf_ctx** out, const gpf_model_desc* desc);
void gpf_destroy(gpf_ctx* ctx);
const char* gpf_last_error(gpf_ctx* ctx);
void* gpf_stream(gpf_ctx* ctx);   /* the ctx's cudaStream_t */
int gpf_sync(gpf_ctx* ctx);

/* state = parameter_cache of every chain.  hyper layout (log10): [y_sigma, prior0_sigma,
 * prior0_lengthscale, prior1_sigma, ...] (H = 1+2P per chain);  F: chains x f x n. */
int gpf_set_state(gpf_ctx* ctx, const double* h_F, const double* h_hyper);
int gpf_get_state(gpf_ctx* ctx, double* h_F, double* h_hyper);
/* asynchronous device -> host copy queued on the ctx stream behind the kernels already launched (history rows to
 * pinned host memory); complete after gpf_sync. */
int gpf_memcpy_d2h(gpf_ctx* ctx, void* h_dst, const void* d_src, size_t bytes);
/* four user events on the ctx stream (idx 0..3): record after queued work, block the host until it has completed --
 * lets the caller drain history chunks from pinned staging buffers while the next sweeps run. */
int gpf_record_event(gpf_ctx* ctx, int idx);
int gpf_wait_event(gpf_ctx* ctx, int idx);
/* device -> host copy on the ctx's COPY stream, ordered after everything queued so far on the ctx stream, with user
 * event idx recorded behind it: the kernels of later gpf_sweep calls run beside the transfer (the source must not be
 * overwritten before gpf_wait_event(ctx, idx) has returned). */
int gpf_memcpy_d2h_event(gpf_ctx* ctx, void* h_dst, const void* d_src, size_t bytes, int idx);
double* gpf_device_F(gpf_ctx* ctx);       /* device pointers to the live state (chains x f x n) */
double* gpf_device_hyper(gpf_ctx* ctx);   /* chains x H */
int gpf_invalidate(gpf_ctx* ctx);         /* call after writing through gpf_device_hyper(): cached factors are stale */
int gpf_n_samplers(gpf_ctx* ctx);         /* 1 + f + 2P: length of the non-Transform sampler list (base.py:57-84) */

/* per-chain status of the last factorisations (max |code| since last clear); h_status: chains ints.
 * returns GPF_ERR_NOT_PD if any chain holds a negative code. */
int gpf_get_status(gpf_ctx* ctx, int* h_status, int clear);

/* counters since creation/clear: [0] factorisations, [1] log-density evaluations in slice steps,
 * [2] jitter retries, [3] function draws, [4] inverse (potri) count, [5] slice steps, [6] factorisations of C executed
 * for function draws (fewer than [3] when functions share their conditional precision), [7] factorisations of
 * K_g + offset I (one per prior group, chain and sweep). */
int gpf_get_counters(gpf_ctx* ctx, unsigned long long* h_counters8, int clear);

/* Kernel.K_inv for prior group g of every chain at the current hyper-parameters (kernel.pyx:51-64 with
 * rbf.pyx:16-19 generated on chip).  The call always leaves L_B = jitchol(K_g + offset I) (kernel.pyx:53-56) inside
 * the ctx for gpf_function_step;  g = -1: all groups.
 * d_out (may be NULL; only for g >= 0): chains x n x n, the inverse itself (L_B^-T L_B^-1, kernel.pyx:62-64) -- an
 * inspection operator: the sampling path never forms K_inv (see gpf_function_step). */
int gpf_group_kinv(gpf_ctx* ctx, int g, double* d_out);

/* Function._sample, sample/function.py:14-40 (+ base.functionResidual, base.py:180-190) for function fn of every chain:
 * m_t, n_t, then a draw from N(A^-1 b, A^-1) with A = diag(n_t)/(sigma_y+offset) + K_o^-1, b = m_t/(sigma_y+offset),
 * K_o = K_g + offset I, written into the state.  The conditional is evaluated in observation form, which is the same
 * Gaussian without the inverse of the cond-1e11 matrix K_o (S = diag(n_t/(sigma_y+offset))^1/2, C = I + S K_o S):
 *     mu = A^-1 b = K_o S C^-1 q0 (q0 = S^-1 b),   Sigma = A^-1 = K_o - K_o S C^-1 S K_o,
 *     draw: f = L_B z1,  q = q0 - S f - z2,  beta = L_B (z1 + L_B^T S C^-1 q)       (z1, z2 standard normal).
 * On a uniform grid with an even number (>= 64) of time points L_B is the block factor of the centrosymmetric split
 * (option 3) and z1 lives in the split domain.
 *   sampler_id, sweep : Philox counter words (position in the reference's sampler list; sweep number)
 *   d_z_inject (may be NULL): chains x 2n standard normals [z1 | z2] used instead of Philox
 *   d_mu_out (chains x n), d_LC_out (chains x n x n: lower Cholesky factor of C), d_mt_out, d_nt_out (chains x n):
 *   inspection outputs, each may be NULL.
 * Factors K_o first if the ctx does not hold L_B at the current hyper-parameters. */
int gpf_function_step(gpf_ctx* ctx, int fn, unsigned sampler_id, unsigned sweep, const double* d_z_inject,
                      double* d_mu_out, double* d_LC_out, double* d_mt_out, double* d_nt_out);

/* Base.observationLikelihood, base.py:196-214, at candidate log10 y_sigma per chain. d_cand,d_out: chains */
int gpf_obs_loglik(gpf_ctx* ctx, const double* d_cand, double* d_out);

/* Base.prior_likelihood, base.py:216-243, for group g; which = 0: candidate replaces sigma,
 * 1: candidate replaces lengthscale.  d_cand, d_out: chains. */
int gpf_prior_loglik(gpf_ctx* ctx, int g, int which, const double* d_cand, double* d_out);

/* Slice._sample, sample/slice.pyx:24-62, on hyper-parameter h (index in the hyper layout) of every chain,
 * with the log densities above; new value written into the state.
 *   d_u_inject (may be NULL): chains x n_inject uniforms used instead of Philox (u0 -> level, u1 -> placement,
 *   u2 -> step split, u3.. -> proposals);  d_nevals_out (may be NULL): chains ints, log-density evaluations. */
int gpf_slice_step(gpf_ctx* ctx, int h, unsigned sampler_id, unsigned sweep, const double* d_u_inject,
                   int n_inject, int* d_nevals_out);

/* SamplerContainer.sample/_sample/store, sample/container.py:29-78: n_sweeps sweeps over the sampler list
 * in the reference's fixed order (base.py:57-84), or in h_order (permutation of sampler ids, container.py:34-35:
 * n_sweeps x n_samplers entries) when non-NULL.  Every `thin`-th sweep (thin<=0: every sweep) the state row
 * [hyper(H), F(f*n)] of every chain is appended to d_history (rows x chains x (H + f*n)); returns rows stored
 * through *rows_out.  sweep0 = index of the first sweep (Philox counter word). */
int gpf_sweep(gpf_ctx* ctx, int n_sweeps, int thin, unsigned sweep0, const int* h_order, double* d_history,
              long long history_capacity_rows, long long* rows_out);

/* timing/accounting of the last gpf_sweep call, measured with CUDA events on the ctx stream:
 * h_ms[0] total, [1] K_o factor / K_inv kernels, [2] function kernels, [3] slice kernels, [4] y_sigma kernel, [5] store.
 * gpf_set_profile(ctx, 1): total + per-kernel brackets (adds event overhead, serialises the groups' slice kernels);
 * gpf_set_profile(ctx, 2): total only (one event pair around the whole call); 0: off. */
/* options: 0 = use the O(n^2) Toeplitz (Schur algorithm) slice log-likelihood when the time grid is uniform
 * (default 1; 0 forces the dense tile engine, which is always used for non-uniform inputs);
 * 1 = balanced design with nothing missing (X^T X diagonal): the function draws of a sweep do not depend on each other
 * and run as ONE fused launch per sweep, one (prior group, chain) item per CTA: factor of K_o, then the group's
 * functions, all in the CTA's L2-resident workspace (default 1; 0 = one launch per sampler in the reference's order;
 * the draws agree to rounding);
 * 2 = let functions of a group with identical conditional precision (balanced design, equal (X^T X)[f,f]) share one
 * factorisation of C (default 1; results agree to rounding);
 * 3 = uniform grid with an even number (>= 64) of time points: K_o is symmetric Toeplitz, hence centrosymmetric,
 * K_o = Q^T diag(B+, B-) Q with B+- = K11 +- K12 J and Q = [[I, J], [I, -J]]/sqrt 2: every n^3 operation of the function
 * step (and gpf_group_kinv's inverse) runs on the two n/2 x n/2 blocks -- a quarter of the flops (default 1;
 * 0 = full-size matrices; changes the square root of K_o the prior draw uses, i.e. the map from z1 to the draw);
 * 4 = balanced design with nothing missing: the residual sum of squares of base.py:196-214 is evaluated about the
 * least-squares fit, ||y - X F||^2 = ||y - X bhat||^2 + sum_f (X^T X)[f,f] ||F_f - bhat_f||^2 (O(f n) per chain instead
 * of O(n m f), all terms non-negative; default 1; 0 = per-observation loop, always used otherwise);
 * 5 = uniform grid, balanced design with nothing missing, n <= 256 and at most four functions per prior group: K_o and
 * C = I + c K_o are symmetric positive definite Toeplitz matrices and the whole function step runs on the Schur
 * algorithm, O(n^2) per function, one warp per (prior group, chain) (default 1 where it applies; 0 = dense fused kernel).
 * 7 = where option 5 applies to every prior group and the samplers run in the reference's fixed order: all sweeps of a
 * gpf_sweep call run in ONE launch with no device-wide barrier between samplers or sweeps -- warps take (sweep, group, chain)
 * units from a ticket counter, a chain's y_sigma step and history row are done by whichever warp finishes the chain's last
 * group of a sweep, chains drift apart freely (csrc/gpf_sweepk.cuh; bit-identical to option 7 = 0, which launches one kernel per
 * sampler stage and sweep; per-kernel profiling, gpf_set_profile(ctx, 1), implies 0).  DEFAULT 0: with the function step and
 * three slice variants in flight on every SM the kernel's hot loops miss the 32 KB instruction cache (ncu: 47 % hit rate, 23
 * no-instruction stall cycles per issue) and it runs four times slower than the per-stage kernels.
 * 8 = where option 5 applies to every prior group: every group's functions and slice kernel run back to back on the group's
 * own stream (a group's slices start while other groups still draw functions; default 1; bit-identical to 8 = 0); 8 = 2:
 * additionally the slice kernel that finishes the last group of a chain's sweep stores the chain's history row and runs the
 * y_sigma step of the next sweep (needs option 4), so only the first sweep of a call launches the y_sigma kernel (fewer launches, but measured slower on the device);
 * 9 = 2 additionally splits the chains into two cohorts with their own streams, so that the stage boundaries of one
 * cohort are filled by the other's kernels (default 1 = one cohort: two measured slower at every chain count -- more
 * kernel variants per SM for the instruction cache, smaller grids).
 * Where option 5 applies, option 3 defaults to 0: the prior draw then uses the full Cholesky factor of K_o in every
 * kernel of the context; setting option 3 to 1 selects the split factors and with them the dense kernels.
 * gpf_get_option returns the EFFECTIVE value (option 6, read-only: 1 if the design is balanced with nothing missing). */
int gpf_set_option(gpf_ctx* ctx, int option, int value);
int gpf_get_option(gpf_ctx* ctx, int option);
int gpf_set_profile(gpf_ctx* ctx, int on);
int gpf_get_profile(gpf_ctx* ctx, double* h_ms6, unsigned long long* h_launches);
/* development aid: cycle counters of the phases of the fused function kernel, read and cleared (all zero unless the
 * library was built with -DGPF_FTRACE; tools/fused_trace.py) */
int gpf_debug_trace(gpf_ctx* ctx, unsigned long long* h_cycles24);

#ifdef __cplusplus
}
#endif
#endif
