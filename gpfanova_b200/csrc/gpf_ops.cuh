// Batched operators beside the sweep: the callers and consumers of the hot path.
//   k_rbf_aux       RBF.dist / RBF._dK                         kernel/rbf.pyx:8-14, 21-38
//   k_gp_loglik     sum_f log N(F_f; 0, K + add I)             prior.py:86-99, base.py:233-238 (MVN part)
//   k_gp_sample     draws from N(0, K) through jitchol(K)      base.py:245-254, prior.py:67-84
//   k_compose_y     y = F X^T + noise                          base.py:256-257
//   k_transform     effect / interaction / FPV transforms      fanova.py:58-81, 106-124
//   k_seq_moments, k_rhat   per-column chain moments and split-R^ over chains
//   k_row_sort      sorted log densities (Chen-Shao HPD)       interval.py:21-37
//   k_band          FunctionInterval                           interval.py:62-131
#pragma once
#include "gpf_common.cuh"

namespace gpf {

// mode 1: RBF.dist (scaled distances); 2: RBF._dK(cross=False); 3: RBF._dK(cross=True).  grid.x = batch, dense output.
__global__ void __launch_bounds__(NTHR) k_rbf_aux(const double* __restrict__ x, int n, int p, const double* __restrict__ lsig,
                                                   const double* __restrict__ lls, double* __restrict__ out, int mode) {
  extern __shared__ __align__(16) double gpf_smem_ops[];
  double* xs = gpf_smem_ops;
  double* xsq = xs + (size_t)n * p;
  const int b = blockIdx.x;
  const double sigma = lsig ? exp10(lsig[b]) : 1.0, ls = exp10(lls[b]);
  rbf_scale_inputs(x, n, p, ls, xs, xsq);
  __syncthreads();
  double* o = out + (size_t)b * n * n;
  for (size_t e = threadIdx.x; e < (size_t)n * n; e += NTHR) {
    const int R = (int)(e / n), Cc = (int)(e % n);
    double dot = 0.0;
    for (int d = 0; d < p; ++d) dot = fma(xs[R * p + d], xs[Cc * p + d], dot);
    const double r2 = fmax(-2.0 * dot + xsq[R] + xsq[Cc], 0.0);      // rbf.pyx:12-13
    const double dist = sqrt(r2);
    if (mode == 1) { o[e] = dist; continue; }
    const double k = sigma * exp(-0.5 * (dist * dist));               // rbf.pyx:19
    const double diff = x[R * p] - x[Cc * p];                         // rbf.pyx:30-32 (first input dimension, unscaled)
    o[e] = (mode == 3) ? (-1.0 / ls) * diff * k : (1.0 / ls) * (1.0 - (1.0 / ls) * (diff * diff)) * k;   // rbf.pyx:34-37
  }
}

struct GpArgs {
  const double* x;
  int n, p, NT, batch, nf, maxtries;
  const double *lsig, *lls;
  const double* F;        // loglik: batch x nf x n function values;  sample: injected normals batch x nf x n or null
  double add_rel;         // diagonal += add_rel * mean(K)            (base.py:222: 1e-6; prior.py: 0)
  double add_abs;         // diagonal += add_abs
  double* out;            // loglik: batch;  sample: batch x nf x n
  double* logdet;         // loglik: batch or null
  int* status;            // batch: jitter retries / GPF_STATUS_*
  double* ws;
  size_t ws_stride;
  unsigned k0, k1, stream_id;
};

// log-likelihood of nf functions under one GP per batch item, with linalg.jitchol's ladder (linalg.py:41-54)
__global__ void __launch_bounds__(NTHR, 1) k_gp_loglik(GpArgs a) {
  const int n = a.n, NT = a.NT, NE = (a.nf + TB - 1) / TB, nti_last = (a.nf - (NE - 1) * TB + 7) / 8;
  EngineSmem sm(smem_base(), NT, NE);
  double* xs = smem_base() + engine_smem_doubles(NT, NE);
  double* xsq = xs + (size_t)n * a.p;
  double* W = a.ws + (size_t)blockIdx.x * a.ws_stride;
  double* Wx = W + (size_t)plt_tiles(NT) * TILE;
  for (int b = blockIdx.x; b < a.batch; b += gridDim.x) {
    const double sigma = exp10(a.lsig[b]), ls = exp10(a.lls[b]);
    int tries = 0, code = 0;
    double jitter = 0.0, addv = 0.0;
    for (;;) {
      __syncthreads();
      rbf_scale_inputs(a.x, n, a.p, ls, xs, xsq);
      __syncthreads();
      const double tot = rbf_fill_tiles(W, NT, n, a.p, xs, xsq, sigma, sm.red());
      addv = a.add_abs + a.add_rel * tot / ((double)n * (double)n);
      for (int t = threadIdx.x; t < NT * TB; t += NTHR) sm.dvec[t] = (t < n) ? addv + jitter : 0.0;
      __syncthreads();
      EngineIO io{W, Wx, RowSrc{a.F + (size_t)b * a.nf * n, nullptr, n, a.nf, n}, nullptr, 0};
      if (chol_engine<AUG_RHS, false>(SrcTiles{W}, io, NT, NE, nti_last, sm)) { code = tries; break; }
      const double dm = sigma + addv;                  // mean of the diagonal
      if (tries == 0) {
        if (!(dm > 0.0)) { code = GPF_STATUS_NONPOS_DIAG; break; }
        jitter = dm * 1e-6;
      } else jitter *= 10.0;
      ++tries;
      if (tries > a.maxtries || !isfinite(jitter)) { code = GPF_STATUS_NOT_PD; break; }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      a.status[b] = code;
      if (code >= 0) {
        const double ld = sm.scal[0], q = engine_quad_total(sm, NT, NE);
        a.out[b] = -0.5 * ((double)a.nf * ((double)n * LOG2PI + ld) + q);
        if (a.logdet) a.logdet[b] = ld;
      } else {
        a.out[b] = nan("");
        if (a.logdet) a.logdet[b] = nan("");
      }
    }
    __syncthreads();
  }
}

// nf draws L z from N(0, K + add I) per batch item, L = jitchol(K + add I); z: Philox (chain word = batch item, sampler
// word = draw, sweep word = stream_id) or injected
__global__ void __launch_bounds__(NTHR, 1) k_gp_sample(GpArgs a) {
  const int n = a.n, NT = a.NT, npad = NT * TB, ldv = npad + 4;
  EngineSmem sm(smem_base(), NT, 0);
  double* xs = smem_base() + engine_smem_doubles(NT, 0);
  double* xsq = xs + (size_t)n * a.p;
  double* va = xsq + n;             // 8 x ldv
  double* vb = va + 8 * ldv;        // 8 x ldv
  double* W = a.ws + (size_t)blockIdx.x * a.ws_stride;
  const Stream rng{a.k0, a.k1};
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int b = blockIdx.x; b < a.batch; b += gridDim.x) {
    const double sigma = exp10(a.lsig[b]), ls = exp10(a.lls[b]);
    int tries = 0, code = 0;
    double jitter = 0.0, addv = 0.0;
    for (;;) {
      __syncthreads();
      rbf_scale_inputs(a.x, n, a.p, ls, xs, xsq);
      __syncthreads();
      const double tot = rbf_fill_tiles(W, NT, n, a.p, xs, xsq, sigma, sm.red());
      addv = a.add_abs + a.add_rel * tot / ((double)n * (double)n);
      for (int t = threadIdx.x; t < npad; t += NTHR) sm.dvec[t] = (t < n) ? addv + jitter : 0.0;
      __syncthreads();
      EngineIO io{W, nullptr, RowSrc{nullptr, nullptr, 0, 0, 0}, nullptr, 0};
      if (chol_engine<AUG_NONE, true>(SrcTiles{W}, io, NT, 0, 0, sm)) { code = tries; break; }
      const double dm = sigma + addv;
      if (tries == 0) {
        if (!(dm > 0.0)) { code = GPF_STATUS_NONPOS_DIAG; break; }
        jitter = dm * 1e-6;
      } else jitter *= 10.0;
      ++tries;
      if (tries > a.maxtries || !isfinite(jitter)) { code = GPF_STATUS_NOT_PD; break; }
    }
    __syncthreads();
    if (threadIdx.x == 0) a.status[b] = code;
    if (code >= 0)
      for (int v0 = 0; v0 < a.nf; v0 += 8) {
        const int R = (a.nf - v0 < 8) ? a.nf - v0 : 8;
        __syncthreads();
        for (int e = threadIdx.x; e < 8 * ldv; e += NTHR) va[e] = 0.0;
        __syncthreads();
        if (a.F) {
          for (int e = threadIdx.x; e < R * n; e += NTHR) va[(e / n) * ldv + e % n] = a.F[((size_t)b * a.nf + v0 + e / n) * n + e % n];
        } else {
          const int np2 = (n + 1) / 2;
          for (int q = threadIdx.x; q < R * np2; q += NTHR) {
            const int v = q / np2, pr = q % np2;
            double z0, z1;
            rng.normal_pair((unsigned)b, a.stream_id, (unsigned)(v0 + v), (unsigned)pr, z0, z1);
            va[v * ldv + 2 * pr] = z0;
            if (2 * pr + 1 < n) va[v * ldv + 2 * pr + 1] = z1;
          }
        }
        __syncthreads();
        trimv_dmma<false, NW>(W, NT, va, vb, ldv, R, warp, lane);
        __syncthreads();
        for (int e = threadIdx.x; e < R * n; e += NTHR) a.out[((size_t)b * a.nf + v0 + e / n) * n + e % n] = vb[(e / n) * ldv + e % n];
      }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------------
// FunctionDerivative._params / _sample (sample/function.py:51-66): for every batch item (one kernel, nf functions)
//   mu_f = K_ba K_a^-1 beta_f,   cov = K_b - K_ba K_a^-1 K_ba^T,   beta'_f = mu_f + L_cov z_f,
// K_a^-1 = (K + offset I)^-1 (kernel.pyx:51-64), K_b = dK(x), K_ba = dK(x, cross=True) (rbf.pyx:21-38), L_cov = jitchol(cov).
// One factorisation of K_a with the rows [beta_f ; K_ba] riding along gives  w_f = L^-1 beta_f  and  V = K_ba L^-T,
// hence mu_f = V w_f and cov = K_b - V V^T without forming the inverse.  (The reference draws through numpy's SVD map;
// equality is in distribution, as for Function._sample.)
// ---------------------------------------------------------------------------------------------------------
struct DerivArgs {
  const double* x;
  int n, p, NT, NE, batch, nf, maxtries;
  const double *lsig, *lls;
  const double* F;        // batch x nf x n
  const double* z;        // injected normals batch x nf x n or null
  double offset;
  double* dF;             // batch x nf x n
  double* mu;             // batch x nf x n or null
  double* cov;            // batch x n x n or null
  int* status;            // batch
  double* ws;             // per CTA: tiles W | Wx | cov tiles | rows (nf+n) x npad | V (nf+n) x npad
  size_t ws_stride;
  unsigned k0, k1, stream_id;
};

__host__ __device__ inline size_t deriv_ws_doubles(int NT, int NE, int nf, int n) {
  return ((size_t)2 * plt_tiles(NT) + (size_t)NE * NT) * TILE + (size_t)2 * (nf + n) * NT * TB;
}

__global__ void __launch_bounds__(NTHR, 1) k_function_derivative(DerivArgs a) {
  const int n = a.n, NT = a.NT, NE = a.NE, npad = NT * TB, nr = a.nf + n, ldv = npad + 4;
  const int nti_last = (nr - (NE - 1) * TB + 7) / 8;
  EngineSmem sm(smem_base(), NT, NE);
  double* xs = smem_base() + engine_smem_doubles(NT, NE);
  double* xsq = xs + (size_t)n * a.p;
  double* va = sm.P;                      // 8 + 8 vectors in the idle panel slots
  double* vb = sm.P + 8 * ldv;
  double* W = a.ws + (size_t)blockIdx.x * a.ws_stride;
  double* Wx = W + (size_t)plt_tiles(NT) * TILE;
  double* Wc = Wx + (size_t)NE * NT * TILE;
  double* rows = Wc + (size_t)plt_tiles(NT) * TILE;
  double* V = rows + (size_t)nr * npad;
  const Stream rng{a.k0, a.k1};
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int b = blockIdx.x; b < a.batch; b += gridDim.x) {
    const double sigma = exp10(a.lsig[b]), ls = exp10(a.lls[b]);
    __syncthreads();
    rbf_scale_inputs(a.x, n, a.p, ls, xs, xsq);
    __syncthreads();
    for (size_t e = threadIdx.x; e < (size_t)nr * npad; e += NTHR) {
      const int r = (int)(e / npad), cidx = (int)(e % npad);
      double v = 0.0;
      if (cidx < n) {
        if (r < a.nf) v = a.F[((size_t)b * a.nf + r) * n + cidx];
        else {
          const int i = r - a.nf;
          v = (-1.0 / ls) * (a.x[i * a.p] - a.x[cidx * a.p]) * rbf_elem(xs, xsq, n, a.p, sigma, i, cidx);   // rbf.pyx:34-35
        }
      }
      rows[e] = v;
    }
    // ---- K_a = L L^T with [beta ; K_ba] riding along (substitution-accurate panel solves: cond(K_a) ~ 1e11) ----
    int tries = 0, code = 0;
    double jitter = 0.0;
    for (;;) {
      __syncthreads();
      for (int t = threadIdx.x; t < npad; t += NTHR) sm.dvec[t] = (t < n) ? a.offset + jitter : 0.0;
      __syncthreads();
      EngineIO io{W, Wx, RowSrc{rows, nullptr, npad, nr, n}, V, npad};
      if (chol_engine<AUG_RHS, true>(SrcRbf{xs, xsq, n, a.p, sigma}, io, NT, NE, nti_last, sm)) { code = tries; break; }
      const double dm = sigma + a.offset;
      if (tries == 0) {
        if (!(dm > 0.0)) { code = GPF_STATUS_NONPOS_DIAG; break; }
        jitter = dm * 1e-6;
      } else jitter *= 10.0;
      ++tries;
      if (tries > a.maxtries || !isfinite(jitter)) { code = GPF_STATUS_NOT_PD; break; }
    }
    __syncthreads();
    if (code >= 0) {
      // ---- mu_f = V w_f  (kept in dF until the draw is added) ----
      for (int e = threadIdx.x; e < a.nf * n; e += NTHR) {
        const int v = e / n, i = e % n;
        const double* vi = V + (size_t)(a.nf + i) * npad;
        const double* wf = V + (size_t)v * npad;
        double m = 0.0;
        for (int k = 0; k < n; ++k) m = fma(vi[k], wf[k], m);
        const size_t o = ((size_t)b * a.nf + v) * n + i;
        if (a.mu) a.mu[o] = m;
        a.dF[o] = m;
      }
      // ---- cov = K_b - V V^T (lower tiles), jitchol(cov) with the ladder of linalg.py:41-54 ----
      int ctries = 0, ccode = 0;
      double cj = 0.0;
      for (;;) {
        __syncthreads();
        double dsum = 0.0, bad = 0.0;
        for (int bi = 0; bi < NT; ++bi)
          for (int bj = 0; bj <= bi; ++bj) {
            double* tile = Wc + (size_t)plt(bi, bj) * TILE;
            for (int e = threadIdx.x; e < TILE; e += NTHR) {
              const int i = bi * TB + (e >> 5), j = bj * TB + (e & 31);
              double v = (i == j) ? 1.0 : 0.0;
              if (i < n && j < n) {
                const double diff = a.x[i * a.p] - a.x[j * a.p];
                v = (1.0 / ls) * (1.0 - (1.0 / ls) * (diff * diff)) * rbf_elem(xs, xsq, n, a.p, sigma, i, j);   // rbf.pyx:37
                const double* vi = V + (size_t)(a.nf + i) * npad;
                const double* vj = V + (size_t)(a.nf + j) * npad;
                double acc = 0.0;
                for (int k = 0; k < n; ++k) acc = fma(vi[k], vj[k], acc);
                v -= acc;
                if (a.cov && ctries == 0) {
                  a.cov[((size_t)b * n + i) * n + j] = v;
                  a.cov[((size_t)b * n + j) * n + i] = v;
                }
                if (i == j) { dsum += v; if (!(v > 0.0)) bad += 1.0; }
              }
              tile[e] = v;
            }
          }
        dsum = block_sum(dsum, sm.red());
        bad = block_sum(bad, sm.red());
        if (bad > 0.0) { ccode = GPF_STATUS_NONPOS_DIAG; break; }
        for (int t = threadIdx.x; t < npad; t += NTHR) sm.dvec[t] = (t < n) ? cj : 0.0;
        __syncthreads();
        EngineIO io{Wc, nullptr, RowSrc{nullptr, nullptr, 0, 0, 0}, nullptr, 0};
        if (chol_engine<AUG_NONE, true>(SrcTiles{Wc}, io, NT, 0, 0, sm)) { ccode = ctries; break; }
        cj = (ctries == 0) ? dsum / n * 1e-6 : cj * 10.0;
        ++ctries;
        if (ctries > a.maxtries || !isfinite(cj)) { ccode = GPF_STATUS_NOT_PD; break; }
      }
      __syncthreads();
      if (ccode < 0) {
        // the reference's K_b / K_ba scaling (rbf.pyx:34-37) only gives a covariance for lengthscales <= 1; beyond that cov is
        // indefinite and there is no Cholesky draw: the means stay available, the draws are NaN, the status says why
        code = ccode;
        for (int e = threadIdx.x; e < a.nf * n; e += NTHR) a.dF[(size_t)b * a.nf * n + e] = nan("");
      } else {
        if (ccode > code) code = ccode;
        for (int v0 = 0; v0 < a.nf; v0 += 8) {
          const int R = (a.nf - v0 < 8) ? a.nf - v0 : 8;
          __syncthreads();
          for (int e = threadIdx.x; e < 8 * ldv; e += NTHR) va[e] = 0.0;
          __syncthreads();
          if (a.z) {
            for (int e = threadIdx.x; e < R * n; e += NTHR) va[(e / n) * ldv + e % n] = a.z[((size_t)b * a.nf + v0 + e / n) * n + e % n];
          } else {
            const int np2 = (n + 1) / 2;
            for (int q = threadIdx.x; q < R * np2; q += NTHR) {
              const int v = q / np2, pr = q % np2;
              double z0, z1;
              rng.normal_pair((unsigned)b, a.stream_id, (unsigned)(v0 + v), (unsigned)pr, z0, z1);
              va[v * ldv + 2 * pr] = z0;
              if (2 * pr + 1 < n) va[v * ldv + 2 * pr + 1] = z1;
            }
          }
          __syncthreads();
          trimv_dmma<false, NW>(Wc, NT, va, vb, ldv, R, warp, lane);
          __syncthreads();
          for (int e = threadIdx.x; e < R * n; e += NTHR) a.dF[((size_t)b * a.nf + v0 + e / n) * n + e % n] += vb[(e / n) * ldv + e % n];
        }
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) a.status[b] = code;
  }
}

// y[b][t][o] = sum_a X[o][a] F[b][a][t] + sd[b] z  (base.py:256-257), z: Philox (chain word = b, sampler word 0xffff, sweep = stream_id)
__global__ void __launch_bounds__(256) k_compose_y(const double* __restrict__ F, const double* __restrict__ X, const double* __restrict__ sd,
                                                    int batch, int n, int m, int f, double* __restrict__ y, unsigned k0, unsigned k1,
                                                    unsigned stream_id) {
  const Stream rng{k0, k1};
  const size_t total = (size_t)batch * n * m;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const int b = (int)(e / ((size_t)n * m)), t = (int)((e / m) % n), o = (int)(e % m);
    double acc = 0.0;
    for (int q = 0; q < f; ++q) acc = fma(X[(size_t)o * f + q], F[((size_t)b * f + q) * n + t], acc);
    const size_t idx = (size_t)t * m + o;
    double z0, z1;
    rng.normal_pair((unsigned)b, stream_id, 0xffffu, (unsigned)(idx >> 1), z0, z1);
    y[e] = acc + sd[b] * ((idx & 1) ? z1 : z0);
  }
}

// ---------------------------------------------------------------------------------------------------------
// FANOVA transforms (fanova.py:58-81, 106-124) of R stored states at once.  Block b maps its contrast functions
// F[in0 .. in0+nin) to nout effect rows, out[r][out0 + j][t] = sum_i coef[j][i] F[r][in0 + i][t]  (effectSample), and, if
// fpv, the finite-population variance row out0 + nout:  diag(a P a^T) / (nout - 1), P = I - 11^T/nout.
// grid = (rows, chunks of 8 output rows); HBM-bound for small blocks, FP64-FMA-bound for the Kronecker interaction blocks.
// ---------------------------------------------------------------------------------------------------------
struct TransformDesc {
  int nchunk;
  const int *chunk_blk, *chunk_j0;                       // per chunk: block and first output row within the block
  const int *in0, *nin, *out0, *nout, *coef_off, *fpv;   // per block
  const double* coef;
  int f, n, Q;                                            // functions, time points, output rows per state
};

__global__ void __launch_bounds__(128) k_transform(const double* __restrict__ F, double* __restrict__ out, TransformDesc d) {
  const int r = blockIdx.x, ck = blockIdx.y;
  const int b = d.chunk_blk[ck], j0 = d.chunk_j0[ck];
  const int nin = d.nin[b], nout = d.nout[b], in0 = d.in0[b];
  const int nj = (nout - j0 < 8) ? nout - j0 : 8;
  const double* cf = d.coef + d.coef_off[b] + (size_t)j0 * nin;
  const double* Fr = F + ((size_t)r * d.f + in0) * d.n;
  double* o = out + ((size_t)r * d.Q + d.out0[b] + j0) * d.n;
  for (int t = threadIdx.x; t < d.n; t += blockDim.x) {
    double acc[8];
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) acc[jj] = 0.0;
    for (int i = 0; i < nin; ++i) {
      const double fv = Fr[(size_t)i * d.n + t];
#pragma unroll
      for (int jj = 0; jj < 8; ++jj)
        if (jj < nj) acc[jj] = fma(cf[(size_t)jj * nin + i], fv, acc[jj]);
    }
#pragma unroll
    for (int jj = 0; jj < 8; ++jj)
      if (jj < nj) o[(size_t)jj * d.n + t] = acc[jj];
  }
}
__global__ void __launch_bounds__(128) k_transform_fpv(double* __restrict__ out, TransformDesc d, int nblk) {
  const int r = blockIdx.x;
  for (int b = 0; b < nblk; ++b) {
    if (!d.fpv[b]) continue;
    const int nout = d.nout[b];
    const double* a = out + ((size_t)r * d.Q + d.out0[b]) * d.n;
    double* o = out + ((size_t)r * d.Q + d.out0[b] + nout) * d.n;
    for (int t = threadIdx.x; t < d.n; t += blockDim.x) {
      double s = 0.0;
      for (int l = 0; l < nout; ++l) s += a[(size_t)l * d.n + t];
      const double mean = s / nout;
      double q = 0.0;
      for (int l = 0; l < nout; ++l) {
        const double v = a[(size_t)l * d.n + t];
        q = fma(v, v - mean, q);                         // a_l (P a)_l
      }
      o[t] = q / (nout - 1);                              // fanova.py:122
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Chain diagnostics on a stored history  hist[s][c][col]  (S rows, C chains): per (chain, half) mean and variance of the
// columns col0 .. col0+ncol, then split-R^ (Gelman et al. 2013), pooled mean and variance per column.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_seq_moments(const double* __restrict__ hist, int S, int C, size_t cols, int col0, int ncol,
                                                      int s0, int s1, double* __restrict__ mean, double* __restrict__ var) {
  // grid = (column chunks of 128, chains); sequence = rows s0 .. s1 of chain blockIdx.y; output index [chain][col]
  const int col = blockIdx.x * blockDim.x + threadIdx.x, c = blockIdx.y;
  if (col >= ncol) return;
  double mu = 0.0, m2 = 0.0;
  int k = 0;
  for (int s = s0; s < s1; ++s) {
    const double v = hist[((size_t)s * C + c) * cols + col0 + col];
    ++k;
    const double dlt = v - mu;
    mu += dlt / k;
    m2 = fma(dlt, v - mu, m2);
  }
  mean[(size_t)c * ncol + col] = mu;
  var[(size_t)c * ncol + col] = (k > 1) ? m2 / (k - 1) : 0.0;
}
// nseq sequences of length len each: rhat, pooled mean, pooled (posterior) variance estimate per column
__global__ void __launch_bounds__(128) k_rhat(const double* __restrict__ mean, const double* __restrict__ var, int nseq, int ncol, int len,
                                               double* __restrict__ rhat, double* __restrict__ pmean, double* __restrict__ pvar) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= ncol) return;
  double mu = 0.0, m2 = 0.0, w = 0.0;
  for (int q = 0; q < nseq; ++q) {
    const double v = mean[(size_t)q * ncol + col];
    const double dlt = v - mu;
    mu += dlt / (q + 1);
    m2 = fma(dlt, v - mu, m2);
    w += var[(size_t)q * ncol + col];
  }
  w /= nseq;
  const double bvar = (nseq > 1) ? m2 / (nseq - 1) : 0.0;       // variance of the sequence means (B / len)
  const double vhat = (double)(len - 1) / len * w + bvar;
  rhat[col] = (w > 0.0) ? sqrt(vhat / w) : nan("");
  pmean[col] = mu;
  pvar[col] = vhat;
}

// ascending sort of every row of v (rows x S, S <= 8192) in shared memory (bitonic); interval.py:32-33
__global__ void __launch_bounds__(256) k_row_sort(double* __restrict__ v, int S, int P2) {
  extern __shared__ __align__(16) double gpf_smem_sort[];
  double* s = gpf_smem_sort;
  double* row = v + (size_t)blockIdx.x * S;
  for (int i = threadIdx.x; i < P2; i += blockDim.x) s[i] = (i < S) ? row[i] : (double)INFINITY;
  __syncthreads();
  for (int k = 2; k <= P2; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < P2; i += blockDim.x) {
        const int l = i ^ j;
        if (l > i) {
          const double a = s[i], b = s[l];
          const bool up = (i & k) == 0;
          // NaNs sort last (numpy's sort order)
          const bool gt = (a > b) || (a != a && b == b);
          if (gt == up) { s[i] = b; s[l] = a; }
        }
      }
      __syncthreads();
    }
  for (int i = threadIdx.x; i < S; i += blockDim.x) row[i] = s[i];
}

// FunctionInterval (interval.py:62-131) for a batch of sample sets: samples[b][s][t] (S x n each).
// out_stats[b]: mean (n), std (n) (population, numpy's default); out_eps[b]: epsilon; out_iters[b]: evaluations of `check`.
__global__ void __launch_bounds__(256) k_band(const double* __restrict__ samples, int S, int n, double alpha_eff, double alpha, double start,
                                               double tol, int maxiter, double* __restrict__ out_mean, double* __restrict__ out_std,
                                               double* __restrict__ out_eps, int* __restrict__ out_iters) {
  extern __shared__ __align__(16) double gpf_smem_band[];
  double* lb = gpf_smem_band;          // mean - 2 std
  double* ub = lb + n;                 // mean + 2 std
  __shared__ int s_cnt;
  const double* X = samples + (size_t)blockIdx.x * S * n;
  for (int t = threadIdx.x; t < n; t += blockDim.x) {
    double s1 = 0.0;
    for (int s = 0; s < S; ++s) s1 += X[(size_t)s * n + t];
    const double mu = s1 / S;
    double s2 = 0.0;
    for (int s = 0; s < S; ++s) { const double dv = X[(size_t)s * n + t] - mu; s2 = fma(dv, dv, s2); }
    const double sd = sqrt(s2 / S);
    out_mean[(size_t)blockIdx.x * n + t] = mu;
    out_std[(size_t)blockIdx.x * n + t] = sd;
    lb[t] = mu - 2.0 * sd;
    ub[t] = mu + 2.0 * sd;
  }
  __syncthreads();
  int evals = 0;
  // fraction of samples entirely inside (lb - x, ub + x)  (interval.py:75)
  auto check = [&](double xq) -> double {
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    int mine = 0;
    for (int s = threadIdx.x; s < S; s += blockDim.x) {
      bool in = true;
      for (int t = 0; t < n && in; ++t) {
        const double v = X[(size_t)s * n + t];
        in = (v > lb[t] - xq) && (v < ub[t] + xq);
      }
      mine += in ? 1 : 0;
    }
    if (mine) atomicAdd(&s_cnt, mine);
    __syncthreads();
    const double r = (double)s_cnt / (double)S;
    __syncthreads();
    ++evals;
    return r;
  };
  double eps = start;
  while (check(eps) < alpha_eff && eps < 1.0e300) eps *= 2.0;      // interval.py:80-82
  double eLb = eps / 2.0, eUb = eps;
  int i = 0;
  for (;;) {                                                       // interval.py:87-110
    if (fabs(check(eLb) - alpha) < tol) { eps = eLb; break; }
    if (fabs(check(eUb) - alpha) < tol) { eps = eUb; break; }
    const double nb = (eLb + eUb) / 2.0;
    if (check(nb) < alpha) eLb = nb; else eUb = nb;
    ++i;
    if (i > maxiter) { eps = eLb; break; }
  }
  if (threadIdx.x == 0) {
    out_eps[blockIdx.x] = eps;
    out_iters[blockIdx.x] = evals;
  }
}

}  // namespace gpf
