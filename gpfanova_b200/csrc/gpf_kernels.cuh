// Kernels of the Gibbs hot path.  One CTA (NTHR threads) works on one chain at a time (grid-stride over chains);
// every n x n matrix is factorised by the tile engine (gpf_tile.cuh) in a per-CTA workspace that stays L2-resident.
#pragma once
#include "gpf_common.cuh"

namespace gpf {

// shared-memory carve-up common to all kernels: [engine | extra]
__device__ __forceinline__ double* smem_base() {
  extern __shared__ __align__(16) double gpf_smem[];
  return gpf_smem;
}

// ---------------------------------------------------------------------------------------------------------
// Stateless operators (dense row-major in/out)
// ---------------------------------------------------------------------------------------------------------

// RBF.dist + RBF._K (kernel/rbf.pyx:8-19) with 10** of kernel.pyx:35-37.  grid.x = batch, dense output.
__global__ void __launch_bounds__(NTHR) k_rbf_dense(const double* __restrict__ x, int n, int p,
                                                     const double* __restrict__ lsig, const double* __restrict__ lls,
                                                     double* __restrict__ K) {
  double* xs = smem_base();
  double* xsq = xs + (size_t)n * p;
  const int b = blockIdx.x;
  const double sigma = exp10(lsig[b]), ls = exp10(lls[b]);
  rbf_scale_inputs(x, n, p, ls, xs, xsq);
  __syncthreads();
  double* out = K + (size_t)b * n * n;
  for (size_t e = threadIdx.x; e < (size_t)n * n; e += NTHR) {
    const int R = (int)(e / n), Cc = (int)(e % n);
    out[e] = rbf_elem(xs, xsq, n, p, sigma, R, Cc);
  }
}

// mean of the diagonal of a dense matrix + any non-positive flag (linalg.py:41-44); all threads get the results
__device__ inline double dense_diag_mean(const double* A, int n, double* red, bool& nonpos) {
  double s = 0.0, bad = 0.0;
  for (int t = threadIdx.x; t < n; t += NTHR) {
    const double d = A[(size_t)t * n + t];
    s += d;
    if (d <= 0.0) bad += 1.0;
  }
  s = block_sum(s, red);
  bad = block_sum(bad, red);
  nonpos = bad > 0.0;
  return s / n;
}

// linalg.jitchol (linalg.py:35-54) per matrix; ws: grid.x workspaces of plt_tiles(NT) tiles.
__global__ void __launch_bounds__(NTHR, 2) k_jitchol_dense(const double* __restrict__ A, int n, int NT, int batch,
                                                           int maxtries, double add_diag, double* __restrict__ L,
                                                           double* __restrict__ logdet, int* __restrict__ status,
                                                           double* ws, size_t ws_stride) {
  EngineSmem sm(smem_base(), NT, 0);
  double* W = ws + (size_t)blockIdx.x * ws_stride;
  for (int b = blockIdx.x; b < batch; b += gridDim.x) {
    const double* Ab = A + (size_t)b * n * n;
    int tries = 0, code = 0;
    double jitter = 0.0;
    for (;;) {
      __syncthreads();
      dense_to_tiles(Ab, n, NT, W);
      for (int t = threadIdx.x; t < NT * TB; t += NTHR) sm.dvec[t] = (t < n) ? add_diag + jitter : 0.0;
      __syncthreads();
      EngineIO io{W, nullptr, RowSrc{nullptr, nullptr, 0, 0, 0}, nullptr, 0};
      if (chol_engine<AUG_NONE, true>(SrcTiles{W}, io, NT, 0, 0, sm)) { code = tries; break; }
      if (tries == 0) {
        bool nonpos;
        const double dm = dense_diag_mean(Ab, n, sm.red(), nonpos) + add_diag;
        if (nonpos) { code = GPF_STATUS_NONPOS_DIAG; break; }
        jitter = dm * 1e-6;
      } else jitter *= 10.0;
      ++tries;
      if (tries > maxtries || !isfinite(jitter)) { code = GPF_STATUS_NOT_PD; break; }
    }
    __syncthreads();
    if (code >= 0) {
      tiles_to_dense(W, n, NT, L + (size_t)b * n * n, 0);
      if (threadIdx.x == 0 && logdet) logdet[b] = sm.scal[0];
    } else if (threadIdx.x == 0 && logdet) logdet[b] = nan("");
    if (threadIdx.x == 0) status[b] = code;
  }
}

// Kernel.K_inv (kernel/kernel.pyx:51-64) on dense input: (A + offset I)^-1 through jitchol + in-place inversion
__global__ void __launch_bounds__(NTHR, 2) k_kinv_dense(const double* __restrict__ A, int n, int NT, int batch,
                                                        int maxtries, double offset, double* __restrict__ Ainv,
                                                        int* __restrict__ status, double* ws, size_t ws_stride) {
  EngineSmem sm(smem_base(), NT, 0);
  double* W = ws + (size_t)blockIdx.x * ws_stride;
  for (int b = blockIdx.x; b < batch; b += gridDim.x) {
    const double* Ab = A + (size_t)b * n * n;
    int tries = 0, code = 0;
    double jitter = 0.0;
    for (;;) {
      __syncthreads();
      dense_to_tiles(Ab, n, NT, W);
      for (int t = threadIdx.x; t < NT * TB; t += NTHR) sm.dvec[t] = (t < n) ? offset + jitter : 0.0;
      __syncthreads();
      EngineIO io{W, nullptr, RowSrc{nullptr, nullptr, 0, 0, 0}, nullptr, 0};
      if (chol_engine<AUG_IDENT, false>(SrcTiles{W}, io, NT, 0, 0, sm)) { code = tries; break; }
      if (tries == 0) {
        bool nonpos;
        const double dm = dense_diag_mean(Ab, n, sm.red(), nonpos) + offset;
        if (nonpos) { code = GPF_STATUS_NONPOS_DIAG; break; }
        jitter = dm * 1e-6;
      } else jitter *= 10.0;
      ++tries;
      if (tries > maxtries || !isfinite(jitter)) { code = GPF_STATUS_NOT_PD; break; }
    }
    __syncthreads();
    if (code >= 0) tiles_to_dense(W, n, NT, Ainv + (size_t)b * n * n, 1);
    if (threadIdx.x == 0) status[b] = code;
  }
}

// packed tiles of one chain -> dense symmetric (inspection output of gpf_group_kinv / L_A of gpf_function_step)
__global__ void __launch_bounds__(NTHR) k_tiles_to_dense(const double* tiles, size_t tile_stride, int n, int NT, int batch,
                                                         double* __restrict__ out, int mode) {
  for (int b = blockIdx.x; b < batch; b += gridDim.x)
    tiles_to_dense(tiles + (size_t)b * tile_stride, n, NT, out + (size_t)b * n * n, mode);
}

// ---------------------------------------------------------------------------------------------------------
// Kernel.K_inv of prior group g for every chain at the current hyper-parameters
//   kernel/kernel.pyx:51-64 (K + 1e-9 I, jitchol, L^-T L^-1) with rbf.pyx:16-19 generated on chip at first touch
//   (MODE 1: uniform grid, one exp per lag; MODE 0: one exp per element; MODE 2: uniform grid with even n -- the
//   centrosymmetric split, two n/2 inverses in the CTA's workspace assembled into c.Kinv).
// The inverse is built in place in its destination (c.Kinv), never as a separate L / L^-1.
// ---------------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(NTHR, 2) k_group_kinv(CtxDev c, int g_first, int g_count) {
  EngineSmem sm(smem_base(), (MODE == 2) ? (c.n / 2 + TB - 1) / TB : c.NT, 0);
  // MODE >= 1: lag table (n); else scaled inputs (n*p) + xsq (n).  MODE 2 keeps two half-size engines in shared memory
  double* xs = smem_base() + ((MODE == 2) ? 2 * engine_smem_doubles((c.n / 2 + TB - 1) / TB, 0) : engine_smem_doubles(c.NT, 0));
  double* xsq = xs + (size_t)c.n * c.p;
  const size_t tstride = (size_t)plt_tiles(c.NT) * TILE;
  __shared__ int s_chain;
  // items = (group, chain) pairs: g_count groups x C chains, handed out dynamically
  for (int item = next_chain(c.sched, &s_chain); item < c.C * g_count; item = next_chain(c.sched, &s_chain)) {
    const int ch = item % c.C, g = g_first + item / c.C;
    const double sigma = exp10(c.hyper[(size_t)ch * c.H + 1 + 2 * g]);
    const double ls = exp10(c.hyper[(size_t)ch * c.H + 2 + 2 * g]);
    double* W = c.Kinv + ((size_t)g * c.C + ch) * tstride;
    __syncthreads();
    if constexpr (MODE >= 1) toeplitz_table(c.d2, c.n, sigma, ls, xs, sm.red());
    else rbf_scale_inputs(c.x, c.n, c.p, ls, xs, xsq);
    int tries = 0, code = 0;
    double jitter = 0.0;
    constexpr bool centro = (MODE == 2);
    const int h = c.n / 2, NTh = (h + TB - 1) / TB;
    double* Wh = c.ws + (size_t)blockIdx.x * c.ws_stride;          // centro: the two half-size inverses
    const size_t hstride = (size_t)plt_tiles(NTh) * TILE;
    for (;;) {
      __syncthreads();
      bool ok = true;
      if constexpr (centro) {
        // the two halves side by side: warps [0, NW/2) invert B+, warps [NW/2, NW) invert B-, each half a sub-engine
        // with its own shared memory and barriers (twice the matrices in flight per SM for these latency-bound sizes)
        constexpr int NWH = NW / 2;
        const int sub = threadIdx.x / (NWH * 32);
        const EngineSmem smh(smem_base() + (size_t)sub * engine_smem_doubles(NTh, 0), NTh, 0);
        for (int t = threadIdx.x - sub * NWH * 32; t < NTh * TB; t += NWH * 32) smh.dvec[t] = (t < h) ? c.offset + jitter : 0.0;
        __syncthreads();
        EngineIO io{Wh + sub * hstride, nullptr, RowSrc{nullptr, nullptr, 0, 0, 0}, nullptr, 0};
        const bool okh = chol_engine<AUG_IDENT, false, SrcCentro, NWH>(SrcCentro{xs, c.n, h, sub ? -1.0 : 1.0}, io, NTh, 0, 0, smh, sub);
        ok = !__syncthreads_or(okh ? 0 : 1);
      }
      if constexpr (!centro) {
        for (int t = threadIdx.x; t < c.NT * TB; t += NTHR) sm.dvec[t] = (t < c.n) ? c.offset + jitter : 0.0;
        __syncthreads();
        EngineIO io{W, nullptr, RowSrc{nullptr, nullptr, 0, 0, 0}, nullptr, 0};
        if constexpr (MODE == 1) ok = chol_engine<AUG_IDENT, false>(SrcToeplitz{xs, c.n}, io, c.NT, 0, 0, sm);
        else ok = chol_engine<AUG_IDENT, false>(SrcRbf{xs, xsq, c.n, c.p, sigma}, io, c.NT, 0, 0, sm);
      }
      if (ok) { code = tries; break; }
      // jitchol ladder (linalg.py:41-54): diag(K + offset I) = sigma + offset everywhere
      const double dm = sigma + c.offset;
      if (tries == 0) {
        if (!(dm > 0.0)) { code = GPF_STATUS_NONPOS_DIAG; break; }
        jitter = dm * 1e-6;
      } else jitter *= 10.0;
      ++tries;
      if (tries > c.maxtries || !isfinite(jitter)) { code = GPF_STATUS_NOT_PD; break; }
    }
    if constexpr (centro) if (code >= 0) {
      // K^-1 = 1/2 [[S, D J], [J D, J S J]],  S = M+ + M-,  D = M+ - M-  (M+- = B+-^-1, symmetric, lower tiles valid)
      __syncthreads();
      const double* M1 = Wh;
      const double* M2 = Wh + hstride;
      const int n = c.n;
      auto msym = [&](const double* M, int a, int b) {
        if (a < b) { const int q = a; a = b; b = q; }
        return M[(size_t)plt(a >> 5, b >> 5) * TILE + (a & 31) * TB + (b & 31)];
      };
      if ((n & 63) == 0) {
        // tile-aligned halves: every output tile is one tile of M+ (+/-) M-, possibly row-flipped (coalesced 16-byte
        // copies) or flipped and transposed (through the warp's idle panel slot)
        constexpr int NWH = NW / 2;
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, T = c.NT, Th = T / 2;
        const int per = (Th < NWH) ? Th : NWH, nwk = 2 * per;           // warps with a panel slot of their own
        const int wq = (warp / NWH) * per + (warp % NWH);
        if ((warp % NWH) < per) {
          const EngineSmem smw(smem_base() + (size_t)(warp / NWH) * engine_smem_doubles(NTh, 0), NTh, 0);
          double* slot = smw.P + (size_t)(warp % NWH) * TB * PS;
          for (int t = wq; t < plt_tiles(T); t += nwk) {
            int bi = 0, bj = t;
            while (bj > bi) { bj -= bi + 1; ++bi; }
            double* tile = W + (size_t)plt(bi, bj) * TILE;
            int ta, tb;
            double sg = 1.0;
            bool flip = false, transp = false;
            if (bi < Th) { ta = bi; tb = bj; }
            else if (bj < Th) {
              sg = -1.0; ta = T - 1 - bi; tb = bj; flip = true;
              if (ta < tb) { const int q = ta; ta = tb; tb = q; transp = true; }   // out[r][c] = D[tb*32 + c][ta*32 + 31 - r]
            } else { ta = T - 1 - bj; tb = T - 1 - bi; flip = true; transp = true; }  // out[r][c] = S[ta*32+31-c][tb*32+31-r]
            const double* s1 = M1 + (size_t)plt(ta, tb) * TILE;
            const double* s2 = M2 + (size_t)plt(ta, tb) * TILE;
            // loads in batches of 8 + 8 before the first store: the source (workspace) and the destination may alias as far
            // as the compiler knows, and a load -> store -> load chain would pay one L2 latency per 16 bytes
            if (!transp) {
#pragma unroll
              for (int q0 = 0; q0 < 16; q0 += 8) {
                double2 a[8], b[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                  const int idx = (q0 + q) * 32 + lane, row = idx >> 4, ch2 = 2 * (idx & 15), srow = flip ? 31 - row : row;
                  a[q] = *reinterpret_cast<const double2*>(s1 + srow * TB + ch2);
                  b[q] = *reinterpret_cast<const double2*>(s2 + srow * TB + ch2);
                }
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                  const int idx = (q0 + q) * 32 + lane, row = idx >> 4, ch2 = 2 * (idx & 15);
                  __stcs(reinterpret_cast<double2*>(tile + row * TB + ch2),
                         make_double2(0.5 * fma(sg, b[q].x, a[q].x), 0.5 * fma(sg, b[q].y, a[q].y)));
                }
              }
            } else {
              __syncwarp();
#pragma unroll
              for (int q0 = 0; q0 < 16; q0 += 8) {
                double2 a[8], b[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                  const int idx = (q0 + q) * 32 + lane, row = idx >> 4, ch2 = 2 * (idx & 15);
                  a[q] = *reinterpret_cast<const double2*>(s1 + row * TB + ch2);
                  b[q] = *reinterpret_cast<const double2*>(s2 + row * TB + ch2);
                }
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                  const int idx = (q0 + q) * 32 + lane, row = idx >> 4, ch2 = 2 * (idx & 15);
                  *reinterpret_cast<double2*>(slot + row * PS + ch2) =
                      make_double2(0.5 * fma(sg, b[q].x, a[q].x), 0.5 * fma(sg, b[q].y, a[q].y));
                }
              }
              __syncwarp();
              // lower-left block: source element (c, 31 - r); lower-right block: source element (31 - c, 31 - r)
              const int srow = (bj < Th) ? lane : 31 - lane;
#pragma unroll 8
              for (int r = 0; r < 32; ++r) __stcs(tile + r * TB + lane, slot[srow * PS + 31 - r]);
            }
          }
        }
      } else {
      // 8 independent element pairs in flight per thread (the loads come from L2: latency, not bandwidth)
      constexpr int U = 8;
      for (int bi = 0; bi < c.NT; ++bi)
        for (int bj = 0; bj <= bi; ++bj) {
          double* tile = W + (size_t)plt(bi, bj) * TILE;
          for (int e0 = threadIdx.x; e0 < TILE; e0 += NTHR * U) {
            double m1[U], m2[U], sg[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
              const int e = e0 + u * NTHR;
              int R = bi * TB + (e >> 5), Cc = bj * TB + (e & 31);
              m1[u] = 0.0; m2[u] = 0.0; sg[u] = 0.0;
              if (e < TILE) {
                if (R >= n || Cc >= n) { m1[u] = (R == Cc) ? 2.0 : 0.0; }
                else {
                  if (R < Cc) { const int q = R; R = Cc; Cc = q; }
                  int a, b;
                  if (R < h) { a = R; b = Cc; sg[u] = 1.0; }
                  else if (Cc < h) { a = n - 1 - R; b = Cc; sg[u] = -1.0; }
                  else { a = n - 1 - R; b = n - 1 - Cc; sg[u] = 1.0; }
                  m1[u] = msym(M1, a, b);
                  m2[u] = msym(M2, a, b);
                }
              }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
              const int e = e0 + u * NTHR;
              if (e < TILE) tile[e] = 0.5 * fma(sg[u], m2[u], m1[u]);
            }
          }
        }
      }
    }
    if (threadIdx.x == 0) {
      status_merge(c.status + ch, code);
      atomicAdd(c.counters + CNT_CHOL, 1ull + (unsigned long long)tries);
      atomicAdd(c.counters + CNT_INV, 1ull);
      if (tries) atomicAdd(c.counters + CNT_JITTER, (unsigned long long)tries);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Base.observationLikelihood (base.py:196-214): the sum of squared residuals is computed once per chain, after
// which every candidate is O(1):  -N (log sd + log(2 pi)/2) - SSE / (2 sd^2) + log U(lb,ub)(sd),  sd = sqrt(10^x).
// mode 0: evaluate at cand[ch] -> out[ch].   mode 1: slice step on y_sigma (sample/slice.pyx:24-62).
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NTHR) k_obs(CtxDev c, int mode, const double* __restrict__ cand, double* __restrict__ outv,
                                              double w, int mstep, unsigned sampler_id, unsigned sweep,
                                              const double* __restrict__ u_inject, int n_inject, int* __restrict__ nevals) {
  __shared__ double red[NW];
  const int n = c.n;
  const Stream rng{c.k0, c.k1};
  for (int ch = blockIdx.x; ch < c.C; ch += gridDim.x) {
    const double* Fc = c.F + (size_t)ch * c.f * n;
    double sse = 0.0, cnt = 0.0;
    if (c.bhat) {
      // balanced design, nothing missing: the residual sum of squares splits orthogonally about the least-squares fit,
      //   ||y - X F||^2 = ||y - X bhat||^2 + sum_f (X^T X)[f,f] ||F_f - bhat_f||^2
      // (all terms non-negative: no cancellation), O(f n) per chain instead of O(n m f)
      for (int e = threadIdx.x; e < c.f * n; e += NTHR) {
        const int q = e / n;
        const double dlt = Fc[e] - c.bhat[e];
        sse = fma(c.G[(size_t)q * c.f + q] * dlt, dlt, sse);
      }
      if (threadIdx.x == 0) { sse += c.rss0; cnt = (double)n * (double)c.m; }
    } else
    for (int t = threadIdx.x; t < n; t += NTHR) {
      for (int o = 0; o < c.m; ++o) {
        const double yv = c.yT[(size_t)o * n + t];
        if (isnan(yv)) continue;
        double pred = 0.0;
        for (int q = 0; q < c.f; ++q) pred = fma(c.X[o * c.f + q], Fc[(size_t)q * n + t], pred);
        const double r = yv - pred;
        sse = fma(r, r, sse);
        cnt += 1.0;
      }
    }
    sse = block_sum(sse, red);
    cnt = block_sum(cnt, red);
    if (mode == 0) {
      if (threadIdx.x == 0) outv[ch] = obs_loglik_eval(sse, cnt, cand[ch], c.prior_lb, c.prior_ub);
    } else {
      const unsigned gch = (unsigned)(ch + c.chain_offset);
      int nev;
      const double x0 = c.hyper[(size_t)ch * c.H];
      const double x1 = slice_step(
          x0, w, mstep, [&](double v) { return obs_loglik_eval(sse, cnt, v, c.prior_lb, c.prior_ub); },
          [&](int k) { return (u_inject && k < n_inject) ? u_inject[(size_t)ch * n_inject + k] : rng.uniform(gch, sweep, sampler_id, (unsigned)k); },
          nev);
      __syncthreads();
      if (threadIdx.x == 0) {
        c.hyper[(size_t)ch * c.H] = x1;
        if (nevals) nevals[ch] = nev;
        atomicAdd(c.counters + CNT_LOGP, (unsigned long long)nev);
        atomicAdd(c.counters + CNT_SLICE, 1ull);
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------------
// Base.prior_likelihood (base.py:216-243) and the slice steps on prior{g}_sigma / prior{g}_lengthscale.
//   cov = K(sigma, ls) + mean(K) 1e-6 I;  ll = 1 + sum_{f in g} log N(beta_f; 0, cov) + log U(lb,ub)(cand)
// Every evaluation that changes the lengthscale is one factorisation with the group's functions riding along as
// right-hand sides (quadratic forms) and the log-determinant from the pivots.  Candidates that only change sigma
// rescale cov exactly, cov(s') = (s'/s) cov(s), so they cost O(1) after one factorisation -- and in the reference's
// sampler order (sigma slice, then lengthscale slice of the same group) the lengthscale slice's first density value
// is such a rescaling too, so the fused mode needs no factorisation for it.
// The slice loops are state machines (SliceSM) around ONE factorisation call site.
//   mode 0: evaluate sigma candidate cand[ch] -> out[ch]      mode 1: evaluate lengthscale candidate
//   mode 2: slice step on prior{g}_sigma                      mode 3: slice step on prior{g}_lengthscale
//   mode 4: sigma slice followed by lengthscale slice (sampler ids sampler_id, sampler_id + 1)
// ---------------------------------------------------------------------------------------------------------

// slice-step state of one chain, kept in shared memory and advanced by thread 0 only (keeps the registers of the
// other 255 threads for the factorisation)
struct PriorState {
  SliceSM sl, ss;
  double s_fact, l_eval, result;
  int base, running, inb, nfact, failed, nev_s;
};

template <bool TOEP>
__global__ void __launch_bounds__(NTHR, 2) k_prior(CtxDev c, PriorArgs a) {
  const int NT = c.NT, n = c.n, g = a.g, mode = a.mode;
  const int ng = c.group_off[g + 1] - c.group_off[g];
  const int NE = (ng + TB - 1) / TB;
  const int nti_last = (ng - (NE - 1) * TB + 7) / 8;
  EngineSmem sm(smem_base(), NT, NE);
  double* xs = smem_base() + engine_smem_doubles(NT, NE);   // TOEP: lag table (n); else scaled inputs (n*p) + xsq (n)
  double* xsq = xs + (size_t)n * c.p;
  double* W = c.ws + (size_t)blockIdx.x * c.ws_stride;
  double* Wx = W + (size_t)plt_tiles(NT) * TILE;
  const int* fns = c.group_fns + c.group_off[g];
  __shared__ PriorState st;
  __shared__ int s_chain;
  for (int ch = next_chain(c.sched, &s_chain); ch < c.C; ch = next_chain(c.sched, &s_chain)) {
    const double cur_s = c.hyper[(size_t)ch * c.H + 1 + 2 * g];
    const double cur_l = c.hyper[(size_t)ch * c.H + 2 + 2 * g];
    if (threadIdx.x == 0) {
      st.base = (mode == 0 || mode == 2 || mode == 4);   // next factorisation is the one at the current (sigma, ls)
      st.s_fact = cur_s;
      st.l_eval = (mode == 1) ? a.cand[ch] : cur_l;
      st.result = 0.0;
      st.sl.init(cur_l, a.w_l, a.m_l);
      st.nfact = st.failed = st.nev_s = 0;
      st.running = 1;
      st.inb = st.base || (st.l_eval >= c.prior_lb && st.l_eval <= c.prior_ub);   // base.py:224-231: outside -> -inf
    }
    for (;;) {
      __syncthreads();
      if (!st.running) break;
      const bool inb = st.inb != 0;
      bool ok = true;
      if (inb) {
        // ---- the factorisation: cov = K(10^s_fact, 10^l_eval) + mean(K) 1e-6 I, group functions as right-hand sides ----
        const double sig = exp10(st.s_fact), ls = exp10(st.l_eval);
        double tot;
        if constexpr (TOEP) tot = toeplitz_table(c.d2, n, sig, ls, xs, sm.red());
        else {
          rbf_scale_inputs(c.x, n, c.p, ls, xs, xsq);
          __syncthreads();
          tot = rbf_fill_tiles(W, NT, n, c.p, xs, xsq, sig, sm.red());
        }
        const double jit = tot / ((double)n * (double)n) * c.prior_jitter;   // base.py:222
        for (int t = threadIdx.x; t < NT * TB; t += NTHR) sm.dvec[t] = (t < n) ? jit : 0.0;
        __syncthreads();
        EngineIO io{W, Wx, RowSrc{c.F + (size_t)ch * c.f * n, fns, n, ng, n}, nullptr, 0};
        if constexpr (TOEP) ok = chol_engine<AUG_RHS, false>(SrcToeplitz{xs, n}, io, NT, NE, nti_last, sm);
        else ok = chol_engine<AUG_RHS, false>(SrcTiles{W}, io, NT, NE, nti_last, sm);
      }
      if (threadIdx.x == 0) {
        // ---- thread 0 advances the slice state machine(s) with this evaluation ----
        const Stream rng{c.k0, c.k1};
        const unsigned gch = (unsigned)(ch + c.chain_offset);
        const double lb = c.prior_lb, ub = c.prior_ub, lprior = -log(ub - lb), nlog2pi = (double)n * LOG2PI;
        double ld = 0.0, q = 0.0;
        if (inb) {
          ++st.nfact;
          if (ok) { ld = sm.scal[0]; q = engine_quad_total(sm, NT, NE); }
          else st.failed = 1;
        }
        const unsigned sid_l = a.sampler_id + (mode == 4 ? 1u : 0u);
        auto unif_s = [&](int k) {
          return (a.u_inject && mode == 2 && k < a.n_inject) ? a.u_inject[(size_t)ch * a.n_inject + k]
                                                              : rng.uniform(gch, a.sweep, a.sampler_id, (unsigned)k);
        };
        auto unif_l = [&](int k) {
          return (a.u_inject && mode == 3 && k < a.n_inject) ? a.u_inject[(size_t)ch * a.n_inject + k]
                                                              : rng.uniform(gch, a.sweep, sid_l, (unsigned)k);
        };
        if (st.base) {
          st.base = 0;
          // density of a sigma candidate v by exact rescaling of the factorisation at cur_s (base.py:235-243)
          auto lp_s = [&](double v, double pr) -> double {
            if (!ok) return -(double)INFINITY;
            const double dl = v - cur_s;
            return 1.0 - 0.5 * ((double)ng * (nlog2pi + ld + (double)n * LN10 * dl) + q * exp10(-dl)) + pr;
          };
          auto prior_of = [&](double v) -> double { return (v < lb || v > ub) ? -(double)INFINITY : lprior; };
          if (mode == 0) {
            st.result = lp_s(a.cand[ch], prior_of(a.cand[ch]));
            st.running = 0;
          } else {
            st.ss.init(cur_s, a.w_s, a.m_s);
            while (!st.ss.done) st.ss.feed(lp_s(st.ss.cand, prior_of(st.ss.cand)), unif_s);
            st.nev_s = st.ss.nev;
            const double new_s = st.ss.cand;
            c.hyper[(size_t)ch * c.H + 1 + 2 * g] = new_s;
            if (mode == 2) st.running = 0;
            else {
              st.s_fact = new_s;
              st.sl.feed(lp_s(new_s, prior_of(cur_l)), unif_l);   // density of the current lengthscale under the new sigma
              st.l_eval = st.sl.cand;
            }
          }
        } else {
          const double lp = (inb && ok) ? 1.0 - 0.5 * ((double)ng * (nlog2pi + ld) + q) + lprior : -(double)INFINITY;
          if (mode == 1) { st.result = lp; st.running = 0; }
          else {
            st.sl.feed(lp, unif_l);
            if (st.sl.done) {
              st.running = 0;
              c.hyper[(size_t)ch * c.H + 2 + 2 * g] = st.sl.cand;
            } else st.l_eval = st.sl.cand;
          }
        }
        st.inb = (st.l_eval >= lb && st.l_eval <= ub);
        if (!st.running) {
          if (mode <= 1) a.outv[ch] = st.result;
          if (st.failed) status_merge(c.status + ch, GPF_STATUS_NOT_PD);
          atomicAdd(c.counters + CNT_CHOL, (unsigned long long)st.nfact);
          if (mode >= 2) {
            const int nev_total = (mode == 2 ? st.nev_s : (mode == 3 ? st.sl.nev : st.nev_s + st.sl.nev));
            if (a.nevals) {
              a.nevals[ch] = (mode == 3) ? st.sl.nev : st.nev_s;
              if (mode == 4) a.nevals[c.C + ch] = st.sl.nev;
            }
            atomicAdd(c.counters + CNT_LOGP, (unsigned long long)nev_total);
            atomicAdd(c.counters + CNT_SLICE, mode == 4 ? 2ull : 1ull);
          }
        }
      }
    }
  }
}

}  // namespace gpf
