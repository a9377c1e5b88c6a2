// Function._sample (sample/function.py:14-40) without ever forming K^-1.
//
// The reference builds  A = K_o^-1 + D  (K_o = K + 1e-9 I through kernel.pyx:51-64, D = diag(n_t)/(sigma_y + 1e-9),
// function.py:27-30), b = m_t/(sigma_y + 1e-9), and draws from N(A^-1 b, A^-1).  K_o has cond 1e10..1e11, so K_o^-1 -- and
// with it the reference's mu and Sigma -- carry 1e-10..1e-9 of rounding.  The same distribution in the
// "observation" form needs no inverse at all.  With  S = D^1/2  and the well-conditioned  C = I + S K_o S  (cond <= 1 +
// |D| |K_o| ~ 1e4..1e7):
//     Sigma = A^-1 = K_o - K_o S C^-1 S K_o          mu = A^-1 b = K_o S C^-1 q0,   q0 = S^-1 b  (0 where n_t = 0)
// and a draw follows by conditioning a prior draw on pseudo-observations (Matheron's rule):
//     f = L_B z1  (L_B = jitchol(K_o), the factor kernel.pyx:56 builds),    q = q0 - S f - z2,    v = C^-1 q,
//     beta = f + K_o S v = L_B (z1 + L_B^T S v)              z1, z2: 2n standard normals (Philox or injected).
// Checked against the 60-digit arbiter (tests/golden/truth.npz): mu and Sigma are 1e-14..3e-12 from the truth where the
// reference's own FP64 path is 4e-11..9e-10 away; against the reference's goldens the difference IS the reference's error.
// Costs per function: one n x n Cholesky of C (the reference: jitchol(A)), per prior group one Cholesky of K_o (the
// reference: Cholesky + triangular inverse + L^-T L^-1, per FUNCTION).
//
// On a uniform grid with an even number of time points K_o is symmetric Toeplitz, hence centrosymmetric:
//     K_o = Q^T diag(B+, B-) Q,   B+- = K11 +- K12 J,   Q = [[I, J], [I, -J]] / sqrt(2)  (orthogonal),
// so L_B is replaced by the factors of the two n/2 x n/2 blocks (z1 then lives in the split domain), and when D = c I
// (balanced design, nothing missing) C splits the same way:  C+- = I + c B+-.  Everything n^3 then runs on half-size
// matrices (a quarter of the flops, half the serial pivot chain).
#pragma once
#include "gpf_common.cuh"

namespace gpf {

constexpr double RSQRT2 = 0.70710678118654752440;

#ifdef GPF_FTRACE
// phase timers of the fused kernel (thread 0 of every 16th CTA), cycles: [0] table, [1] factor of K_o, [2] right-hand sides,
// [3] factor of C, [4] back substitution, [5] triangular products, [6] write-out, [7] items
__device__ unsigned long long gpf_ftrace[8];
#define FT_INIT long long ft_prev = clock64();
#define FT(s)                                                                 \
  do {                                                                        \
    if (threadIdx.x == 0 && (blockIdx.x & 15) == 0) {                         \
      const long long now_ = clock64();                                       \
      atomicAdd(gpf_ftrace + (s), (unsigned long long)(now_ - ft_prev));      \
      ft_prev = now_;                                                         \
    }                                                                         \
  } while (0)
#define FT_COUNT if (threadIdx.x == 0 && (blockIdx.x & 15) == 0) atomicAdd(gpf_ftrace + 7, 1ull);
#else
#define FT_INIT
#define FT(s)
#define FT_COUNT
#endif


// ---------------------------------------------------------------------------------------------------------
// jitchol(K_g + offset I) of every chain (kernel.pyx:55-56 with rbf.pyx:16-19 generated on chip), kept in c.LB for the
// function steps of a sequential sweep.  MODE 0: general inputs, 1: uniform grid (lag table), 2: uniform grid with even n:
// the two half-size blocks B+-, factored side by side by two sub-engines.
// ---------------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(NTHR, 2) k_group_factor(CtxDev c, int g_first, int g_count) {
  constexpr bool centro = (MODE == 2);
  constexpr int NSUB = centro ? 2 : 1, NWE = NW / NSUB, NTE = NWE * 32;
  const int n = c.n, hh = centro ? n / 2 : n, NTm = (hh + TB - 1) / TB;
  const size_t esz = engine_smem_doubles(NTm, 0), pltsz = (size_t)plt_tiles(NTm) * TILE;
  const int sub = threadIdx.x / NTE, tid = threadIdx.x - sub * NTE;
  const EngineSmem sm(smem_base() + (size_t)sub * esz, NTm, 0);
  const EngineSmem sm0(smem_base(), NTm, 0);
  double* xs = smem_base() + (size_t)NSUB * esz;   // MODE >= 1: lag table (n); else scaled inputs (n*p) + xsq (n)
  double* xsq = xs + (size_t)n * c.p;
  __shared__ int s_chain;
  for (int item = next_chain(c.sched, &s_chain); item < c.C * g_count; item = next_chain(c.sched, &s_chain)) {
    const int ch = item % c.C, g = g_first + item / c.C;
    const double sigma = exp10(c.hyper[(size_t)ch * c.H + 1 + 2 * g]);
    const double ls = exp10(c.hyper[(size_t)ch * c.H + 2 + 2 * g]);
    double* W = c.LB + ((size_t)g * c.C + ch) * c.lb_stride + (size_t)sub * pltsz;
    __syncthreads();
    if constexpr (MODE >= 1) toeplitz_table(c.d2, n, sigma, ls, xs, sm0.red());
    else rbf_scale_inputs(c.x, n, c.p, ls, xs, xsq);
    int tries = 0, code = 0;
    double jitter = 0.0;
    for (;;) {
      __syncthreads();
      for (int t = tid; t < NTm * TB; t += NTE) sm.dvec[t] = (t < hh) ? c.offset + jitter : 0.0;
      __syncthreads();
      EngineIO io{W, nullptr, RowSrc{nullptr, nullptr, 0, 0, 0}, nullptr, 0};
      bool okh;
      if constexpr (MODE == 2) okh = chol_engine<AUG_NONE, true, SrcCentro, NWE>(SrcCentro{xs, n, hh, sub ? -1.0 : 1.0}, io, NTm, 0, 0, sm, sub);
      else if constexpr (MODE == 1) okh = chol_engine<AUG_NONE, true>(SrcToeplitz{xs, n}, io, NTm, 0, 0, sm);
      else okh = chol_engine<AUG_NONE, true>(SrcRbf{xs, xsq, n, c.p, sigma}, io, NTm, 0, 0, sm);
      const bool ok = !__syncthreads_or(okh ? 0 : 1);
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
    if (threadIdx.x == 0) {
      c.Bjit[(size_t)g * c.C + ch] = jitter;
      status_merge(c.status + ch, code);
      atomicAdd(c.counters + CNT_CHOL, 1ull + (unsigned long long)tries);
      atomicAdd(c.counters + CNT_BFACT, 1ull + (unsigned long long)tries);
      if (tries) atomicAdd(c.counters + CNT_JITTER, (unsigned long long)tries);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Function._sample + base.functionResidual (base.py:180-190) for one function of every chain -- any design (unbalanced,
// missing observations), sequential sampler order.  Needs k_group_factor of the function's group.
// Residual projection: at time points where nothing is missing, m_t = (X^T y)[f,t] - sum_{q != f} (X^T X)[f,q] F_q(t)
// and n_t = (X^T X)[f,f] (precomputed Z, G); time points with missing observations take the per-observation loop.
// Outputs for inspection: m_t, n_t, mu, and L_C (the factor of C = I + S K_o S; Sigma = K_o - K_o S C^-1 S K_o).
// ---------------------------------------------------------------------------------------------------------
struct FnOut { double *mu, *LC, *mt, *nt; };

__host__ __device__ inline size_t function_smem_doubles(int n, int p, int NT, int mode) {
  const size_t tab = ((mode >= 1 ? (size_t)n : (size_t)n * p + n) + 3) & ~(size_t)3;
  return engine_smem_doubles(NT, 1, 8) + tab + (size_t)7 * NT * TB;
}

template <int MODE>
__global__ void __launch_bounds__(NTHR, 2) k_function(CtxDev c, int fn_single, unsigned sid_single, const int* __restrict__ fn_list,
                                                      const unsigned* __restrict__ sid_list, int nfn, unsigned sweep,
                                                      const double* __restrict__ z_inject, FnOut out) {
  constexpr bool centro = (MODE == 2);
  const int NT = c.NT, n = c.n, npad = NT * TB;
  const int hh = centro ? n / 2 : n, NTh = (hh + TB - 1) / TB, hpad = NTh * TB, ldv = (centro ? 2 * hpad : npad) + 4;
  const size_t pltsz_h = (size_t)plt_tiles(NTh) * TILE;
  const int nrows = out.mu ? 2 : 1;          // row 0: the draw; row 1: the conditional mean (inspection only)
  EngineSmem sm(smem_base(), NT, 1, 8);
  double* tab = smem_base() + engine_smem_doubles(NT, 1, 8);   // MODE >= 1: lag table (n); else scaled inputs + xsq
  double* xsq = tab + (size_t)n * c.p;
  double* sdv = tab + (((MODE >= 1 ? (size_t)n : (size_t)n * c.p + n) + 3) & ~(size_t)3);   // S = sqrt(n_t / (sigma_y + offset))
  double* z1s = sdv + npad;       // z1
  double* q = z1s + npad;         // 2 x npad right-hand sides riding along the factorisation
  double* xs = q + 2 * npad;      // 2 x npad solved rows
  double* fv = xs + 2 * npad;     // prior draw f (time domain)
  double* va = sm.P;              // 2 + 2 vectors for the triangular products (the panel slots are idle then)
  double* vb = sm.P + 2 * ldv;
  const size_t tstride = (size_t)plt_tiles(NT) * TILE;
  double* W = c.ws + (size_t)blockIdx.x * c.ws_stride;
  double* Wx = W + tstride;
  const Stream rng{c.k0, c.k1};
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __shared__ int s_chain;
  for (int item = next_chain(c.sched, &s_chain); item < c.C * nfn; item = next_chain(c.sched, &s_chain)) {
    const int ch = item % c.C;
    const int fn = fn_list ? fn_list[item / c.C] : fn_single;
    const unsigned sid = sid_list ? sid_list[item / c.C] : sid_single;
    const unsigned gch = (unsigned)(ch + c.chain_offset);
    const int g = c.group_of_fn[fn];
    const double* Grow = c.G + (size_t)fn * c.f;
    const double* Fc = c.F + (size_t)ch * c.f * n;
    const double* LBp = c.LB + ((size_t)g * c.C + ch) * c.lb_stride;
    const double jitB = c.Bjit[(size_t)g * c.C + ch];
    const double sigma = exp10(c.hyper[(size_t)ch * c.H + 1 + 2 * g]);
    const double ls = exp10(c.hyper[(size_t)ch * c.H + 2 + 2 * g]);
    const double yinv = 1.0 / (exp10(c.hyper[(size_t)ch * c.H]) + c.offset);   // white.py:7-9 through kernel.pyx:51-64
    __syncthreads();
    // ---- m_t, n_t (function.py:16-25), S, q0, the normals ----
    for (int t = threadIdx.x; t < npad; t += NTHR) {
      double mt = 0.0, nt = 0.0, z1 = 0.0, z2 = 0.0;
      if (t < n) {
        if (c.full[t]) {
          mt = c.Z[(size_t)fn * n + t];
          for (int qq = 0; qq < c.f; ++qq) {
            const double gq = Grow[qq];
            if (qq != fn && gq != 0.0) mt = fma(-gq, Fc[(size_t)qq * n + t], mt);
          }
          nt = Grow[fn];
        } else {
          for (int o = 0; o < c.m; ++o) {
            const double xf = c.X[o * c.f + fn];
            if (xf == 0.0) continue;
            const double yv = c.yT[(size_t)o * n + t];
            if (isnan(yv)) continue;
            double pred = 0.0;
            for (int qq = 0; qq < c.f; ++qq)
              if (qq != fn) pred = fma(c.X[o * c.f + qq], Fc[(size_t)qq * n + t], pred);
            mt = fma(xf, yv - pred, mt);
            nt = fma(xf, xf, nt);
          }
        }
        if (out.mt) out.mt[(size_t)ch * n + t] = mt;
        if (out.nt) out.nt[(size_t)ch * n + t] = nt;
        if (z_inject) { z1 = z_inject[(size_t)ch * 2 * n + t]; z2 = z_inject[(size_t)ch * 2 * n + n + t]; }
        else { z1 = normal_elem(rng, gch, sweep, sid, t); z2 = normal_elem(rng, gch, sweep, sid, n + t); }
      }
      const double d = nt * yinv;                       // function.py:30
      const double sd = sqrt(d);
      sdv[t] = sd;
      q[npad + t] = (d > 0.0) ? yinv * mt / sd : 0.0;   // q0 = S^-1 b, b = y_inv m  (function.py:32)
      q[t] = -z2;
      z1s[t] = z1;
      sm.dvec[t] = (t < n) ? 1.0 + d * (c.offset + jitB) : 0.0;
      xs[t] = 0.0;
      xs[npad + t] = 0.0;
    }
    for (int e = threadIdx.x; e < 4 * ldv; e += NTHR) va[e] = 0.0;   // va, vb
    __syncthreads();
    // ---- prior draw f = L_B z1 (split domain on uniform even grids) ----
    if constexpr (centro) {
      for (int e = threadIdx.x; e < 2 * hh; e += NTHR) va[(e / hh) * hpad + e % hh] = z1s[e];
      __syncthreads();
      for (int s2 = 0; s2 < 2; ++s2)
        trimv_dmma<false, NW>(LBp + s2 * pltsz_h, NTh, va + s2 * hpad, vb + s2 * hpad, ldv, 1, warp, lane);
      __syncthreads();
      for (int i = threadIdx.x; i < hh; i += NTHR) {
        fv[i] = (vb[i] + vb[hpad + i]) * RSQRT2;
        fv[n - 1 - i] = (vb[i] - vb[hpad + i]) * RSQRT2;
      }
    } else {
      for (int t = threadIdx.x; t < n; t += NTHR) va[t] = z1s[t];
      __syncthreads();
      trimv_dmma<false, NW>(LBp, NT, va, vb, ldv, 1, warp, lane);
      __syncthreads();
      for (int t = threadIdx.x; t < n; t += NTHR) fv[t] = vb[t];
    }
    __syncthreads();
    for (int t = threadIdx.x; t < n; t += NTHR) q[t] = q[npad + t] - sdv[t] * fv[t] + q[t];   // q = q0 - S f - z2
    // ---- C = I + S K_o S, factored with q (and q0) riding along: rows come out as L_C^-1 q ----
    if constexpr (MODE >= 1) toeplitz_table(c.d2, n, sigma, ls, tab, sm.red());
    else rbf_scale_inputs(c.x, n, c.p, ls, tab, xsq);
    __syncthreads();
    EngineIO io{W, Wx, RowSrc{q, nullptr, npad, nrows, n}, xs, npad};
    bool ok;
    if constexpr (MODE >= 1) ok = chol_engine<AUG_RHS, true>(SrcDiagScaled<SrcToeplitz>{SrcToeplitz{tab, n}, sdv, n}, io, NT, 1, 1, sm);
    else ok = chol_engine<AUG_RHS, true>(SrcDiagScaled<SrcRbf>{SrcRbf{tab, xsq, n, c.p, sigma}, sdv, n}, io, NT, 1, 1, sm);
    __syncthreads();
    if (ok) {
      back_solve<2>(W, NT, xs, npad, sm, nullptr);                       // v = C^-1 q  (row 1: C^-1 q0)
      // ---- beta = L_B (z1 + L_B^T S v),  mu = L_B L_B^T S C^-1 q0 ----
      for (int e = threadIdx.x; e < 4 * ldv; e += NTHR) va[e] = 0.0;
      __syncthreads();
      if constexpr (centro) {
        for (int e = threadIdx.x; e < nrows * hh; e += NTHR) {
          const int r = e / hh, i = e % hh;
          const double a = sdv[i] * xs[r * npad + i], b = sdv[n - 1 - i] * xs[r * npad + n - 1 - i];
          va[r * ldv + i] = (a + b) * RSQRT2;
          va[r * ldv + hpad + i] = (a - b) * RSQRT2;
        }
        __syncthreads();
        for (int s2 = 0; s2 < 2; ++s2)
          trimv_dmma<true, NW>(LBp + s2 * pltsz_h, NTh, va + s2 * hpad, vb + s2 * hpad, ldv, nrows, warp, lane);
        __syncthreads();
        for (int e = threadIdx.x; e < 2 * hh; e += NTHR) vb[(e / hh) * hpad + e % hh] += z1s[e];
        __syncthreads();
        for (int s2 = 0; s2 < 2; ++s2)
          trimv_dmma<false, NW>(LBp + s2 * pltsz_h, NTh, vb + s2 * hpad, va + s2 * hpad, ldv, nrows, warp, lane);
        __syncthreads();
        for (int i = threadIdx.x; i < hh; i += NTHR) {
          double* Fo = c.F + ((size_t)ch * c.f + fn) * n;
          Fo[i] = (va[i] + va[hpad + i]) * RSQRT2;
          Fo[n - 1 - i] = (va[i] - va[hpad + i]) * RSQRT2;
          if (out.mu) {
            out.mu[(size_t)ch * n + i] = (va[ldv + i] + va[ldv + hpad + i]) * RSQRT2;
            out.mu[(size_t)ch * n + n - 1 - i] = (va[ldv + i] - va[ldv + hpad + i]) * RSQRT2;
          }
        }
      } else {
        for (int e = threadIdx.x; e < nrows * n; e += NTHR) {
          const int r = e / n, t = e % n;
          va[r * ldv + t] = sdv[t] * xs[r * npad + t];
        }
        __syncthreads();
        trimv_dmma<true, NW>(LBp, NT, va, vb, ldv, nrows, warp, lane);
        __syncthreads();
        for (int t = threadIdx.x; t < n; t += NTHR) vb[t] += z1s[t];
        __syncthreads();
        trimv_dmma<false, NW>(LBp, NT, vb, va, ldv, nrows, warp, lane);
        __syncthreads();
        for (int t = threadIdx.x; t < n; t += NTHR) {
          c.F[((size_t)ch * c.f + fn) * n + t] = va[t];
          if (out.mu) out.mu[(size_t)ch * n + t] = va[ldv + t];
        }
      }
      if (out.LC) tiles_to_dense(W, n, NT, out.LC + (size_t)ch * n * n, 0);
    }
    if (threadIdx.x == 0) {
      status_merge(c.status + ch, ok ? 0 : GPF_STATUS_NOT_PD);
      atomicAdd(c.counters + CNT_CHOL, 1ull);
      atomicAdd(c.counters + CNT_DRAW, 1ull);
      atomicAdd(c.counters + CNT_FNCHOL, 1ull);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Balanced design with nothing missing (X^T X diagonal): the function draws of a sweep do not depend on each other and
// n_t = (X^T X)[f,f] =: n_f for every t, so D = c I with c = n_f / (sigma_y + offset), and functions of one prior group with
// equal n_f (a CLASS; with helmertConvert=True every function of a group) share C = I + c K_o.
// One item = (prior group, chain), everything stays on the SM or in the CTA's L2-resident workspace -- neither K^-1 nor
// any factor ever goes to HBM:
//   1. jitchol(K_o) (centro: of B+ and B-, side by side on two sub-engines)
//   2. per class, eight functions per pass: prior draws f = L_B z1 (one tensor-pipe pass over L_B), right-hand sides
//      q = sqrt(c) (ybar - f) - z2, ONE factorisation of C with the q riding along as rows, back substitution,
//      beta = f + sqrt(c) K_o v  (uniform grids: K_o v straight from the lag table; else L_B (L_B^T v)).
// ybar = (X^T y)[f] / n_f (c.bhat).
// ---------------------------------------------------------------------------------------------------------
struct FusedArgs {
  const int* gorder;      // P groups, largest work first
  const int* gcls_off;    // P+1: classes of group g are gcls_off[g] .. gcls_off[g+1]
  const int* cls_off;     // CSR offsets of the classes into cls_fn / cls_sid
  const int* cls_fn;
  const unsigned* cls_sid;
  unsigned sweep;
};

constexpr int FUSED_RB = 8;   // functions per factorisation pass (right-hand-side rows riding along)

__host__ __device__ inline size_t fused_smem_doubles(int n, int p, int mode) {
  const int nsub = mode == 2 ? 2 : 1;
  const int hh = mode == 2 ? n / 2 : n, NTm = (hh + TB - 1) / TB;
  const size_t tab = ((mode >= 1 ? (size_t)n : (size_t)n * p + n) + 3) & ~(size_t)3;
  return (size_t)nsub * engine_smem_doubles(NTm, 1, FUSED_RB) + tab;
}
// per-CTA global scratch (doubles): q rows, solved rows, prior draws -- FUSED_RB rows of nsub * hpad each
__host__ __device__ inline size_t fused_scratch_doubles(int n, int mode) {
  const int nsub = mode == 2 ? 2 : 1;
  const int hh = mode == 2 ? n / 2 : n, NTm = (hh + TB - 1) / TB;
  return (size_t)3 * FUSED_RB * nsub * NTm * TB;
}

// y[v][i] = sum_j M[i][j] x[v][j],  i < hh,  for the R vectors x + v*ld, with M = K_o (Toeplitz: tab[|i-j|]) or one of its
// centrosymmetric blocks B+- (tab[|i-j|] +- tab[n-1-i-j]) generated from the lag table, plus dg * x (offset + jitter on the
// diagonal).  Threads tid, tid + nthr, ... of the calling group each own rows; x is read as shared-memory broadcasts.
template <bool CENTRO, int RMAX>
__device__ __forceinline__ void table_matvec(const double* tab, int n, int hh, double sgn, double dg, const double* x, double* y,
                                             int ld, int R, int tid, int nthr) {
  for (int i = tid; i < hh; i += nthr) {
    double acc[RMAX];
#pragma unroll
    for (int v = 0; v < RMAX; ++v) acc[v] = 0.0;
    for (int j = 0; j < hh; ++j) {
      const int d = i - j;
      double t = tab[d < 0 ? -d : d];
      if constexpr (CENTRO) t = fma(sgn, tab[n - 1 - i - j], t);
#pragma unroll
      for (int v = 0; v < RMAX; ++v)
        if (v < R) acc[v] = fma(t, x[v * ld + j], acc[v]);
    }
#pragma unroll
    for (int v = 0; v < RMAX; ++v)
      if (v < R) y[v * ld + i] = fma(dg, x[v * ld + i], acc[v]);
  }
}

template <int MODE>
__global__ void __launch_bounds__(NTHR, 2) k_group_fused(CtxDev c, FusedArgs a) {
  constexpr bool centro = (MODE == 2);
  constexpr int NSUB = centro ? 2 : 1, NWE = NW / NSUB, NTE = NWE * 32;
  const int n = c.n, hh = centro ? n / 2 : n, NTm = (hh + TB - 1) / TB, hpad = NTm * TB, ldv = hpad + 4;
  const size_t esz = engine_smem_doubles(NTm, 1, FUSED_RB);
  const size_t pltsz = (size_t)plt_tiles(NTm) * TILE;
  const int sub = threadIdx.x / NTE, tid = threadIdx.x - sub * NTE, warp = tid >> 5, lane = tid & 31;
  double* ebase = smem_base() + (size_t)sub * esz;
  const EngineSmem sg(ebase, NTm, 1, FUSED_RB);
  double* dvec = sg.dvec;
  double* rinv = sg.rinv;
  double* scratch = sg.P;                             // panel slots, idle between factorisations: 8 + 8 vectors, then z2
  double* scratch1 = smem_base() + (size_t)(NSUB - 1) * esz + (scratch - ebase);   // the other half's (sub 1) scratch
  double* scratch0 = smem_base() + (scratch - ebase);
  double* red0 = EngineSmem(smem_base(), NTm, 1, FUSED_RB).red();
  double* tab = smem_base() + (size_t)NSUB * esz;   // MODE >= 1: lag table (n); else scaled inputs (n*p) + xsq (n)
  double* xsq = tab + (size_t)n * c.p;
  double* xa = scratch;
  double* xb = scratch + FUSED_RB * ldv;
  double* z2b = xb + FUSED_RB * ldv;                  // centro: FUSED_RB x n full-domain z2 (fits: see fused_smem check)
  double* Wb = c.ws + (size_t)blockIdx.x * c.ws_stride;
  double* LBt = Wb + (size_t)sub * pltsz;                      // factor(s) of K_o
  double* WC = Wb + (size_t)(NSUB + sub) * pltsz;              // factor(s) of C
  double* Wx = Wb + (size_t)2 * NSUB * pltsz + (size_t)sub * NTm * TILE;
  const int rowlen = NSUB * hpad;
  double* qrow = c.fws + (size_t)blockIdx.x * c.fws_stride + (size_t)sub * hpad;   // FUSED_RB rows of q, solved rows, prior draws
  double* wrow = qrow + (size_t)FUSED_RB * rowlen;
  double* frow = wrow + (size_t)FUSED_RB * rowlen;
  const double sgn = sub ? -1.0 : 1.0;
  const Stream rng{c.k0, c.k1};
  auto gsync = [&]() {
    if constexpr (NSUB == 1) __syncthreads();
    else bar_sync<NTE>(1 + 3 * sub);
  };
  // one factorisation on this half's engine; RHSROWS: q rows ride along.  All threads of the CTA get the combined verdict.
  auto factor = [&](auto src, bool with_rhs, double* Wout, int cnt) -> bool {
    bool okh;
    if (with_rhs)
      okh = chol_engine<AUG_RHS, true, decltype(src), NWE>(src, EngineIO{Wout, Wx, RowSrc{qrow, nullptr, rowlen, cnt, hh}, wrow, rowlen},
                                                         NTm, 1, (cnt + 7) / 8, sg, sub);
    else
      okh = chol_engine<AUG_NONE, true, decltype(src), NWE>(src, EngineIO{Wout, nullptr, RowSrc{nullptr, nullptr, 0, 0, 0}, nullptr, 0},
                                                          NTm, 0, 0, sg, sub);
    return !__syncthreads_or(okh ? 0 : 1);
  };
  __shared__ int s_chain;
  for (int item = next_chain(c.sched, &s_chain); item < c.C * c.P; item = next_chain(c.sched, &s_chain)) {
    const int ch = item % c.C, g = a.gorder[item / c.C];
    const unsigned gch = (unsigned)(ch + c.chain_offset);
    const double sigma = exp10(c.hyper[(size_t)ch * c.H + 1 + 2 * g]);
    const double ls = exp10(c.hyper[(size_t)ch * c.H + 2 + 2 * g]);
    const double yinv = 1.0 / (exp10(c.hyper[(size_t)ch * c.H]) + c.offset);
    __syncthreads();
    FT_INIT
    FT_COUNT
    if constexpr (MODE >= 1) toeplitz_table(c.d2, n, sigma, ls, tab, red0);
    else rbf_scale_inputs(c.x, n, c.p, ls, tab, xsq);
    FT(0);
    // ---------------- 1. jitchol(K_o) ----------------
    int tries = 0, code = 0;
    double jitter = 0.0;
    for (;;) {
      __syncthreads();
      for (int t = tid; t < hpad; t += NTE) dvec[t] = (t < hh) ? c.offset + jitter : 0.0;
      __syncthreads();
      bool ok;
      if constexpr (MODE == 2) ok = factor(SrcCentro{tab, n, hh, sgn}, false, LBt, 0);
      else if constexpr (MODE == 1) ok = factor(SrcToeplitz{tab, n}, false, LBt, 0);
      else ok = factor(SrcRbf{tab, xsq, n, c.p, sigma}, false, LBt, 0);
      if (ok) { code = tries; break; }
      const double dm = sigma + c.offset;                 // jitchol ladder (linalg.py:41-54)
      if (tries == 0) {
        if (!(dm > 0.0)) { code = GPF_STATUS_NONPOS_DIAG; break; }
        jitter = dm * 1e-6;
      } else jitter *= 10.0;
      ++tries;
      if (tries > c.maxtries || !isfinite(jitter)) { code = GPF_STATUS_NOT_PD; break; }
    }
    int nfact = 0, ndraw = 0;
    FT(1);
    // ---------------- 2. the classes of this group ----------------
    if (code >= 0)
      for (int cls = a.gcls_off[g]; cls < a.gcls_off[g + 1]; ++cls) {
        const int* fns = a.cls_fn + a.cls_off[cls];
        const unsigned* sids = a.cls_sid + a.cls_off[cls];
        const int nf = a.cls_off[cls + 1] - a.cls_off[cls];
        const double cc = c.G[(size_t)fns[0] * c.f + fns[0]] * yinv;     // function.py:30 with n_t = (X^T X)[f,f]
        const double sc = sqrt(cc);
        for (int j0 = 0; j0 < nf; j0 += FUSED_RB) {
          const int R = (nf - j0 < FUSED_RB) ? nf - j0 : FUSED_RB;
          // ---- normals: z1 in the split domain (this half's block), z2 in the time domain ----
          gsync();
          for (int e = tid; e < FUSED_RB * ldv; e += NTE) xa[e] = 0.0;
          gsync();
          {
            const int e0 = sub * hh, p0 = e0 >> 1, p1 = (e0 + hh - 1) >> 1, np1 = p1 - p0 + 1;
            for (int q = tid; q < R * np1; q += NTE) {
              const int v = q / np1, pr = p0 + q % np1;
              double za, zb;
              rng.normal_pair(gch, a.sweep, sids[j0 + v], (unsigned)pr, za, zb);
              const int ea = 2 * pr - e0, eb = ea + 1;
              if (ea >= 0 && ea < hh) xa[v * ldv + ea] = za;
              if (eb >= 0 && eb < hh) xa[v * ldv + eb] = zb;
            }
            // z2 = elements n .. 2n-1: pairs (n >> 1) ..
            const int q0 = n >> 1, q1 = (2 * n - 1) >> 1, nq = q1 - q0 + 1;
            for (int q = tid; q < R * nq; q += NTE) {
              const int v = q / nq, pr = q0 + q % nq;
              double za, zb;
              rng.normal_pair(gch, a.sweep, sids[j0 + v], (unsigned)pr, za, zb);
              const int ea = 2 * pr - n, eb = ea + 1;
              if (ea >= 0 && ea < n) z2b[v * n + ea] = za;
              if (eb >= 0 && eb < n) z2b[v * n + eb] = zb;
            }
          }
          gsync();
          // ---- prior draws f = L_B z1, right-hand sides q = sqrt(c) (ybar - f) - z2 ----
          trimv_dmma<false, NWE>(LBt, NTm, xa, xb, ldv, R, warp, lane);
          gsync();
          for (int e = tid; e < R * hh; e += NTE) {
            const int v = e / hh, i = e % hh;
            const double* yb = c.bhat + (size_t)fns[j0 + v] * n;
            double yt, z2;
            if constexpr (centro) {
              yt = (yb[i] + sgn * yb[n - 1 - i]) * RSQRT2;
              z2 = (z2b[v * n + i] + sgn * z2b[v * n + n - 1 - i]) * RSQRT2;
            } else {
              yt = yb[i];
              z2 = z2b[v * n + i];
            }
            const double fv = xb[v * ldv + i];
            frow[(size_t)v * rowlen + i] = fv;
            qrow[(size_t)v * rowlen + i] = sc * (yt - fv) - z2;
          }
          // ---- C = I + c K_o with the q riding along ----
          gsync();
          FT(2);
          for (int t = tid; t < hpad; t += NTE) dvec[t] = (t < hh) ? 1.0 + cc * (c.offset + jitter) : 0.0;
          __syncthreads();
          bool okc;
          if constexpr (MODE == 2) okc = factor(SrcScaled<SrcCentro>{SrcCentro{tab, n, hh, sgn}, cc, hh}, true, WC, R);
          else if constexpr (MODE == 1) okc = factor(SrcScaled<SrcToeplitz>{SrcToeplitz{tab, n}, cc, n}, true, WC, R);
          else okc = factor(SrcScaled<SrcRbf>{SrcRbf{tab, xsq, n, c.p, sigma}, cc, n}, true, WC, R);
          FT(3);
          ++nfact;
          if (!okc) { code = GPF_STATUS_NOT_PD; break; }
          // ---- v = L_C^-T w ----
          for (int e = tid; e < FUSED_RB * hpad; e += NTE) {
            const int v = e / hpad, i = e % hpad;
            xa[v * ldv + i] = (v < R && i < hh) ? wrow[(size_t)v * rowlen + i] : 0.0;
          }
          if (R <= 2) back_solve_g<2, NWE>(WC, NTm, xa, ldv, rinv, sub);
          else back_solve_g<8, NWE>(WC, NTm, xa, ldv, rinv, sub);
          FT(4);
          // ---- beta = f + sqrt(c) K_o v ----
          if constexpr (MODE >= 1) {
            if (R <= 2) table_matvec<centro, 2>(tab, n, hh, sgn, c.offset + jitter, xa, xb, ldv, R, tid, NTE);
            else table_matvec<centro, FUSED_RB>(tab, n, hh, sgn, c.offset + jitter, xa, xb, ldv, R, tid, NTE);
            gsync();
            for (int e = tid; e < R * hh; e += NTE) {
              const int v = e / hh, i = e % hh;
              xa[v * ldv + i] = fma(sc, xb[v * ldv + i], frow[(size_t)v * rowlen + i]);
            }
          } else {
            trimv_dmma<true, NWE>(LBt, NTm, xa, xb, ldv, R, warp, lane);       // K_o v = L_B (L_B^T v)
            gsync();
            trimv_dmma<false, NWE>(LBt, NTm, xb, xa, ldv, R, warp, lane);
            gsync();
            for (int e = tid; e < R * hh; e += NTE) {
              const int v = e / hh, i = e % hh;
              xa[v * ldv + i] = fma(sc, xa[v * ldv + i], frow[(size_t)v * rowlen + i]);
            }
          }
          __syncthreads();
          FT(5);
          for (int e = threadIdx.x; e < R * hh; e += NTHR) {
            const int v = e / hh, i = e % hh;
            double* Fo = c.F + ((size_t)ch * c.f + fns[j0 + v]) * n;
            if constexpr (centro) {
              const double bp = scratch0[v * ldv + i], bm = scratch1[v * ldv + i];
              Fo[i] = (bp + bm) * RSQRT2;
              Fo[n - 1 - i] = (bp - bm) * RSQRT2;
            } else Fo[i] = scratch0[v * ldv + i];
          }
          __syncthreads();
          FT(6);
          ndraw += R;
        }
        if (code < 0) break;
      }
    if (threadIdx.x == 0) {
      status_merge(c.status + ch, code);
      atomicAdd(c.counters + CNT_CHOL, 1ull + (unsigned long long)tries + (unsigned long long)nfact);
      atomicAdd(c.counters + CNT_BFACT, 1ull + (unsigned long long)tries);
      atomicAdd(c.counters + CNT_FNCHOL, (unsigned long long)nfact);
      atomicAdd(c.counters + CNT_DRAW, (unsigned long long)ndraw);
      if (tries) atomicAdd(c.counters + CNT_JITTER, (unsigned long long)tries);
    }
  }
}

}  // namespace gpf
