// Slice-sampler GP log-likelihood on a uniform time grid through the Schur algorithm.
//
// On a uniform grid the slice covariance  cov = K(sigma, ls) + mean(K) 1e-6 I  (base.py:221-222) is a symmetric
// positive definite TOEPLITZ matrix  T[i][j] = t[|i-j|].  Its displacement  T - Z T Z^T = u^T u - v^T v  has rank 2
// (u = t / sqrt(t0), v = u with v0 = 0), and the Schur algorithm turns the generator (u, v) into the columns of the
// lower Cholesky factor, one column per step, with one hyperbolic rotation and O(n) work per step:
//     column k of L  = u^(k)                      (stored from the diagonal down: us[j] = L[k+j][k])
//     rho = v[k+1] / us[0],  c = sqrt(1 - rho^2)  (|rho| < 1  <=>  T positive definite)
//     us'[j] = (us[j] - rho v[k+1+j]) / c,   v'[k+1+j] = c v[k+1+j] - rho us'[j]        (mixed, stable form)
// so a whole evaluation is ~n^2 multiply-adds instead of n^3/3, needs no n x n storage, and L[k+1][k+1] = L[k][k] c.
// The log-determinant comes from the rotation cosines, and the quadratic forms beta^T cov^-1 beta = |L^-1 beta|^2 of
// the group's functions from a column-oriented forward substitution that consumes each column as it appears.
// One WARP works on one chain (vectors in shared memory, lanes over the vector index); the Neal slice state machine
// (SliceSM) runs redundantly in the warp's registers.  General (non-uniform) inputs use the dense engine (k_prior).
// Stability: Bojanczyk, Brent, de Hoog & Sweet, "On the stability of the Bareiss and related Toeplitz factorization
// algorithms" (1995) -- for positive definite T the computed factor satisfies a Cholesky-like backward error bound.
#pragma once
#include "gpf_common.cuh"
#include "gpf_schur_args.cuh"

namespace gpf {


struct SchurOut { double logdet, quad; bool ok; };

// all lanes of the warp call this with identical arguments; smem regions are the warp's own
__device__ __forceinline__ SchurOut schur_eval(const CtxDev& c, int n, int ng, const int* __restrict__ fns,
                                               const double* __restrict__ Fc, double sigma, double ls, double* us, double* v,
                                               double* ev, double* r, int lane) {
  // lag table and its full-matrix sum (base.py:222's cov.mean())
  const double s = -0.5 / (ls * ls);
  double part = 0.0;
  for (int k = lane; k < n; k += 32) {
    const double t = sigma * exp(c.d2[k] * s);
    us[k] = t;
    part += (k == 0 ? (double)n : 2.0 * (double)(n - k)) * t;
  }
  part = warp_sum(part);
  const double jit = part / ((double)n * (double)n) * c.prior_jitter;
  __syncwarp();
  const double t0 = us[0] + jit;
  SchurOut out{0.0, 0.0, true};
  if (!(t0 > 0.0) || !(t0 < 1.0e300)) { out.ok = false; return out; }
  double uinv = rsqrt_nr(t0);                 // 1 / L[0][0]
  __syncwarp();
  for (int k = lane; k < n; k += 32) {
    const double u = (k == 0 ? t0 : us[k]) * uinv;
    us[k] = u;
    v[k] = (k == 0) ? 0.0 : u;
  }
  for (int q = 0; q < ng; ++q)
    for (int k = lane; k < n; k += 32) r[q * n + k] = Fc[(size_t)fns[q] * n + k];
  __syncwarp();
  double quad = 0.0;
  bool ok = true;
  for (int k = 0; k < n; ++k) {
    const int m = n - 1 - k;                  // entries below the diagonal in column k
    // ---- forward substitution with column k:  w = r[k] / L[k][k];  r[i] -= L[i][k] w  (i > k) ----
    for (int q = 0; q < ng; ++q) {
      double* rq = r + q * n + k;
      const double w = rq[0] * uinv;
      quad = fma(w, w, quad);
      for (int j = 1 + lane; j <= m; j += 32) rq[j] = fma(-us[j], w, rq[j]);
    }
    if (m == 0) break;
    // ---- hyperbolic rotation that produces column k+1 ----
    const double rho = v[k + 1] * uinv;
    const double e = fma(-rho, rho, 1.0);
    if (!(e > 0.0)) { ok = false; break; }
    const double cinv = rsqrt_nr(e);
    const double cc = e * cinv;
    if (lane == 0) ev[k] = e;
    uinv *= cinv;                             // L[k+1][k+1] = L[k][k] c
    __syncwarp();
    for (int j = lane; j < m; j += 32) {
      const double un = (us[j] - rho * v[k + 1 + j]) * cinv;
      const double vn = fma(cc, v[k + 1 + j], -rho * un);
      us[j] = un;
      v[k + 1 + j] = vn;
    }
    __syncwarp();
  }
  if (!ok) { out.ok = false; return out; }
  // logdet = 2 sum_k log L[k][k],  L[k][k] = sqrt(t0) prod_{j<k} sqrt(e_j)
  __syncwarp();
  double ld = 0.0;
  for (int j = lane; j < n - 1; j += 32) ld += (double)(n - 1 - j) * log(ev[j]);
  ld = warp_sum(ld);
  out.logdet = (double)n * log(t0) + ld;
  out.quad = quad;
  return out;
}

// Register-resident version for n <= 256 and at most NG functions per group: lane l owns positions 8l .. 8l+7 of every
// vector (column of L, v, right-hand sides).  v and the right-hand sides never move; the column shifts down by one
// position per step, which with the step loop unrolled by 8 is a compile-time renaming of the lane's registers plus ONE
// shuffle (the value crossing the lane boundary).  The shared-memory version above moves 32 bytes per multiply-add
// through the LSU and is bound by it (measured 53k cycles per evaluation per SM); this one is bound by the FP64 pipe.
// Positions above the diagonal of the current column hold dead values that are never read by live ones.
// log det = -2 sum_k log(1/L[k][k]) is accumulated as a product with the exponent split off every 8 steps.
template <int NG>
__device__ __forceinline__ SchurOut schur_eval_reg(const CtxDev& c, int n, int ng, const int* __restrict__ fns,
                                                   const double* __restrict__ Fc, double sigma, double ls, int lane) {
  double col[8], v[8], r[NG][8];
  const double sc = -0.5 / (ls * ls);
  double part = 0.0;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int i = lane * 8 + q;
    const double t = (i < n) ? sigma * exp(c.d2[i] * sc) : 0.0;
    col[q] = t;
    if (i < n) part += (i == 0 ? (double)n : 2.0 * (double)(n - i)) * t;
  }
  part = warp_sum(part);
  const double jit = part / ((double)n * (double)n) * c.prior_jitter;   // base.py:222
  const double t0 = __shfl_sync(FULLMASK, col[0], 0) + jit;
  SchurOut out{0.0, 0.0, true};
  if (!(t0 > 0.0) || !(t0 < 1.0e300)) { out.ok = false; return out; }
  double uinv = rsqrt_nr(t0);                 // 1 / L[0][0]
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int i = lane * 8 + q;
    const double u = ((i == 0) ? t0 : col[q]) * uinv;
    col[q] = u;
    v[q] = (i == 0) ? 0.0 : u;
  }
#pragma unroll
  for (int qq = 0; qq < NG; ++qq)
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int i = lane * 8 + q;
      r[qq][q] = (qq < ng && i < n) ? Fc[(size_t)fns[qq] * n + i] : 0.0;
    }
  double prod = 1.0, quad = 0.0;
  int e2 = 0;
  bool ok = true;
  const int nblk = (n + 7) >> 3;
  for (int K = 0; K < nblk; ++K) {
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      const int k = 8 * K + s;
      // ---- forward substitution with column k (logical position q of this lane is col[(q - s) & 7]) ----
#pragma unroll
      for (int qq = 0; qq < NG; ++qq) {
        // (positions >= n pick up shifted-out column entries: they are dead, and steps k >= n contribute nothing)
        const double w = (k < n) ? __shfl_sync(FULLMASK, r[qq][s], K) * uinv : 0.0;
        quad = fma(w, w, quad);
#pragma unroll
        for (int q = 0; q < 8; ++q) r[qq][q] = fma(-col[(q - s) & 7], w, r[qq][q]);
      }
      if (k < n) prod *= uinv;
      // ---- hyperbolic rotation that produces column k+1 ----
      double vk1 = (s < 7) ? __shfl_sync(FULLMASK, v[(s + 1) & 7], K) : __shfl_sync(FULLMASK, v[0], (K + 1) & 31);
      const double rho = (k + 1 < n) ? vk1 * uinv : 0.0;
      const double e = fma(-rho, rho, 1.0);
      if (!(e > 0.0)) ok = false;
      const double cinv = rsqrt_nr(e);
      const double cc = e * cinv;
      uinv *= cinv;                             // L[k+1][k+1] = L[k][k] c
      // shift the column down by one position: only the register holding logical position 7 crosses lanes
      col[(7 - s) & 7] = __shfl_up_sync(FULLMASK, col[(7 - s) & 7], 1);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int pc = (q - s - 1) & 7;
        const double un = (col[pc] - rho * v[q]) * cinv;
        v[q] = fma(cc, v[q], -rho * un);
        col[pc] = un;
      }
    }
    int ex;
    prod = frexp(prod, &ex);
    e2 += ex;
    if (!ok) break;
  }
  if (!ok) { out.ok = false; return out; }
  out.logdet = -2.0 * (log(prod) + (double)e2 * 0.6931471805599453);
  out.quad = quad;
  return out;
}

// ---- staged, square-root-free version (the one the kernels call) ---------------------------------------------------------------------
// Two changes against schur_eval_reg above, same recursion:
//  (1) SCALED GENERATOR.  With (u, v) multiplied by the running product of the rotation cosines the rotation needs neither the
//      square root nor the division by c:   u' = u - rho v,   v' = e v - rho u'   (e = 1 - rho^2; three multiply-adds per position
//      instead of four).  The pivot of the scaled column is the Schur complement  g_k = L[k][k]^2 = g_{k-1} e_{k-1},  so
//          rho_k = v[k+1] / g_k,      log det = sum_k log g_k,      L[p][k] w_k = u[p] (r[k] / g_k),     w_k^2 = r[k]^2 / g_k
//      and the only non-FMA operation left on the loop-carried chain is one reciprocal (1 / e).  Same mixed (stable) downdating
//      form, up to the diagonal scaling -- checked against 60-digit factorisations next to the normalised recursion
//      (tools/schur_scaled_probe.py: errors equal within noise).
//  (2) STAGES.  Column k only has n - k live positions.  A lane owns W consecutive slots; after 16 W steps the lower sixteen
//      lanes hold nothing live any more, so the upper sixteen lanes' slots are spread over all 32 lanes at W/2 slots per lane
//      (W shuffles per vector) and the recursion goes on with half the instructions per step: 8 -> 4 -> 2 -> 1 slots per lane,
//      1376 instead of 2048 vector instructions per vector for n = 256.  n <= 128 / 64 / 32 enter at W = 4 / 2 / 1.
struct SchurScal {
  double ginv, prod, quad;   // 1 / g_k, running product of the ginv (exponent split off into e2), sum of w_k^2
  int e2;
  bool ok;
};

// 1/x from the MUFU seed (rel. error r ~ 2^-23) and one third-order step y (1 + r + r^2); x > 0 finite (callers check)
__device__ __forceinline__ double rcp_nr(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double r = fma(-x, y, 1.0);
  return fma(y, fma(r, r, r), y);
}

// steps  base .. base + 16 W - 1  (W == 1: 32 steps) of the recursion; slot p = base + W lane + q is held in register q
template <int W, int NG>
__device__ __forceinline__ void schur_stage(int n, int base, int lane, double (&col)[W], double (&v)[W], double (&r)[NG][W],
                                            SchurScal& sc) {
  constexpr int NK = (W == 1) ? 32 : 16;
  constexpr int FR = (W >= 8) ? 1 : 8 / W;      // blocks between two exponent splits of prod (every 8 steps)
  for (int K = 0; K < NK; ++K) {
    if (base + W * K >= n || !sc.ok) break;
#pragma unroll
    for (int s = 0; s < W; ++s) {
      const int k = base + W * K + s;
      // ---- forward substitution with column k (logical slot q of this lane is col[(q - s) mod W]) ----
#pragma unroll
      for (int qq = 0; qq < NG; ++qq) {
        const double rraw = __shfl_sync(FULLMASK, r[qq][s], K);
        const double rk = (k < n) ? rraw : 0.0;
        const double wt = rk * sc.ginv;
        sc.quad = fma(rk, wt, sc.quad);
#pragma unroll
        for (int q = 0; q < W; ++q) r[qq][q] = fma(-col[(q - s + W) % W], wt, r[qq][q]);
      }
      if (k < n) sc.prod *= sc.ginv;
      // ---- hyperbolic rotation that produces column k+1 ----
      const double vk1 = (s < W - 1) ? __shfl_sync(FULLMASK, v[(s + 1) % W], K) : __shfl_sync(FULLMASK, v[0], (K + 1) & 31);
      const double rho = (k + 1 < n) ? vk1 * sc.ginv : 0.0;
      const double e = fma(-rho, rho, 1.0);
      if (!(e > 0.0)) sc.ok = false;
      sc.ginv *= rcp_nr(sc.ok ? e : 1.0);       // g_{k+1} = g_k e
      // shift the column down by one slot: only the register holding the lane's last logical slot crosses lanes
      col[(W - 1 - s + W) % W] = __shfl_up_sync(FULLMASK, col[(W - 1 - s + W) % W], 1);
#pragma unroll
      for (int q = 0; q < W; ++q) {
        const int pc = (q - s - 1 + 2 * W) % W;
        const double un = fma(-rho, v[q], col[pc]);
        v[q] = fma(-rho, un, e * v[q]);
        col[pc] = un;
      }
    }
    if ((K + 1) % FR == 0 || K == NK - 1) {
      int ex;
      sc.prod = frexp(sc.prod, &ex);
      sc.e2 += ex;
    }
  }
}

// the upper sixteen lanes' 16 W slots, re-dealt to all 32 lanes at W/2 slots each
template <int W>
__device__ __forceinline__ void schur_repack(const double (&in)[W], double (&out)[W / 2], int lane) {
  const int src = 16 + (lane >> 1);
#pragma unroll
  for (int q = 0; q < W / 2; ++q) {
    const double lo = __shfl_sync(FULLMASK, in[q], src);
    const double hi = __shfl_sync(FULLMASK, in[q + W / 2], src);
    out[q] = (lane & 1) ? hi : lo;
  }
}

template <int W, int NG>
__device__ __forceinline__ void schur_chain(int n, int base, int lane, double (&col)[W], double (&v)[W], double (&r)[NG][W],
                                            SchurScal& sc) {
  schur_stage<W, NG>(n, base, lane, col, v, r, sc);
  if constexpr (W > 1) {
    if (base + 16 * W < n && sc.ok) {
      double col2[W / 2], v2[W / 2], r2[NG][W / 2];
      schur_repack<W>(col, col2, lane);
      schur_repack<W>(v, v2, lane);
#pragma unroll
      for (int qq = 0; qq < NG; ++qq) schur_repack<W>(r[qq], r2[qq], lane);
      schur_chain<W / 2, NG>(n, base + 16 * W, lane, col2, v2, r2, sc);
    }
  }
}

// lag table, jitter, generator and right-hand sides at W0 slots per lane, then the staged recursion
template <int W0, int NG>
__device__ __forceinline__ SchurOut schur_eval_entry(const CtxDev& c, int n, int ng, const int* __restrict__ fns,
                                                     const double* __restrict__ Fc, double sigma, double ls, int lane) {
  double col[W0], v[W0], r[NG][W0];
  const double scl = -0.5 / (ls * ls);
  double part = 0.0;
#pragma unroll
  for (int q = 0; q < W0; ++q) {
    const int i = lane * W0 + q;
    const double t = (i < n) ? sigma * exp(c.d2[i] * scl) : 0.0;
    col[q] = t;
    if (i < n) part += (i == 0 ? (double)n : 2.0 * (double)(n - i)) * t;
  }
  part = warp_sum(part);
  const double jit = part / ((double)n * (double)n) * c.prior_jitter;   // base.py:222
  const double t0 = __shfl_sync(FULLMASK, col[0], 0) + jit;
  SchurOut out{0.0, 0.0, true};
  if (!(t0 > 0.0) || !(t0 < 1.0e300)) { out.ok = false; return out; }
#pragma unroll
  for (int q = 0; q < W0; ++q) {
    const int i = lane * W0 + q;
    if (i == 0) col[q] = t0;
    v[q] = (i == 0) ? 0.0 : col[q];
  }
#pragma unroll
  for (int qq = 0; qq < NG; ++qq)
#pragma unroll
    for (int q = 0; q < W0; ++q) {
      const int i = lane * W0 + q;
      r[qq][q] = (qq < ng && i < n) ? __ldcg(Fc + (size_t)fns[qq] * n + i) : 0.0;
    }
  SchurScal sc{rcp_nr(t0), 1.0, 0.0, 0, true};
  schur_chain<W0, NG>(n, 0, lane, col, v, r, sc);
  if (!sc.ok) { out.ok = false; return out; }
  out.logdet = -(log(sc.prod) + (double)sc.e2 * 0.6931471805599453);
  out.quad = sc.quad;
  return out;
}

template <int NG>
__device__ __forceinline__ SchurOut schur_eval_staged(const CtxDev& c, int n, int ng, const int* __restrict__ fns,
                                                      const double* __restrict__ Fc, double sigma, double ls, int lane) {
  if (n > 128) return schur_eval_entry<8, NG>(c, n, ng, fns, Fc, sigma, ls, lane);
  if (n > 64) return schur_eval_entry<4, NG>(c, n, ng, fns, Fc, sigma, ls, lane);
  if (n > 32) return schur_eval_entry<2, NG>(c, n, ng, fns, Fc, sigma, ls, lane);
  return schur_eval_entry<1, NG>(c, n, ng, fns, Fc, sigma, ls, lane);
}

// warp-level chain hand-out
__device__ __forceinline__ int next_chain_warp(int* sched, int lane) {
  unsigned t = 0u;
  if (lane == 0) t = (unsigned)atomicAdd(sched, 1);
  return (int)__reduce_max_sync(FULLMASK, t);      // (REDUX writes a uniform register: the chain loop is provably convergent)
}


// one chain's slice step(s) on the hyper-parameters of prior group a.g (all lanes of the warp; modes as k_prior)
template <int NG>   // NG > 0: register-resident evaluation (n <= 256, group size <= NG); 0: shared-memory evaluation
__device__ __forceinline__ void schur_slice_chain(const CtxDev& c, const PriorArgs& a, int ch, int lane, double* us, double* v, double* ev,
                                                  double* r) {
  const int n = c.n, g = a.g, mode = a.mode;
  const int ng = c.group_off[g + 1] - c.group_off[g];
  const int* fns = c.group_fns + c.group_off[g];
  const Stream rng{c.k0, c.k1};
  const double lb = c.prior_lb, ub = c.prior_ub, lprior = -log(ub - lb), nlog2pi = (double)n * LOG2PI;
  {
    const unsigned gch = (unsigned)(ch + c.chain_offset);
    const double* Fc = c.F + (size_t)ch * c.f * n;
    const double cur_s = __ldcg(c.hyper + (size_t)ch * c.H + 1 + 2 * g);   // (L2: another SM may have written it)
    const double cur_l = __ldcg(c.hyper + (size_t)ch * c.H + 2 + 2 * g);
    const unsigned sid_l = a.sampler_id + (mode == 4 ? 1u : 0u);
    auto unif_s = [&](int k) {
      return (a.u_inject && mode == 2 && k < a.n_inject) ? a.u_inject[(size_t)ch * a.n_inject + k]
                                                          : rng.uniform(gch, a.sweep, a.sampler_id, (unsigned)k);
    };
    auto unif_l = [&](int k) {
      return (a.u_inject && mode == 3 && k < a.n_inject) ? a.u_inject[(size_t)ch * a.n_inject + k]
                                                          : rng.uniform(gch, a.sweep, sid_l, (unsigned)k);
    };
    bool basef = (mode == 0 || mode == 2 || mode == 4);
    double s_fact = cur_s, l_eval = (mode == 1) ? a.cand[ch] : cur_l, result = 0.0;
    SliceSM sl;
    sl.init(cur_l, a.w_l, a.m_l);
    int nfact = 0, failed = 0, nev_s = 0;
    bool running = true;
    while (running) {
      const bool inb = basef || (l_eval >= lb && l_eval <= ub);   // base.py:224-231: outside the prior -> -inf
      SchurOut o{0.0, 0.0, true};
      if (inb) {
        if constexpr (NG > 0) {
#ifdef GPF_SCHUR_LEGACY
          o = schur_eval_reg<NG>(c, n, ng, fns, Fc, exp10(s_fact), exp10(l_eval), lane);
#else
          o = schur_eval_staged<NG>(c, n, ng, fns, Fc, exp10(s_fact), exp10(l_eval), lane);
#endif
        }
        else o = schur_eval(c, n, ng, fns, Fc, exp10(s_fact), exp10(l_eval), us, v, ev, r, lane);
        ++nfact;
        if (!o.ok) failed = 1;
      }
      if (basef) {
        basef = false;
        auto lp_s = [&](double x, double pr) -> double {   // exact rescaling cov(s') = (s'/s) cov(s), base.py:235-243
          if (!o.ok) return -(double)INFINITY;
          const double dl = x - cur_s;
          return 1.0 - 0.5 * ((double)ng * (nlog2pi + o.logdet + (double)n * LN10 * dl) + o.quad * exp10(-dl)) + pr;
        };
        auto prior_of = [&](double x) -> double { return (x < lb || x > ub) ? -(double)INFINITY : lprior; };
        if (mode == 0) {
          result = lp_s(a.cand[ch], prior_of(a.cand[ch]));
          running = false;
        } else {
          SliceSM ss;
          ss.init(cur_s, a.w_s, a.m_s);
          while (!ss.done) ss.feed(lp_s(ss.cand, prior_of(ss.cand)), unif_s);
          nev_s = ss.nev;
          const double new_s = ss.cand;
          if (lane == 0) c.hyper[(size_t)ch * c.H + 1 + 2 * g] = new_s;
          if (mode == 2) running = false;
          else {
            s_fact = new_s;
            sl.feed(lp_s(new_s, prior_of(cur_l)), unif_l);
            l_eval = sl.cand;
          }
        }
      } else {
        const double lp = (inb && o.ok) ? 1.0 - 0.5 * ((double)ng * (nlog2pi + o.logdet) + o.quad) + lprior : -(double)INFINITY;
        if (mode == 1) { result = lp; running = false; }
        else {
          sl.feed(lp, unif_l);
          if (sl.done) {
            running = false;
            if (lane == 0) c.hyper[(size_t)ch * c.H + 2 + 2 * g] = sl.cand;
          } else l_eval = sl.cand;
        }
      }
    }
    if (lane == 0) {
      if (mode <= 1) a.outv[ch] = result;
      if (failed) status_merge(c.status + ch, GPF_STATUS_NOT_PD);
      atomicAdd(c.counters + CNT_CHOL, (unsigned long long)nfact);
      if (mode >= 2) {
        const int nev_total = (mode == 2 ? nev_s : (mode == 3 ? sl.nev : nev_s + sl.nev));
        if (a.nevals) {
          a.nevals[ch] = (mode == 3) ? sl.nev : nev_s;
          if (mode == 4) a.nevals[c.C + ch] = sl.nev;
        }
        atomicAdd(c.counters + CNT_LOGP, (unsigned long long)nev_total);
        atomicAdd(c.counters + CNT_SLICE, mode == 4 ? 2ull : 1ull);
      }
    }
    __syncwarp();
  }
}

// the y_sigma slice step of one chain by one warp; the residual sum of squares in k_obs's summation order
__device__ __forceinline__ void schur_ysigma_warp(const CtxDev& c, const Stream& rng, int ch, unsigned sweep, double w_y, int m_y,
                                                  unsigned sid_y, int lane) {
  const int n = c.n, fn_ = c.f * n;
  const double* Fc = c.F + (size_t)ch * c.f * n;
  double sse = 0.0;
#pragma unroll
  for (int vw = 0; vw < NW; ++vw) {              // virtual warp vw of k_obs's CTA: threads 32 vw + lane
    double part = 0.0;
    for (int e = 32 * vw + lane; e < fn_; e += NTHR) {
      const int q = e / n;
      const double dlt = __ldcg(Fc + e) - c.bhat[e];
      part = fma(c.G[(size_t)q * c.f + q] * dlt, dlt, part);
    }
    if (vw == 0 && lane == 0) part += c.rss0;
    sse += warp_sum(part);
  }
  const double cnt = (double)n * (double)c.m;
  const unsigned gch = (unsigned)(ch + c.chain_offset);
  int nev;
  const double x0 = __ldcg(c.hyper + (size_t)ch * c.H);
  const double x1 = slice_step(
      x0, w_y, m_y, [&](double v) { return obs_loglik_eval(sse, cnt, v, c.prior_lb, c.prior_ub); },
      [&](int k) { return rng.uniform(gch, sweep, sid_y, (unsigned)k); }, nev);
  if (lane == 0) {
    c.hyper[(size_t)ch * c.H] = x1;
    atomicAdd(c.counters + CNT_LOGP, (unsigned long long)nev);
    atomicAdd(c.counters + CNT_SLICE, 1ull);
  }
}


// end of a chain's group unit in a pipelined sweep: count it; the last group of the chain's sweep stores the row and samples
// y_sigma for the next sweep (the kernels of the next sweep start after all slice kernels of this one have finished)
__device__ __forceinline__ void schur_fold_tail(const CtxDev& c, const PriorArgs& a, int ch, int lane) {
  __threadfence();
  __syncwarp();
  int old = 0;
  if (lane == 0) old = atomicAdd(a.fold_done + ch, 1);
  old = __shfl_sync(FULLMASK, old, 0);
  if (old + 1 != a.fold_target) return;
  __threadfence();                                // the other groups' units counted before this one: their data is in L2
  const int n = c.n;
  if (a.fold_row != nullptr) {
    double* dst = a.fold_row + (size_t)ch * a.fold_cols;
    const double* hy = c.hyper + (size_t)ch * c.H;
    const double* Fc = c.F + (size_t)ch * c.f * n;
    for (int e = lane; e < c.H; e += 32) dst[e] = __ldcg(hy + e);
    for (int e = lane; e < c.f * n; e += 32) __stcs(dst + c.H + e, __ldcg(Fc + e));
  }
  if (a.fold_next_y) schur_ysigma_warp(c, Stream{c.k0, c.k1}, ch, a.sweep + 1u, a.w_y, a.m_y, 0u, lane);
}

template <int NG>
__global__ void __launch_bounds__(NG > 0 ? 128 : NTHR, schur_minb(NG)) k_prior_schur(CtxDev c, PriorArgs a) {
  extern __shared__ __align__(16) double gpf_smem_schur[];
  const int n = c.n, g = a.g;
  const int ng = c.group_off[g + 1] - c.group_off[g];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* base = gpf_smem_schur + (size_t)warp * schur_warp_doubles(n, ng);
  double *us = base, *v = base + n, *ev = base + 2 * n, *r = base + 4 * n;   // base + 3n: spare (alignment of r)
  for (int ch = next_chain_warp(c.sched, lane); ch < c.C; ch = next_chain_warp(c.sched, lane)) {
    schur_slice_chain<NG>(c, a, ch, lane, us, v, ev, r);
    if (NG > 0 && a.fold_done != nullptr) schur_fold_tail(c, a, ch, lane);
  }
}

}  // namespace gpf
