// Function._sample (sample/function.py:14-40) on a uniform time grid in O(n^2) per function: the Schur algorithm instead
// of dense factorisations.
//
// Balanced design with nothing missing on a uniform grid: both matrices of the observation form (gpf_cond.cuh)
//     K_o = K + (offset + jitter) I          and          C = I + c K_o,   c = n_f / (sigma_y + offset),
// are symmetric positive definite TOEPLITZ matrices, so the Schur algorithm (gpf_schur.cuh) yields the columns of their
// Cholesky factors one per step with one hyperbolic rotation and O(n) work -- no n x n storage, no n^3.  One WARP does one
// (prior group, chain) item, every vector register-resident (n <= 256, at most four functions per group; the scaled,
// square-root-free form of the recursion and the staged register layout are those of gpf_schur.cuh):
//   1. columns of L_B = chol(K_o):  prior draws  f_j = L_B z1_j  accumulate as the columns appear (never stored);
//   2. right-hand sides  q_j = sqrt(c) (ybar_j - f_j) - z2_j  for every function of the group;
//   3. per class (functions sharing c): the generalized Schur recursion on the embedding [[C, I], [I, 0]] -- whose
//      displacement generator is the Toeplitz one extended by the rows [e0 e0]/sqrt(t0) -- yields column k of L_C AND column
//      k of L_C^-T from the same rotation: forward substitution  y_k = (L_C^-1 q)_k  and accumulation  v += L_C^-T[:, k] y_k
//      as the columns appear.  No column is ever stored and there is no back substitution (measured against 50-digit
//      solves at cond(C) = 5e2 .. 2e8: 3 to 80 times closer to the truth than LAPACK's Cholesky solve);
//   4. beta_j = f_j + sqrt(c) K_o v_j = ybar_j - (z2_j + v_j) / sqrt(c)      (c K_o = C - I, so K_o v never has to be formed).
// The draw contract is function_draw_obs of the oracle with the FULL Cholesky factor of K_o as its square root.
// C has cond <= 1 + c |K_o| (1e4 .. 1e7), far better than the slice covariance the same recursion already factors; for
// K_o (cond ~ 1e11) a breakdown (a reflection coefficient reaching 1) takes the jitter ladder of linalg.py:41-54, exactly
// like a failed dpotrf.
#pragma once
#include "gpf_schur.cuh"

namespace gpf {


// Register layouts.  The recursion works on two sets of vectors that only meet in the scalars rho, e, 1/g of a step:
//   * the COLUMN set (column k of L, v, right-hand sides / prior-draw accumulators) is live on positions >= k: it starts at
//     W0 slots per lane (W0 = 8, 4, 2, 1 for n <= 256, 128, 64, 32) and is re-dealt at half the width whenever its lower
//     sixteen lanes have died (schur_repack) -- slot p = 32 (W0 - WC) + WC lane + q at width WC;
//   * the INVERSE set (column k of L^-T, its partner b, the solution accumulators) is live on positions <= k: it starts at
//     one slot per lane and is spread out to twice the width eight steps before position 32 WA would be touched
//     (sfn_expand) -- slot p = WA lane + q.
// Every switch falls on a multiple of 8 steps, where the register renaming of both sets is back at the identity.

// first column of  scale * K + diag_add I  as the (unnormalised) generator; returns 1 / g_0
template <int W>
__device__ __forceinline__ double sfn_init(const CtxDev& c, int n, double sigma, double ls, double scale, double diag_add,
                                           double (&col)[W], double (&v)[W], int lane, bool& ok) {
  const double sc = -0.5 / (ls * ls);
#pragma unroll
  for (int q = 0; q < W; ++q) {
    const int i = lane * W + q;
    col[q] = (i < n) ? scale * sigma * exp(c.d2[i] * sc) : 0.0;
  }
  const double t0 = __shfl_sync(FULLMASK, col[0], 0) + diag_add;
  ok = (t0 > 0.0) && (t0 < 1.0e300);
#pragma unroll
  for (int q = 0; q < W; ++q) {
    const int i = lane * W + q;
    if (i == 0) col[q] = t0;
    v[q] = (i == 0) ? 0.0 : col[q];
  }
  return ok ? rcp_nr(t0) : 1.0;
}

// steps k0 .. k1-1 (multiples of max(WC, WA)); op.column(s, k, Kc, col, a, ginv) sees every column before it is rotated away.
// INV: the same shift + rotation on the inverse set (a, b): a becomes column k+1 of L^-T (times the common scale).
template <int WC, int WA, bool INV, class Op>
__device__ __forceinline__ void sfn_segment(int n, int k0, int k1, int baseC, int lane, double (&col)[WC], double (&v)[WC],
                                            double (&a)[WA], double (&b)[WA], double& ginv, bool& ok, Op& op) {
  constexpr int U = (WC > WA) ? WC : WA;
  for (int kb = k0; kb < k1; kb += U) {
    if (!ok) break;
    const int Kc0 = (kb - baseC) / WC;
#pragma unroll
    for (int s = 0; s < U; ++s) {
      const int k = kb + s, Kc = Kc0 + s / WC;
      const int sC = s % WC, sA = s % WA;
      op.column(s, k, Kc, col, a, ginv);
      const double vk1 = (sC < WC - 1) ? __shfl_sync(FULLMASK, v[(sC + 1) % WC], Kc) : __shfl_sync(FULLMASK, v[0], (Kc + 1) & 31);
      const double rho = (k + 1 < n) ? vk1 * ginv : 0.0;
      const double e = fma(-rho, rho, 1.0);
      if (!(e > 0.0)) ok = false;
      ginv *= rcp_nr(ok ? e : 1.0);             // g_{k+1} = g_k e
      col[(2 * WC - 1 - sC) % WC] = __shfl_up_sync(FULLMASK, col[(2 * WC - 1 - sC) % WC], 1);
#pragma unroll
      for (int q = 0; q < WC; ++q) {
        const int pc = (q - sC - 1 + 2 * WC) % WC;
        const double un = fma(-rho, v[q], col[pc]);
        v[q] = fma(-rho, un, e * v[q]);
        col[pc] = un;
      }
      if constexpr (INV) {
        const double up = __shfl_up_sync(FULLMASK, a[(2 * WA - 1 - sA) % WA], 1);
        a[(2 * WA - 1 - sA) % WA] = (lane == 0) ? 0.0 : up;       // position 0 of the shifted column is an exact zero
#pragma unroll
        for (int q = 0; q < WA; ++q) {
          const int pa = (q - sA - 1 + 2 * WA) % WA;
          const double an = fma(-rho, b[q], a[pa]);
          b[q] = fma(-rho, an, e * b[q]);
          a[pa] = an;
        }
      }
    }
  }
}

// the inverse set at twice the width: lanes 0..15 take over the slots of two old lanes each, lanes 16..31 start at zero
template <int W>
__device__ __forceinline__ void sfn_expand(const double (&in)[W], double (&out)[2 * W], int lane) {
#pragma unroll
  for (int q = 0; q < 2 * W; ++q) {
    const double val = __shfl_sync(FULLMASK, in[q % W], (2 * lane + (q >= W ? 1 : 0)) & 31);
    out[q] = (lane < 16) ? val : 0.0;
  }
}

// f_j += L[:, k] z1_j[k]  with L[p][k] = u[p] / sqrt(g_k)  (positions above the diagonal of the column hold dead values:
// masked); z1 in shared memory
template <int WC, int NGT>
struct OpPriorDraw {
  int nld, lane;
  const double* zs;               // NGT x nld (rows of absent functions are zero)
  double (&f)[NGT][WC];
  __device__ __forceinline__ void column(int s, int k, int Kc, const double (&col)[WC], const double (&)[1], double ginv) {
    const int sC = s % WC;
    const double su = ginv * rsqrt_nr(ginv);      // 1 / L[k][k]
#pragma unroll
    for (int j = 0; j < NGT; ++j) {
      const double zk = zs[j * nld + k] * su;
#pragma unroll
      for (int q = 0; q < WC; ++q) {
        const bool live = (q >= sC) ? (lane >= Kc) : (lane > Kc);
        if (live) f[j][q] = fma(col[(q - sC + WC) % WC], zk, f[j][q]);
      }
    }
  }
};

// one column of the solve  v = C^-1 q  for up to SFN_NC right-hand sides:  wt = r_k / g_k,  r -= u wt,  v += a wt
// (a = column k of L^-T times the scale of u; it is exactly zero below position k: no masking needed)
template <int WC, int WA, int NM>
struct OpSolve {
  int n;                          // time points
  double (&r)[NM][WC];
  double (&vo)[NM][WA];
  __device__ __forceinline__ void column(int s, int k, int Kc, const double (&col)[WC], const double (&a)[WA], double ginv) {
    const int sC = s % WC, sA = s % WA;
#pragma unroll
    for (int j = 0; j < NM; ++j) {
      const double rraw = __shfl_sync(FULLMASK, r[j][sC], Kc);
      const double wt = (k < n) ? rraw * ginv : 0.0;
#pragma unroll
      for (int q = 0; q < WC; ++q) r[j][q] = fma(-col[(q - sC + WC) % WC], wt, r[j][q]);
#pragma unroll
      for (int q = 0; q < WA; ++q) vo[j][q] = fma(a[(q - sA + WA) % WA], wt, vo[j][q]);
    }
  }
};

// prior draws: the column set through its stages; finished positions of f (below the current column) are parked in the
// shared-memory rows of z1 they no longer need
template <int W0, int WC, int NGT>
__device__ __forceinline__ void sfn_prior_chain(int n, int n8, int lane, double (&col)[WC], double (&v)[WC], double (&f)[NGT][WC],
                                                double& ginv, bool& ok, int nld, double* zs) {
  constexpr int baseC = 32 * (W0 - WC);
  constexpr int K1 = (WC > 1) ? baseC + 16 * WC : 32 * W0;
  double da[1] = {0.0}, db[1] = {0.0};
  {
    OpPriorDraw<WC, NGT> op{nld, lane, zs, f};
    sfn_segment<WC, 1, false>(n, baseC, (K1 < n8) ? K1 : n8, baseC, lane, col, v, da, db, ginv, ok, op);
  }
  if constexpr (WC > 1) {
    if (K1 < n8 && ok) {
      __syncwarp();
      if (lane < 16) {
#pragma unroll
        for (int j = 0; j < NGT; ++j)
#pragma unroll
          for (int q = 0; q < WC; ++q) zs[j * nld + baseC + WC * lane + q] = f[j][q];
      }
      double col2[WC / 2], v2[WC / 2], f2[NGT][WC / 2];
      schur_repack<WC>(col, col2, lane);
      schur_repack<WC>(v, v2, lane);
#pragma unroll
      for (int j = 0; j < NGT; ++j) schur_repack<WC>(f[j], f2[j], lane);
      sfn_prior_chain<W0, WC / 2, NGT>(n, n8, lane, col2, v2, f2, ginv, ok, nld, zs);
      return;
    }
  }
  __syncwarp();
#pragma unroll
  for (int j = 0; j < NGT; ++j)
#pragma unroll
    for (int q = 0; q < WC; ++q) {
      const int p = baseC + WC * lane + q;
      if (p < nld) zs[j * nld + p] = f[j][q];
    }
}

// the embedded recursion through the stages of both sets (compile-time schedule: K0 = first step of this segment)
template <int W0, int WC, int WA, int K0, int NM>
__device__ __forceinline__ void sfn_solve_chain(int n, int n8, int lane, double (&col)[WC], double (&v)[WC], double (&a)[WA], double (&b)[WA],
                                                double (&rq)[NM][WC], double (&vo)[NM][WA], double& ginv, bool& ok, double (&vout)[NM][W0]) {
  constexpr int baseC = 32 * (W0 - WC);
  constexpr int kc = (WC > 1) ? baseC + 16 * WC : (1 << 30);      // the column set is re-dealt here
  constexpr int ka = (WA < W0) ? 32 * WA - 8 : (1 << 30);         // the inverse set is spread out here
  constexpr int kx = (kc < ka) ? kc : ka;
  constexpr int K1 = (kx < 32 * W0) ? kx : 32 * W0;
  {
    OpSolve<WC, WA, NM> op{n, rq, vo};
    sfn_segment<WC, WA, true>(n, K0, (K1 < n8) ? K1 : n8, baseC, lane, col, v, a, b, ginv, ok, op);
  }
  if constexpr (K1 < 32 * W0) {
    if (K1 < n8 && ok) {
      if constexpr (ka < kc) {
        double a2[2 * WA], b2[2 * WA], vo2[NM][2 * WA];
        sfn_expand<WA>(a, a2, lane);
        sfn_expand<WA>(b, b2, lane);
#pragma unroll
        for (int j = 0; j < NM; ++j) sfn_expand<WA>(vo[j], vo2[j], lane);
        sfn_solve_chain<W0, WC, 2 * WA, K1, NM>(n, n8, lane, col, v, a2, b2, rq, vo2, ginv, ok, vout);
      } else {
        double col2[WC / 2], v2[WC / 2], rq2[NM][WC / 2];
        schur_repack<WC>(col, col2, lane);
        schur_repack<WC>(v, v2, lane);
#pragma unroll
        for (int j = 0; j < NM; ++j) schur_repack<WC>(rq[j], rq2[j], lane);
        sfn_solve_chain<W0, WC / 2, WA, K1, NM>(n, n8, lane, col2, v2, a, b, rq2, vo, ginv, ok, vout);
      }
      return;
    }
  }
#pragma unroll
  for (int j = 0; j < NM; ++j)
#pragma unroll
    for (int q = 0; q < W0; ++q) {
      if constexpr (WA == W0) vout[j][q] = vo[j][q];
      else vout[j][q] = 0.0;                      // only reached after a breakdown (n > 16 W0 otherwise passes every expansion)
    }
}

// elements e0 + W lane + q of a function's normal vector, valid below n_valid_to
template <int W>
__device__ __forceinline__ void sfn_normals(const Stream& rng, unsigned gch, unsigned sweep, unsigned sid, int e0, int n_valid_to, int lane,
                                            double (&out)[W]) {
#pragma unroll
  for (int q = 0; q < W; ++q) out[q] = 0.0;
  const int base = e0 + W * lane;
  if (base >= n_valid_to) return;
  if (W > 1 && (base & 1) == 0) {
#pragma unroll
    for (int q = 0; q < W; q += 2) {
      const int e = base + q;
      if (e < n_valid_to) {
        double za, zb;
        rng.normal_pair(gch, sweep, sid, (unsigned)(e >> 1), za, zb);
        out[q] = za;
        if (q + 1 < W && e + 1 < n_valid_to) out[q + 1] = zb;
      }
    }
  } else {
#pragma unroll
    for (int q = 0; q < W; ++q) {
      const int e = base + q;
      if (e < n_valid_to) out[q] = normal_elem(rng, gch, sweep, sid, e);
    }
  }
}

// prior draws of NGT functions (rows of zs: z1 in, f = L_B z1 out); false on breakdown
template <int W0, int NGT>
__device__ __forceinline__ bool sfn_prior(const CtxDev& c, int n, int nld, int lane, double sigma, double ls, double diag_add, double* zs) {
  double col[W0], v[W0], f[NGT][W0];
  bool ok;
  double ginv = sfn_init<W0>(c, n, sigma, ls, 1.0, diag_add, col, v, lane, ok);
#pragma unroll
  for (int j = 0; j < NGT; ++j)
#pragma unroll
    for (int q = 0; q < W0; ++q) f[j][q] = 0.0;
  if (ok) sfn_prior_chain<W0, W0, NGT>(n, nld, lane, col, v, f, ginv, ok, nld, zs);
  return ok;
}

// v = C^-1 q for NM right-hand sides (rows mem[0..NM) of zs) by the embedded recursion; false on breakdown
template <int W0, int NM>
__device__ __forceinline__ bool sfn_solve(const CtxDev& c, int n, int nld, int lane, double sigma, double ls, double cc, double diag_add,
                                          const double* zs, const int* mem, double (&vout)[SFN_NC][W0]) {
  double rq[NM][W0], vo[NM][1], vres[NM][W0];
#pragma unroll
  for (int j = 0; j < NM; ++j) {
#pragma unroll
    for (int q = 0; q < W0; ++q) rq[j][q] = (W0 * lane + q < nld) ? zs[mem[j] * nld + W0 * lane + q] : 0.0;
    vo[j][0] = 0.0;
  }
  double col[W0], v[W0], ia[1], ib[1];
  bool ok;
  double ginv = sfn_init<W0>(c, n, sigma, ls, cc, diag_add, col, v, lane, ok);
  ia[0] = ib[0] = (lane == 0) ? 1.0 : 0.0;          // generator rows [e0 e0] (times the scale sqrt(t0) of the column set)
  if (ok) sfn_solve_chain<W0, W0, 1, 0, NM>(n, nld, lane, col, v, ia, ib, rq, vo, ginv, ok, vres);
  if (ok) {
#pragma unroll
    for (int j = 0; j < NM; ++j)
#pragma unroll
      for (int q = 0; q < W0; ++q) vout[j][q] = vres[j][q];
  }
  return ok;
}

// one (prior group, chain) item at W0 slots per lane
template <int W0>
__device__ __forceinline__ void sfn_item(const CtxDev& c, const SchurFnArgs& a, const Stream& rng, int ch, int g, int lane, double* zs) {
  const int n = c.n, nld = ((n + 7) >> 3) << 3;
  const unsigned gch = (unsigned)(ch + c.chain_offset);
  const double sigma = exp10(__ldcg(c.hyper + (size_t)ch * c.H + 1 + 2 * g));      // (L2: written by other SMs in the sweep kernel)
  const double ls = exp10(__ldcg(c.hyper + (size_t)ch * c.H + 2 + 2 * g));
  const double yinv = 1.0 / (exp10(__ldcg(c.hyper + (size_t)ch * c.H)) + c.offset);
  const int f0 = c.group_off[g], ng = c.group_off[g + 1] - f0;          // functions of the group: group_fns[f0 .. f0+ng)
  const int* gf = c.group_fns + f0;
  // sampler id of every function of the group (from the class lists) and its class's c
  unsigned sid[SFN_NG];
  double ccls[SFN_NG];
#pragma unroll
  for (int j = 0; j < SFN_NG; ++j) { sid[j] = 0u; ccls[j] = 1.0; }
  for (int cls = a.gcls_off[g]; cls < a.gcls_off[g + 1]; ++cls)
    for (int t = a.cls_off[cls]; t < a.cls_off[cls + 1]; ++t) {
      const int fn = a.cls_fn[t];
#pragma unroll
      for (int j = 0; j < SFN_NG; ++j)
        if (j < ng && gf[j] == fn) { sid[j] = a.cls_sid[t]; ccls[j] = c.G[(size_t)fn * c.f + fn] * yinv; }
    }
  // ---------------- 1. prior draws through the columns of chol(K_o), jitter ladder on breakdown ----------------
  int tries = 0, code = 0;
  double jitter = 0.0;
  for (;;) {
    __syncwarp();
#pragma unroll
    for (int j = 0; j < SFN_NG; ++j) {             // z1 (regenerated on a retry: finished draws overwrite it); absent rows zero
      double z[W0];
      if (j < ng) sfn_normals<W0>(rng, gch, a.sweep, sid[j], 0, n, lane, z);
      else {
#pragma unroll
        for (int q = 0; q < W0; ++q) z[q] = 0.0;
      }
#pragma unroll
      for (int q = 0; q < W0; ++q)
        if (W0 * lane + q < nld) zs[j * nld + W0 * lane + q] = z[q];
    }
    __syncwarp();
    bool ok;
    if (ng == 1) ok = sfn_prior<W0, 1>(c, n, nld, lane, sigma, ls, c.offset + jitter, zs);
    else if (ng == 2) ok = sfn_prior<W0, 2>(c, n, nld, lane, sigma, ls, c.offset + jitter, zs);
    else ok = sfn_prior<W0, SFN_NG>(c, n, nld, lane, sigma, ls, c.offset + jitter, zs);
    if (ok) { code = tries; break; }
    const double dm = sigma + c.offset;                // jitchol ladder (linalg.py:41-54)
    if (tries == 0) {
      if (!(dm > 0.0)) { code = GPF_STATUS_NONPOS_DIAG; break; }
      jitter = dm * 1e-6;
    } else jitter *= 10.0;
    ++tries;
    if (tries > c.maxtries || !isfinite(jitter)) { code = GPF_STATUS_NOT_PD; break; }
  }
  int nfact = 0, ndraw = 0;
  if (code >= 0) {
    // ---------------- 2. q_j = sqrt(c_j) (ybar_j - f_j) - z2_j, parked in shared memory ----------------
    __syncwarp();
#pragma unroll
    for (int j = 0; j < SFN_NG; ++j)
      if (j < ng) {
        double z2[W0];
        sfn_normals<W0>(rng, gch, a.sweep, sid[j], n, 2 * n, lane, z2);
        const double* yb = c.bhat + (size_t)gf[j] * n;
        const double sc = sqrt(ccls[j]);
#pragma unroll
        for (int q = 0; q < W0; ++q) {
          const int i = W0 * lane + q;
          if (i < nld) zs[j * nld + i] = (i < n) ? sc * (yb[i] - zs[j * nld + i]) - z2[q] : 0.0;
        }
      }
    __syncwarp();
    // ---------------- 3. per class, two functions per pass: v = C^-1 q by the embedded recursion ----------------
    for (int cls = a.gcls_off[g]; cls < a.gcls_off[g + 1] && code >= 0; ++cls) {
      int mem[SFN_NG], nmem = 0;
      double cc = 1.0;
      for (int t = a.cls_off[cls]; t < a.cls_off[cls + 1]; ++t) {
        const int fn = a.cls_fn[t];
#pragma unroll
        for (int j = 0; j < SFN_NG; ++j)
          if (j < ng && gf[j] == fn) { mem[nmem++] = j; cc = ccls[j]; }
      }
      const double isc = 1.0 / sqrt(cc);
      for (int m0 = 0; m0 < nmem; m0 += SFN_NC) {
        const int nm = (nmem - m0 < SFN_NC) ? nmem - m0 : SFN_NC;
        double vout[SFN_NC][W0];
        int mm[SFN_NC];
#pragma unroll
        for (int j = 0; j < SFN_NC; ++j) {
          mm[j] = mem[(m0 + j < nmem) ? m0 + j : m0];
#pragma unroll
          for (int q = 0; q < W0; ++q) vout[j][q] = 0.0;
        }
        const double dadd = 1.0 + cc * (c.offset + jitter);
        const bool ok = (nm == 1) ? sfn_solve<W0, 1>(c, n, nld, lane, sigma, ls, cc, dadd, zs, mm, vout)
                                  : sfn_solve<W0, SFN_NC>(c, n, nld, lane, sigma, ls, cc, dadd, zs, mm, vout);
        ++nfact;
        if (!ok) { code = GPF_STATUS_NOT_PD; break; }
        // ---------------- 4. beta_j = ybar_j - (z2_j + v_j) / sqrt(c) ----------------
#pragma unroll
        for (int j = 0; j < SFN_NC; ++j)
          if (j < nm) {
            const int gj = mm[j];
            double z2[W0];
            sfn_normals<W0>(rng, gch, a.sweep, sid[gj], n, 2 * n, lane, z2);
            const double* yb = c.bhat + (size_t)gf[gj] * n;
            double* Fo = c.F + ((size_t)ch * c.f + gf[gj]) * n;
#pragma unroll
            for (int q = 0; q < W0; ++q) {
              const int i = W0 * lane + q;
              if (i < n) Fo[i] = yb[i] - (z2[q] + vout[j][q]) * isc;
            }
          }
        ndraw += nm;
      }
    }
  }
  if (lane == 0) {
    status_merge(c.status + ch, code);
    atomicAdd(c.counters + CNT_CHOL, 1ull + (unsigned long long)tries + (unsigned long long)nfact);
    atomicAdd(c.counters + CNT_BFACT, 1ull + (unsigned long long)tries);
    atomicAdd(c.counters + CNT_FNCHOL, (unsigned long long)nfact);
    atomicAdd(c.counters + CNT_DRAW, (unsigned long long)ndraw);
    if (tries) atomicAdd(c.counters + CNT_JITTER, (unsigned long long)tries);
  }
  __syncwarp();
}

#ifndef GPF_NO_SCHURFN_KERNEL      // (translation units that only need sfn_item: gpf_sweepk_tu.cu)
#ifndef GPF_SFN_WPB
#define GPF_SFN_WPB 4              // warps per CTA (12 warps per SM by registers: 12 / GPF_SFN_WPB CTAs per SM)
#endif
__global__ void __launch_bounds__(32 * GPF_SFN_WPB, 12 / GPF_SFN_WPB) k_group_schurfn(CtxDev c, SchurFnArgs a) {
  extern __shared__ __align__(16) double gpf_smem_sfn[];      // per warp: SFN_SMEM_ROWS x nld (z1 / finished prior draws, then the right-hand sides q)
  const int n = c.n, nld = ((n + 7) >> 3) << 3;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const Stream rng{c.k0, c.k1};
  double* zs = gpf_smem_sfn + (size_t)warp * (SFN_SMEM_ROWS * nld);
  // Items are handed out GPF_SFN_WPB AT A TIME to a CTA (consecutive chains of one prior group): its warps then run the same
  // recursion loops side by side.  The loops of this kernel add up to several times the 32 KB instruction cache; with warps
  // taking items one by one every warp of an SM sat in a different loop (ncu: 87 % instruction-cache hits, 0.8 no-instruction
  // stall cycles per issue).
  __shared__ int s_ticket;
  const int nitem = c.C * (a.ngroups > 0 ? a.ngroups : c.P);
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) s_ticket = atomicAdd(c.sched, 1);
    __syncthreads();
    const int base = s_ticket * GPF_SFN_WPB;
    if (base >= nitem) break;
    const int item = (int)__reduce_max_sync(FULLMASK, (unsigned)(base + warp));
    if (item >= nitem) continue;
    const int ch = item % c.C, g = a.gorder[item / c.C];
    if (n > 128) sfn_item<8>(c, a, rng, ch, g, lane, zs);
    else if (n > 64) sfn_item<4>(c, a, rng, ch, g, lane, zs);
    else if (n > 32) sfn_item<2>(c, a, rng, ch, g, lane, zs);
    else sfn_item<1>(c, a, rng, ch, g, lane, zs);
  }
}

#endif

}  // namespace gpf
