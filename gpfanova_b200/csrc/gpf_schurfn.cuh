// Function._sample (sample/function.py:14-40) on a uniform time grid in O(n^2) per function: the Schur algorithm instead
// of dense factorisations.
//
// Balanced design with nothing missing on a uniform grid: both matrices of the observation form (gpf_cond.cuh)
//     K_o = K + (offset + jitter) I          and          C = I + c K_o,   c = n_f / (sigma_y + offset),
// are symmetric positive definite TOEPLITZ matrices, so the Schur algorithm (gpf_schur.cuh) yields the columns of their
// Cholesky factors one per step with one hyperbolic rotation and O(n) work -- no n x n storage, no n^3.  One WARP does one
// (prior group, chain) item, every vector register-resident (lane l owns positions 8l .. 8l+7, n <= 256, at most four
// functions per group):
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
#include "gpf_cond.cuh"

namespace gpf {

constexpr int SFN_NG = 4;          // functions per prior group held in registers

struct SchurFnArgs {
  const int* gorder;      // P groups, largest first
  const int* gcls_off;    // P+1
  const int* cls_off;     // CSR offsets of the classes into cls_fn / cls_sid
  const int* cls_fn;
  const unsigned* cls_sid;
  unsigned sweep;
};

// first column of  scale * (K + add0 I)  [+ 1 on the diagonal when plus_one]: registers col/v of the Schur generator
// (lane owns positions 8 lane + q).  Returns 1 / L[0][0], or a NaN-free 0 with ok = false if the pivot is not positive.
__device__ __forceinline__ double schurfn_init(const CtxDev& c, int n, double sigma, double ls, double scale, double diag_add,
                                               double (&col)[8], double (&v)[8], int lane, bool& ok) {
  const double sc = -0.5 / (ls * ls);
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int i = lane * 8 + q;
    col[q] = (i < n) ? scale * sigma * exp(c.d2[i] * sc) : 0.0;
  }
  const double t0 = __shfl_sync(FULLMASK, col[0], 0) + diag_add;
  ok = (t0 > 0.0) && (t0 < 1.0e300);
  const double uinv = ok ? rsqrt_nr(t0) : 0.0;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int i = lane * 8 + q;
    const double u = ((i == 0) ? t0 : col[q]) * uinv;
    col[q] = u;
    v[q] = (i == 0) ? 0.0 : u;
  }
  return uinv;
}

// the hyperbolic rotation that turns column k into column k+1 (step s of block K); returns false on breakdown.
// INV: the same shift + rotation on the inverse-factor rows (a, b): a becomes column k+1 of L^-T.
template <int s, bool INV>
__device__ __forceinline__ bool schurfn_rotate(int n, int K, int lane, double (&col)[8], double (&v)[8], double (&a)[8], double (&b)[8],
                                               double& uinv) {
  const int k = 8 * K + s;
  const double vk1 = (s < 7) ? __shfl_sync(FULLMASK, v[(s + 1) & 7], K) : __shfl_sync(FULLMASK, v[0], (K + 1) & 31);
  const double rho = (k + 1 < n) ? vk1 * uinv : 0.0;
  const double e = fma(-rho, rho, 1.0);
  const bool ok = e > 0.0;
  const double cinv = rsqrt_nr(ok ? e : 1.0);
  const double cc = e * cinv;
  uinv *= cinv;                             // L[k+1][k+1] = L[k][k] c
  // shift the column down by one position: only the register holding logical position 7 crosses lanes
  col[(7 - s) & 7] = __shfl_up_sync(FULLMASK, col[(7 - s) & 7], 1);
  if constexpr (INV) {
    const double up = __shfl_up_sync(FULLMASK, a[(7 - s) & 7], 1);
    a[(7 - s) & 7] = (lane == 0) ? 0.0 : up;           // position 0 of the shifted column is an exact zero
  }
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int pc = (q - s - 1) & 7;
    const double un = (col[pc] - rho * v[q]) * cinv;
    v[q] = fma(cc, v[q], -rho * un);
    col[pc] = un;
    if constexpr (INV) {
      const double an = (a[pc] - rho * b[q]) * cinv;
      b[q] = fma(cc, b[q], -rho * an);
      a[pc] = an;
    }
  }
  return ok;
}

template <int s, bool INV, class Op>
__device__ __forceinline__ bool schurfn_block_step(int n, int K, int lane, double (&col)[8], double (&v)[8], double (&a)[8],
                                                   double (&b)[8], double& uinv, Op& op) {
  op.template column<s>(K, col, a, uinv);   // column k = 8K + s: logical position q of this lane is col[(q - s) & 7] (a likewise)
  return schurfn_rotate<s, INV>(n, K, lane, col, v, a, b, uinv);
}

// run the recursion over all columns; op.column<s>(K, col, a, uinv) sees every column before it is rotated away
template <bool INV, class Op>
__device__ __forceinline__ bool schurfn_run(int n, int lane, double (&col)[8], double (&v)[8], double (&a)[8], double (&b)[8], double uinv,
                                            Op& op) {
  bool ok = true;
  const int nblk = (n + 7) >> 3;
  for (int K = 0; K < nblk && ok; ++K) {
    ok = schurfn_block_step<0, INV>(n, K, lane, col, v, a, b, uinv, op) && ok;
    ok = schurfn_block_step<1, INV>(n, K, lane, col, v, a, b, uinv, op) && ok;
    ok = schurfn_block_step<2, INV>(n, K, lane, col, v, a, b, uinv, op) && ok;
    ok = schurfn_block_step<3, INV>(n, K, lane, col, v, a, b, uinv, op) && ok;
    ok = schurfn_block_step<4, INV>(n, K, lane, col, v, a, b, uinv, op) && ok;
    ok = schurfn_block_step<5, INV>(n, K, lane, col, v, a, b, uinv, op) && ok;
    ok = schurfn_block_step<6, INV>(n, K, lane, col, v, a, b, uinv, op) && ok;
    ok = schurfn_block_step<7, INV>(n, K, lane, col, v, a, b, uinv, op) && ok;
  }
  return ok;
}

// f_j += L[:, k] z1_j[k]   (positions above the diagonal of the column hold dead values: masked); z1 in shared memory
struct OpPriorDraw {
  int n, nld, lane, ng;
  const double* zs;               // ng x nld
  double (&f)[SFN_NG][8];
  template <int s>
  __device__ __forceinline__ void column(int K, const double (&col)[8], const double (&)[8], double) {
    const int k = 8 * K + s;
#pragma unroll
    for (int j = 0; j < SFN_NG; ++j) {
      if (j < ng) {
        const double zk = zs[j * nld + k];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const bool live = (q >= s) ? (lane >= K) : (lane > K);
          if (live) f[j][q] = fma(col[(q - s) & 7], zk, f[j][q]);
        }
      }
    }
  }
};

// one column of the solve  v = C^-1 q  for up to SFN_NC right-hand sides: y_k = r_k / L[k][k], r -= L[:, k] y_k,
// v += L^-T[:, k] y_k   (L^-T[:, k] is exactly zero below position k: no masking needed)
constexpr int SFN_NC = 2;
struct OpSolve {
  int n, nm;                      // time points, right-hand sides in this pass
  double (&r)[SFN_NC][8];
  double (&vo)[SFN_NC][8];
  template <int s>
  __device__ __forceinline__ void column(int K, const double (&col)[8], const double (&a)[8], double uinv) {
    const int k = 8 * K + s;
#pragma unroll
    for (int j = 0; j < SFN_NC; ++j) {
      if (j < nm) {
        const double w = (k < n) ? __shfl_sync(FULLMASK, r[j][s], K) * uinv : 0.0;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          r[j][q] = fma(-col[(q - s) & 7], w, r[j][q]);
          vo[j][q] = fma(a[(q - s) & 7], w, vo[j][q]);
        }
      }
    }
  }
};

__device__ __forceinline__ void sfn_normals8(const Stream& rng, unsigned gch, unsigned sweep, unsigned sid, int e0, int n_valid_from,
                                             int n_valid_to, int lane, double (&out)[8]) {
  // elements e0 + 8 lane + q of the function's normal vector; valid element range [n_valid_from, n_valid_to)
#pragma unroll
  for (int q = 0; q < 8; ++q) out[q] = 0.0;
  const int base = e0 + 8 * lane;
  if (base >= n_valid_to) return;
  if ((base & 1) == 0) {
#pragma unroll
    for (int q = 0; q < 8; q += 2) {
      const int e = base + q;
      if (e < n_valid_to) {
        double za, zb;
        rng.normal_pair(gch, sweep, sid, (unsigned)(e >> 1), za, zb);
        out[q] = za;
        if (e + 1 < n_valid_to) out[q + 1] = zb;
      }
    }
  } else {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int e = base + q;
      if (e < n_valid_to) out[q] = normal_elem(rng, gch, sweep, sid, e);
    }
  }
  (void)n_valid_from;
}

#ifndef GPF_SFN_MINB
#define GPF_SFN_MINB 3
#endif
__global__ void __launch_bounds__(128, GPF_SFN_MINB) k_group_schurfn(CtxDev c, SchurFnArgs a) {
  extern __shared__ __align__(16) double gpf_smem_sfn[];      // per warp: SFN_NG x nld (z1 of the group's functions, then their q)
  const int n = c.n, nld = ((n + 7) >> 3) << 3;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const Stream rng{c.k0, c.k1};
  double* zs = gpf_smem_sfn + (size_t)warp * (SFN_NG * nld);
  for (int item = next_chain_warp(c.sched, lane); item < c.C * c.P; item = next_chain_warp(c.sched, lane)) {
    const int ch = item % c.C, g = a.gorder[item / c.C];
    const unsigned gch = (unsigned)(ch + c.chain_offset);
    const double sigma = exp10(c.hyper[(size_t)ch * c.H + 1 + 2 * g]);
    const double ls = exp10(c.hyper[(size_t)ch * c.H + 2 + 2 * g]);
    const double yinv = 1.0 / (exp10(c.hyper[(size_t)ch * c.H]) + c.offset);
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
    double r[SFN_NG][8];           // f, then q, then y, then v
    int tries = 0, code = 0;
    double jitter = 0.0;
    {
      __syncwarp();
      for (int j = 0; j < ng; ++j) {
        double z[8];
        sfn_normals8(rng, gch, a.sweep, sid[j], 0, 0, n, lane, z);
        if (8 * lane < nld) {
#pragma unroll
          for (int q = 0; q < 8; ++q) zs[j * nld + 8 * lane + q] = z[q];
        }
      }
      __syncwarp();
      for (;;) {
        double col[8], v[8], da[8], db[8];      // (da, db unused without the inverse factor)
        bool ok;
        double uinv = schurfn_init(c, n, sigma, ls, 1.0, c.offset + jitter, col, v, lane, ok);
#pragma unroll
        for (int j = 0; j < SFN_NG; ++j)
#pragma unroll
          for (int q = 0; q < 8; ++q) r[j][q] = 0.0;
        if (ok) {
          OpPriorDraw op{n, nld, lane, ng, zs, r};
          ok = schurfn_run<false>(n, lane, col, v, da, db, uinv, op);
        }
        if (ok) { code = tries; break; }
        const double dm = sigma + c.offset;                // jitchol ladder (linalg.py:41-54)
        if (tries == 0) {
          if (!(dm > 0.0)) { code = GPF_STATUS_NONPOS_DIAG; break; }
          jitter = dm * 1e-6;
        } else jitter *= 10.0;
        ++tries;
        if (tries > c.maxtries || !isfinite(jitter)) { code = GPF_STATUS_NOT_PD; break; }
      }
    }
    int nfact = 0, ndraw = 0;
    if (code >= 0) {
      // ---------------- 2. q_j = sqrt(c_j) (ybar_j - f_j) - z2_j, parked in shared memory ----------------
      __syncwarp();
#pragma unroll
      for (int j = 0; j < SFN_NG; ++j)
        if (j < ng) {
          double z2[8];
          sfn_normals8(rng, gch, a.sweep, sid[j], n, n, 2 * n, lane, z2);
          const double* yb = c.bhat + (size_t)gf[j] * n;
          const double sc = sqrt(ccls[j]);
          if (8 * lane < nld) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const int i = 8 * lane + q;
              zs[j * nld + i] = (i < n) ? sc * (yb[i] - r[j][q]) - z2[q] : 0.0;
            }
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
          double rq[SFN_NC][8], vo[SFN_NC][8];
#pragma unroll
          for (int j = 0; j < SFN_NC; ++j)
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              rq[j][q] = (j < nm && 8 * lane < nld) ? zs[mem[m0 + j] * nld + 8 * lane + q] : 0.0;
              vo[j][q] = 0.0;
            }
          double col[8], v[8], ia[8], ib[8];
          bool ok;
          double uinv = schurfn_init(c, n, sigma, ls, cc, 1.0 + cc * (c.offset + jitter), col, v, lane, ok);
#pragma unroll
          for (int q = 0; q < 8; ++q) ia[q] = ib[q] = (lane == 0 && q == 0) ? uinv : 0.0;      // generator rows [e0 e0] / sqrt(t0)
          if (ok) {
            OpSolve op{n, nm, rq, vo};
            ok = schurfn_run<true>(n, lane, col, v, ia, ib, uinv, op);
          }
          ++nfact;
          if (!ok) { code = GPF_STATUS_NOT_PD; break; }
          // ---------------- 4. beta_j = ybar_j - (z2_j + v_j) / sqrt(c) ----------------
#pragma unroll
          for (int j = 0; j < SFN_NC; ++j)
            if (j < nm) {
              const int gj = mem[m0 + j];
              double z2[8];
              sfn_normals8(rng, gch, a.sweep, sid[gj], n, n, 2 * n, lane, z2);
              const double* yb = c.bhat + (size_t)gf[gj] * n;
              double* Fo = c.F + ((size_t)ch * c.f + gf[gj]) * n;
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const int i = 8 * lane + q;
                if (i < n) Fo[i] = yb[i] - (z2[q] + vo[j][q]) * isc;
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
}

}  // namespace gpf
