// Shared device-side state and helpers of the gpfanova_b200 kernels.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "gpf_philox.cuh"
#include "gpf_tile.cuh"

namespace gpf {

constexpr double LOG2PI = 1.8378770664093453;   // log(2 pi)
constexpr double LN10 = 2.302585092994046;

// counters (gpf_get_counters)
enum { CNT_CHOL = 0, CNT_LOGP = 1, CNT_JITTER = 2, CNT_DRAW = 3, CNT_INV = 4, CNT_SLICE = 5, CNT_FNCHOL = 6, CNT_BFACT = 7, CNT_N = 8 };

// device view of one sampler context (passed by value to every kernel)
struct CtxDev {
  int n, p, m, f, P, C, H, NT, NEmax, chain_offset, maxtries;
  int centro;                 // uniform grid, n even and >= 64: K_inv through the centrosymmetric split (two n/2 inverses)
  const double* x;            // n x p
  const double* yT;           // m x n (transposed observations, NaN = missing)
  const double* X;            // m x f
  const int* group_of_fn;     // f
  const int* group_fns;       // functions of every group, concatenated
  const int* group_off;       // P+1 offsets into group_fns
  double* F;                  // C x f x n
  double* hyper;              // C x H
  int* status;                // C
  unsigned long long* counters;
  double* Kinv;               // C x plt_tiles(NT) tiles: scratch of the K_inv operator (gpf_group_kinv with output), else null
  double* LB;                 // P x C x lb_stride: Cholesky factor(s) of K_g + offset I of every chain (sequential sweeps);
                              // centro: the factors of the two half-size blocks, one after the other
  size_t lb_stride;
  double* Bjit;               // P x C: jitter the ladder of linalg.py:41-54 added to K_g + offset I
  double* ws;                 // per-CTA workspace, ws_stride doubles each
  size_t ws_stride;
  double offset, prior_jitter, prior_lb, prior_ub;
  unsigned k0, k1;            // Philox key
  const double* d2;           // uniform grid (p = 1): squared distance at lag k, n entries (K is Toeplitz); else null
  const double* G;            // f x f   X^T X
  const double* Z;            // f x n   X^T y over the observed entries (NaN -> 0)
  const unsigned char* full;  // n       1 if no observation is missing at time t
  int* sched;                 // chain hand-out counter of the running launch (reset to 0 by the host before it)
  double* fws;                // per-CTA scratch of the fused function kernel (right-hand-side rows): fws_stride doubles each
  size_t fws_stride;
  const double* bhat;         // f x n  least-squares functions Z[f,t] / G[f,f] (balanced design, nothing missing), else null
  double rss0;                // sum of squared residuals of y about X bhat (host, extended precision)
};

// dynamic hand-out of chains to CTAs (chains differ in the number of slice evaluations they need): returns the next
// chain index to every thread; >= C when the launch is exhausted.  slot: one int of shared memory.
__device__ __forceinline__ int next_chain(int* sched, int* slot) {
  __syncthreads();
  if (threadIdx.x == 0) *slot = atomicAdd(sched, 1);
  __syncthreads();
  return *slot;
}

// per-chain status word: a negative code (not PD) wins over everything, the more severe (more negative) first; otherwise
// the largest number of jitter retries.  Several CTAs (items of the same chain in one launch, concurrent slice kernels of
// different groups) may report at once: compare-and-swap loop, no update is lost.
__device__ __forceinline__ void status_merge(int* st, int code) {
  int old = *reinterpret_cast<volatile int*>(st);
  for (;;) {
    int want;
    if (code < 0) want = (old >= 0 || code < old) ? code : old;
    else want = (old >= 0 && code > old) ? code : old;
    if (want == old) return;
    const int seen = atomicCAS(st, old, want);
    if (seen == old) return;
    old = seen;
  }
}

// scaled inputs of RBF.dist (rbf.pyx:9-11): xs = x / lengthscale, xsq = sum(xs^2, 1)
__device__ __forceinline__ void rbf_scale_inputs(const double* __restrict__ x, int n, int p, double ls, double* xs,
                                                 double* xsq) {
  for (int t = threadIdx.x; t < n; t += blockDim.x) {
    double s = 0.0;
    for (int d = 0; d < p; ++d) {
      const double v = x[t * p + d] / ls;
      xs[t * p + d] = v;
      s += v * v;
    }
    xsq[t] = s;
  }
}

// one element of RBF._K (rbf.pyx:12-19); outside the n x n block the tile padding is the identity
__device__ __forceinline__ double rbf_elem(const double* xs, const double* xsq, int n, int p, double sigma, int R, int Cc) {
  if (R >= n || Cc >= n) return (R == Cc) ? 1.0 : 0.0;
  double dot = 0.0;
  for (int d = 0; d < p; ++d) dot = fma(xs[R * p + d], xs[Cc * p + d], dot);
  double r2 = -2.0 * dot + xsq[R] + xsq[Cc];
  r2 = fmax(r2, 0.0);
  const double d = sqrt(r2);
  return sigma * exp(-0.5 * (d * d));
}

// Fill the packed lower tiles of K(sigma, ls) into W (all threads).  Returns the sum over the full n x n matrix
// (for base.py:222's cov.mean()) to every thread; red = NW doubles of shared memory.
__device__ inline double rbf_fill_tiles(double* W, int NT, int n, int p, const double* xs, const double* xsq, double sigma,
                                        double* red) {
  double part = 0.0;
  for (int bi = 0; bi < NT; ++bi)
    for (int bj = 0; bj <= bi; ++bj) {
      double* tile = W + (size_t)plt(bi, bj) * TILE;
      const double wgt = (bi == bj) ? 1.0 : 2.0;
      for (int e = threadIdx.x; e < TILE; e += NTHR) {
        const int R = bi * TB + (e >> 5), Cc = bj * TB + (e & 31);
        const double v = rbf_elem(xs, xsq, n, p, sigma, R, Cc);
        tile[e] = v;
        if (R < n && Cc < n) part += wgt * v;
      }
    }
  return block_sum(part, red);
}

// dense row-major n x n (lower triangle trusted, mirrored) -> packed tiles with identity padding
__device__ inline void dense_to_tiles(const double* __restrict__ A, int n, int NT, double* W) {
  for (int bi = 0; bi < NT; ++bi)
    for (int bj = 0; bj <= bi; ++bj) {
      double* tile = W + (size_t)plt(bi, bj) * TILE;
      for (int e = threadIdx.x; e < TILE; e += NTHR) {
        const int R = bi * TB + (e >> 5), Cc = bj * TB + (e & 31);
        double v;
        if (R >= n || Cc >= n) v = (R == Cc) ? 1.0 : 0.0;
        else v = (R >= Cc) ? A[(size_t)R * n + Cc] : A[(size_t)Cc * n + R];
        tile[e] = v;
      }
    }
}

// packed tiles -> dense row-major n x n.  mode 0: lower triangle, upper zeroed (Cholesky factor);
// mode 1: symmetric (mirror the lower tiles).
__device__ inline void tiles_to_dense(const double* W, int n, int NT, double* __restrict__ A, int mode) {
  for (int R = 0; R < n; ++R)
    for (int Cc = threadIdx.x; Cc < n; Cc += NTHR) {
      double v;
      if (Cc <= R) v = W[(size_t)plt(R >> 5, Cc >> 5) * TILE + (R & 31) * TB + (Cc & 31)];
      else v = mode ? W[(size_t)plt(Cc >> 5, R >> 5) * TILE + (Cc & 31) * TB + (R & 31)] : 0.0;
      A[(size_t)R * n + Cc] = v;
    }
}

// Neal (2003) step-out / shrinkage slice sampler, sample/slice.pyx:24-62.  Executed redundantly by every thread of
// the CTA (all inputs are CTA-uniform), so LogP may contain __syncthreads.  Unif(k) returns the k-th uniform of this
// step: u0 -> level (slice.pyx:30, expon.rvs(1) has loc = 1), u1 -> placement, u2 -> step budget, u3.. -> proposals.
template <class LogP, class Unif>
__device__ inline double slice_step(double x, double w, int m, LogP&& logp, Unif&& unif, int& nev) {
  nev = 1;
  const double f0 = logp(x);
  const double z = f0 - (1.0 - log(unif(0)));
  double u = unif(1);
  double l = x - w * u;
  double r = l + w;
  const double v = unif(2);
  int j = (int)(m * v);
  int k = m - 1 - j;
  while (j > 0) {
    ++nev;
    if (!(z < logp(l))) break;
    --j;
    l -= w;
  }
  while (k > 0) {
    ++nev;
    if (!(z < logp(r))) break;
    --k;
    r += w;
  }
  int d = 3;
  u = unif(d++);
  double x1 = l + u * (r - l);
  for (;;) {
    ++nev;
    if (!(z > logp(x1))) break;
    if (x1 < x) l = x1; else r = x1;
    u = unif(d++);
    x1 = l + u * (r - l);
    if (d > 4096) break;   // unreachable in exact arithmetic; guards a degenerate (NaN-free) infinite shrink
  }
  return x1;
}

// first-touch source: RBF kernel evaluated element by element (general inputs)
struct SrcRbf {
  const double *xs, *xsq;
  int n, p;
  double sigma;
  __device__ __forceinline__ double at(int R, int Cc) const { return rbf_elem(xs, xsq, n, p, sigma, R, Cc); }
  __device__ __forceinline__ void rows(double (&a)[32], int bi, int bj, int lane) const {
#pragma unroll
    for (int c = 0; c < 32; ++c) a[c] = rbf_elem(xs, xsq, n, p, sigma, bi * TB + lane, bj * TB + c);
  }
  __device__ __forceinline__ void cfrag(CFrag& f, int bi, int bj, int g, int tig) const {
#pragma unroll
    for (int ti = 0; ti < 4; ++ti)
#pragma unroll
      for (int tj = 0; tj < 4; ++tj) {
        f.c[ti][tj][0] = rbf_elem(xs, xsq, n, p, sigma, bi * TB + 8 * ti + g, bj * TB + 8 * tj + 2 * tig);
        f.c[ti][tj][1] = rbf_elem(xs, xsq, n, p, sigma, bi * TB + 8 * ti + g, bj * TB + 8 * tj + 2 * tig + 1);
      }
  }
  __device__ __forceinline__ void fill_slot(double* slot, int bi, int bj, int lane) const {
    for (int r = 0; r < 32; ++r) slot[r * PS + lane] = rbf_elem(xs, xsq, n, p, sigma, bi * TB + r, bj * TB + lane);
  }
};

// uniform grid: tab[k] = sigma * exp(-d2[k] / (2 l^2)), k < n (one exp per lag instead of one per matrix element).
// Returns the sum of the full n x n matrix to every thread (base.py:222's cov.mean() * n^2).
__device__ inline double toeplitz_table(const double* __restrict__ d2, int n, double sigma, double ls, double* tab, double* red) {
  const double s = -0.5 / (ls * ls);
  double part = 0.0;
  for (int k = threadIdx.x; k < n; k += NTHR) {
    const double v = sigma * exp(d2[k] * s);
    tab[k] = v;
    part += (k == 0 ? (double)n : 2.0 * (double)(n - k)) * v;
  }
  return block_sum(part, red);   // contains the __syncthreads that publish tab
}

// Slice._sample (sample/slice.pyx:24-62) as a resumable state machine: the caller evaluates the log density at `cand`
// (one call site, so the factorisation code is instantiated once) and feeds the value back until `done`.
struct SliceSM {
  double x, w, z, l, r, cand;
  int m, j, k, phase, nev, d;
  bool done;
  __device__ void init(double x0, double w0, int m0) {
    x = x0; w = w0; m = m0; cand = x0; phase = 0; nev = 1; d = 3; done = false;
    z = 0.0; l = r = x0; j = k = 0;
  }
  template <class Unif>
  __device__ void feed(double lp, Unif&& unif) {
    if (phase == 0) {
      z = lp - (1.0 - log(unif(0)));               // slice.pyx:30: expon.rvs(1) has loc = 1
      const double u = unif(1);
      l = x - w * u;
      r = l + w;
      const double v = unif(2);
      j = (int)(m * v);
      k = m - 1 - j;
      phase = 1;
    } else if (phase == 1) {
      if (z < lp) { --j; l -= w; } else phase = 2;
    } else if (phase == 2) {
      if (z < lp) { --k; r += w; } else phase = 3;
    } else {
      if (!(z > lp)) { done = true; return; }
      if (cand < x) l = cand; else r = cand;
      if (d > 4096) { done = true; return; }       // unreachable in exact arithmetic
    }
    if (phase == 1) {
      if (j > 0) { cand = l; ++nev; return; }
      phase = 2;
    }
    if (phase == 2) {
      if (k > 0) { cand = r; ++nev; return; }
      phase = 3;
    }
    const double u = unif(d++);
    cand = l + u * (r - l);
    ++nev;
  }
};

// ---- shared by the dense kernels (gpf_kernels.cuh, gpf_cond.cuh) and the Schur kernels (gpf_schur*.cuh) ----
// Base.observationLikelihood (base.py:196-214) for one candidate, given the sum of squared residuals
__device__ __forceinline__ double obs_loglik_eval(double sse, double nobs, double x, double lb, double ub) {
  const double sd = sqrt(exp10(x));
  if (sd < lb || sd > ub) return -(double)INFINITY;
  return -0.5 * sse / (sd * sd) - nobs * (log(sd) + 0.5 * LOG2PI) - log(ub - lb);
}


// arguments of the prior slice kernels (k_prior, k_prior_schur): modes in gpf_kernels.cuh
struct PriorArgs {
  int g, mode;
  const double* cand;
  double* outv;
  double w_s, w_l;
  int m_s, m_l;
  unsigned sampler_id, sweep;
  const double* u_inject;
  int n_inject;
  int* nevals;       // mode 2/3: evaluations of that step; mode 4: sigma step, nevals[C + ch]: lengthscale step
  // "fold" (register-resident Schur kernels in a pipelined sweep, mode 4): the warp that finishes the LAST prior group of a
  // chain's sweep stores the chain's history row and runs the y_sigma step of the next sweep itself
  int* fold_done;    // per chain: group units finished since the start of the gpf_sweep call (null: no fold)
  int fold_target;   // P * (sweeps finished including this one)
  double* fold_row;  // history row of this sweep (chain 0 of the launch), or null
  size_t fold_cols;  // its row stride per chain
  int fold_next_y;   // run the y_sigma step of sweep + 1
  double w_y;        // slice width / step budget of y_sigma
  int m_y;
};

// element e of the 2n standard normals of one function draw (z1 = elements [0, n), z2 = [n, 2n)): Box-Muller pair e >> 1
__device__ __forceinline__ double normal_elem(const Stream& rng, unsigned gch, unsigned sweep, unsigned sid, int e) {
  double z0, z1;
  rng.normal_pair(gch, sweep, sid, (unsigned)(e >> 1), z0, z1);
  return (e & 1) ? z1 : z0;
}

}  // namespace gpf
