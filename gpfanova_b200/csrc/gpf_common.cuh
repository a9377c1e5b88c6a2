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
enum { CNT_CHOL = 0, CNT_LOGP = 1, CNT_JITTER = 2, CNT_DRAW = 3, CNT_INV = 4, CNT_SLICE = 5, CNT_N = 8 };

// device view of one sampler context (passed by value to every kernel)
struct CtxDev {
  int n, p, m, f, P, C, H, NT, NEmax, chain_offset, maxtries;
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
  double* Kinv;               // P x C x plt_tiles(NT) tiles
  double* ws;                 // per-CTA workspace, ws_stride doubles each
  size_t ws_stride;
  double offset, prior_jitter, prior_lb, prior_ub;
  unsigned k0, k1;            // Philox key
};

__device__ __forceinline__ void status_merge(int* st, int code) {
  const int old = *st;
  if (code < 0) { if (old >= 0 || code < old) *st = code; }
  else if (old >= 0 && code > old) *st = code;
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

}  // namespace gpf
