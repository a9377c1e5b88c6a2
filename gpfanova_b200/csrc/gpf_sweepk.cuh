// Whole sweeps in ONE launch, without a device-wide barrier between the samplers or between sweeps
// (sample/container.py:29-78 for a balanced design on a uniform grid).
//
// Chains never interact, and inside one chain a sweep is a small dependency graph:
//     y_sigma slice  ->  per prior group g:  functions of g (sfn_item)  ->  sigma + lengthscale slices of g (schur_slice_chain)
//     all groups done  ->  store the row  ->  y_sigma slice of the next sweep
// The separate kernels of gpf_sweep put five device-wide barriers into every sweep; each waits for the slowest of its items
// (a slice step takes 4..12 evaluations), and with few chains per GPU the waiting dominates (512 chains: 1.43 ms per sweep
// where the work is 0.74 ms).  Here a WARP takes "units" from one ticket counter:
//     tickets 0 .. C-1                      Y unit: y_sigma slice of the launch's first sweep for chain t
//     then, sweep-major, largest group first: G unit (s, g, chain): functions + slices of group g in sweep s
// A G unit waits (acquire-polling one word per chain) until the y_sigma step of its sweep has been done; the warp that finishes
// the LAST group unit of a chain's sweep stores the chain's history row and runs the y_sigma step of the next sweep itself.  A
// unit only ever waits for units with smaller tickets, which are running or finished, so the oldest unit in flight never
// waits: no deadlock, and the chains drift apart freely -- the time of a launch is (total work) / (throughput) plus one unit.
// Results are bit-identical to the kernel-per-sampler sweeps: same Philox counters, same arithmetic, same summation order
// in the y_sigma residual (the CTA-wide reduction of k_obs is replayed by one warp).
#pragma once
#include "gpf_schurfn.cuh"

namespace gpf {


__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int* p, int v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

__global__ void __launch_bounds__(128, GPF_SFN_MINB) k_chain_sweeps(CtxDev c, SweepKArgs a) {
  extern __shared__ __align__(16) double gpf_smem_swk[];      // per warp: SFN_NG x nld (scratch of the function step)
  const int n = c.n, nld = ((n + 7) >> 3) << 3;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const Stream rng{c.k0, c.k1};
  double* zs = gpf_smem_swk + (size_t)warp * (SFN_SMEM_ROWS * nld);
  const long long per_sweep = (long long)c.C * c.P;
  const long long total = (long long)c.C + per_sweep * a.n_sweeps;
  const size_t cols = (size_t)c.H + (size_t)c.f * n;
  for (;;) {
    const long long t = (long long)(unsigned)next_chain_warp(c.sched, lane);
    if (t >= total) break;
    int ch, s;
    bool last = false;
    if (t < c.C) {                                // ---- Y unit of the first sweep ----
      ch = (int)t;
      s = -1;
      last = true;
    } else {                                      // ---- G unit ----
      const long long u = t - c.C;
      s = (int)(u / per_sweep);
      const int rem = (int)(u % per_sweep);
      ch = rem % c.C;
      const int g = a.fn.gorder[rem / c.C];
      // wait for the y_sigma step of sweep s of this chain
      int have = 0;
      for (unsigned spin = 0;; ++spin) {
        if (lane == 0) have = ld_acquire(a.prog + ch);
        have = __shfl_sync(FULLMASK, have, 0);
        if (have > s) break;
        if (spin > (1u << 24)) {                  // ~10 s: something is broken; give up instead of hanging the device
          if (lane == 0) atomicExch(a.err, 1);
          break;
        }
        __nanosleep(200);
      }
      if (have <= s) break;
      __threadfence();
      const unsigned sweep = a.sweep0 + (unsigned)s;
      SchurFnArgs fa = a.fn;
      fa.sweep = sweep;
      if (n > 128) sfn_item<8>(c, fa, rng, ch, g, lane, zs);
      else if (n > 64) sfn_item<4>(c, fa, rng, ch, g, lane, zs);
      else if (n > 32) sfn_item<2>(c, fa, rng, ch, g, lane, zs);
      else sfn_item<1>(c, fa, rng, ch, g, lane, zs);
      __syncwarp();
      const PriorArgs pa{g, 4, nullptr, nullptr, a.slice_w[1 + 2 * g], a.slice_w[2 + 2 * g], a.slice_m[1 + 2 * g], a.slice_m[2 + 2 * g],
                         a.sid_sigma[g], sweep, nullptr, 0, nullptr};
      const int ng = c.group_off[g + 1] - c.group_off[g];
      if (ng == 1) schur_slice_chain<1>(c, pa, ch, lane, nullptr, nullptr, nullptr, nullptr);
      else if (ng == 2) schur_slice_chain<2>(c, pa, ch, lane, nullptr, nullptr, nullptr, nullptr);
      else schur_slice_chain<4>(c, pa, ch, lane, nullptr, nullptr, nullptr, nullptr);
      // publish: this group's functions and hyper-parameters are in memory before the count moves
      __threadfence();                            // every lane has written function values
      __syncwarp();
      int old = 0;
      if (lane == 0) old = atomicAdd(a.done + ch, 1);
      old = __shfl_sync(FULLMASK, old, 0);
      last = (old + 1 == c.P * (s + 1));
      if (last) __threadfence();                  // the other groups' units counted before this one: their data is in L2
    }
    if (last) {
      // ---- the chain's sweep s is complete: store its row, then the y_sigma step of sweep s+1 ----
      if (s >= 0 && a.history != nullptr && (a.thin <= 0 || (s + 1) % a.thin == 0)) {
        const long long row = (a.thin <= 0) ? s : (s + 1) / a.thin - 1;
        if (row < a.rows_cap) {
          double* dst = a.history + ((size_t)row * c.C + ch) * cols;
          const double* hy = c.hyper + (size_t)ch * c.H;
          const double* Fc = c.F + (size_t)ch * c.f * n;
          for (int e = lane; e < c.H; e += 32) dst[e] = __ldcg(hy + e);
          for (int e = lane; e < c.f * n; e += 32) __stcs(dst + c.H + e, __ldcg(Fc + e));
        }
      }
      if (s + 1 < a.n_sweeps) {
        schur_ysigma_warp(c, rng, ch, a.sweep0 + (unsigned)(s + 1), a.slice_w[0], a.slice_m[0], a.sid_y, lane);
        __syncwarp();
        if (lane == 0) {
          __threadfence();
          st_release(a.prog + ch, s + 2);
        }
      }
    }
    __syncwarp();
  }
}

}  // namespace gpf
