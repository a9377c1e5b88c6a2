// Argument blocks, launch constants and host-side launchers of the Schur kernels.  The kernels themselves are compiled in
// their own translation units (gpf_schur_tu.cu, gpf_sweepk_tu.cu: they are most of the library's compile time); the C ABI
// (gpf_api.cu) only sees this header.
#pragma once
#include "gpf_common.cuh"

namespace gpf {

// shared memory per warp, in doubles: us, v, e (rotation 1 - rho^2 per step), tab, r[ng][n]
__host__ __device__ inline size_t schur_warp_doubles(int n, int ng) { return (size_t)(4 + ng) * n; }

// same modes as k_prior (PriorArgs): 0/1 evaluate sigma / lengthscale candidate, 2/3 slice step, 4 fused sigma + ls
// resident CTAs per SM of the register-resident variants: groups of one or two functions need fewer registers
#ifndef GPF_SCHUR_MINB_SMALL
#define GPF_SCHUR_MINB_SMALL 3
#endif
__host__ __device__ constexpr int schur_minb(int NG) { return NG == 0 ? 2 : (NG <= 2 ? GPF_SCHUR_MINB_SMALL : 3); }

constexpr int SFN_NG = 4;          // functions per prior group held in registers
constexpr int SFN_NC = 2;          // right-hand sides per pass of the embedded recursion
constexpr int SFN_SMEM_ROWS = SFN_NG;       // shared-memory rows of nld doubles per warp: z1 / f / q of every function
// (keeping z2 in four more rows instead of regenerating it for step 4 was SLOWER: 64 KB per CTA leave too little L1)

struct SchurFnArgs {
  const int* gorder;      // P groups, largest first
  const int* gcls_off;    // P+1
  const int* cls_off;     // CSR offsets of the classes into cls_fn / cls_sid
  const int* cls_fn;
  const unsigned* cls_sid;
  unsigned sweep;
  int ngroups;            // groups of this launch: gorder[0 .. ngroups)  (0: all P groups)
};

#ifndef GPF_SFN_MINB
#define GPF_SFN_MINB 3
#endif

struct SweepKArgs {
  SchurFnArgs fn;            // class lists of the function step (fn.sweep unused)
  unsigned sweep0;           // Philox sweep counter of the launch's first sweep
  int n_sweeps, thin;        // thin <= 0: every sweep is stored
  double* history;           // rows x C x (H + f n), or null
  long long rows_cap;
  const unsigned* sid_sigma; // P: sampler id of prior{g}_sigma (lengthscale: + 1)
  const double* slice_w;     // 1 + 2P slice widths / step budgets in hyper order
  const int* slice_m;
  unsigned sid_y;
  int* prog;                 // C: number of sweeps of this launch whose y_sigma step is done (zeroed before the launch)
  int* done;                 // C: group units of this launch finished (zeroed before the launch)
  int* err;                  // set if a wait ran out of patience (never in a correct run; keeps a bug from hanging the GPU)
};

// ---- launchers (defined next to the kernels) ----
// NG = 1, 2, 4: register-resident variants (block 128, no shared memory); 0: shared-memory variant
cudaError_t schur_launch_prior(int NG, int grid, int block, size_t smem, cudaStream_t st, const CtxDev& dev, const PriorArgs& a);
// items = chains x groups of the launch; block / grid / shared memory are chosen next to the kernel (GPF_SFN_WPB warps per CTA)
cudaError_t schur_launch_functions(int items, int sms, cudaStream_t st, const CtxDev& dev, const SchurFnArgs& a);
cudaError_t schur_launch_chain_sweeps(int grid, size_t smem, cudaStream_t st, const CtxDev& dev, const SweepKArgs& a);
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) of the kernels above
cudaError_t schur_set_attributes();
cudaError_t sweepk_set_attributes();

}  // namespace gpf
