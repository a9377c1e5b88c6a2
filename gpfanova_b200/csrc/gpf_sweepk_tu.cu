// Translation unit of the one-launch sweep kernel k_chain_sweeps (gpf_sweepk.cuh)
#include "../../include/gpfanova_b200.h"
#define GPF_NO_SCHURFN_KERNEL
#include "gpf_sweepk.cuh"

namespace gpf {

cudaError_t schur_launch_chain_sweeps(int grid, size_t smem, cudaStream_t st, const CtxDev& dev, const SweepKArgs& a) {
  k_chain_sweeps<<<grid, 128, smem, st>>>(dev, a);
  return cudaGetLastError();
}

cudaError_t sweepk_set_attributes() {
  return cudaFuncSetAttribute(k_chain_sweeps, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)4 * SFN_SMEM_ROWS * 256 * sizeof(double)));
}

}  // namespace gpf
