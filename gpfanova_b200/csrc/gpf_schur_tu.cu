// Translation unit of the Schur kernels: k_prior_schur<0,1,2,4> (gpf_schur.cuh) and k_group_schurfn (gpf_schurfn.cuh)
#include "../../include/gpfanova_b200.h"
#include "gpf_schur.cuh"
#include "gpf_schurfn.cuh"
#include <algorithm>

namespace gpf {

cudaError_t schur_launch_prior(int NG, int grid, int block, size_t smem, cudaStream_t st, const CtxDev& dev, const PriorArgs& a) {
  if (NG == 1) k_prior_schur<1><<<grid, block, smem, st>>>(dev, a);
  else if (NG == 2) k_prior_schur<2><<<grid, block, smem, st>>>(dev, a);
  else if (NG == 4) k_prior_schur<4><<<grid, block, smem, st>>>(dev, a);
  else k_prior_schur<0><<<grid, block, smem, st>>>(dev, a);
  return cudaGetLastError();
}

cudaError_t schur_launch_functions(int items, int sms, cudaStream_t st, const CtxDev& dev, const SchurFnArgs& a) {
  const int nld = ((dev.n + 7) / 8) * 8;
  const size_t smem = (size_t)GPF_SFN_WPB * SFN_SMEM_ROWS * nld * sizeof(double);
  const int grid = std::min((items + GPF_SFN_WPB - 1) / GPF_SFN_WPB, (12 / GPF_SFN_WPB) * sms);
  k_group_schurfn<<<grid, 32 * GPF_SFN_WPB, smem, st>>>(dev, a);
  return cudaGetLastError();
}

cudaError_t schur_set_attributes() {
  cudaError_t e = cudaFuncSetAttribute(k_prior_schur<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(k_group_schurfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)GPF_SFN_WPB * SFN_SMEM_ROWS * 256 * sizeof(double)));
  return e;
}

}  // namespace gpf
