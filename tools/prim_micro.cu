// standalone timing of potrf_tile / trsm_slot / tile_mma (cycles per call)
#include <cstdio>
#include <vector>
#include "gpf_tile.cuh"
using namespace gpf;
__global__ void __launch_bounds__(NTHR, 2) k_prim(const double* A, double* out, long long* cyc, int reps, int nw, int what) {
  extern __shared__ __align__(16) double sm_[];
  double (*DT)[TILE] = (double (*)[TILE])sm_;
  double (*slot)[TB * PS] = (double (*)[TB * PS])(sm_ + 8 * TILE);
  double (*rinv)[32] = (double (*)[32])(sm_ + 8 * TILE + 8 * TB * PS);
  double (*piv)[32] = (double (*)[32])(sm_ + 8 * TILE + 8 * TB * PS + 256);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c = 0; c < 32; ++c) { DT[warp][c * 32 + lane] = A[lane * 32 + c]; slot[warp][lane * PS + c] = 0.01 * (lane + c); }
  rinv[warp][lane] = 1.0 + 0.01 * lane;
  __syncthreads();
  double acc = 0;
  long long t0 = clock64();
  if (warp < nw) {
    for (int r = 0; r < reps; ++r) {
      if (what == 0) {
        for (int c = 0; c < 32; ++c) DT[warp][c * 32 + lane] = A[lane * 32 + c];
        __syncwarp();
        acc += potrf_tile(DT[warp], rinv[warp], piv[warp]) ? 1.0 : 0.0;
      } else {
        acc += trsm_slot(slot[warp], DT[warp], rinv[warp], false);
      }
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = (t1 - t0) / reps;
  if (acc == 12345.0) out[0] = acc;
}
int main() {
  std::vector<double> h(TILE);
  for (int r = 0; r < 32; ++r) for (int c = 0; c < 32; ++c) h[r * 32 + c] = (r == c ? 40.0 : 0.0) + 1.0 / (1 + abs(r - c));
  double *dA, *dout; long long* dc;
  cudaMalloc(&dA, TILE * 8); cudaMalloc(&dout, 8); cudaMalloc(&dc, 296 * 8);
  cudaMemcpy(dA, h.data(), TILE * 8, cudaMemcpyHostToDevice);
  long long hc[296];
  const int SMB = (8 * TILE + 8 * TB * PS + 512) * 8;
  cudaFuncSetAttribute(k_prim, cudaFuncAttributeMaxDynamicSharedMemorySize, SMB);
  for (int ctas : {148}) for (int nw : {1, 4, 8}) for (int what : {0, 1}) {
    k_prim<<<ctas, NTHR, SMB>>>(dA, dout, dc, 20, nw, what); cudaMemcpy(hc, dc, 8 * ctas, cudaMemcpyDeviceToHost);
    printf("%s ctas %3d warps %d: %lld cycles/call\n", what ? "trsm_slot " : "potrf_tile", ctas, nw, hc[0]);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
