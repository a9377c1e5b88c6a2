// micro-benchmarks of the tile engine (cycles, per-phase trace of CTA 0)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DGPF_TRACE -I../gpfanova_b200/csrc -o engine_micro engine_micro.cu
#include <cstdio>
#include <vector>
#include "gpf_tile.cuh"
using namespace gpf;

extern __shared__ __align__(16) double smem[];

template <int AUG, bool KEEPL>
__global__ void __launch_bounds__(NTHR, 2) k_engine(int n, int NT, int NE, int reps, double* ws, size_t stride, const double* F,
                                                    long long* cyc, double* out, long long* bs) {
  EngineSmem sm(smem, NT, NE);
  double* tab = smem + engine_smem_doubles(NT, NE);
  double* W = ws + (size_t)blockIdx.x * stride;
  for (int k = threadIdx.x; k < n; k += NTHR) tab[k] = exp(-0.5 * (k * 2.0 / 255) * (k * 2.0 / 255) / 0.25);
  for (int t = threadIdx.x; t < NT * TB; t += NTHR) sm.dvec[t] = (t < n) ? 1e-6 : 0.0;
  __syncthreads();
  long long t0 = clock64();
  bool ok = true;
  for (int r = 0; r < reps; ++r) {
    EngineIO io{W, W + (size_t)plt_tiles(NT) * TILE, RowSrc{F, nullptr, n, 4, n}, nullptr, 0};
    ok = chol_engine<AUG, KEEPL>(SrcToeplitz{tab, n}, io, NT, NE, 1, sm) && ok;
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) { cyc[blockIdx.x] = (t1 - t0) / reps; out[blockIdx.x] = sm.scal[0] + (ok ? 0 : 1e9); }
  if (KEEPL && bs) {
    double* xs = tab;   // 2 x NT*32 doubles (the lag table is no longer needed)
    for (int t = threadIdx.x; t < 2 * NT * TB; t += NTHR) xs[t] = 1.0;
    __syncthreads();
    long long a0 = clock64();
    for (int r = 0; r < reps; ++r) back_solve<2>(W, NT, xs, NT * TB, sm, nullptr);
    long long a1 = clock64();
    long long a2 = clock64();
    if (threadIdx.x == 0) { bs[2 * blockIdx.x] = (a1 - a0) / reps; bs[2 * blockIdx.x + 1] = (a2 - a1) / reps; }
  }
}

template <int AUG, bool KEEPL>
void run(const char* name, int ctas, int NE, int n = 256) {
  const int NT = (n + 31) / 32, reps = 10;
  size_t stride = (size_t)(plt_tiles(NT) + NE * NT) * TILE;
  double *ws, *F, *out; long long* cyc;
  cudaMalloc(&ws, ctas * stride * 8); cudaMalloc(&F, 4 * n * 8); cudaMalloc(&out, ctas * 8); cudaMalloc(&cyc, ctas * 8);
  std::vector<double> hF(4 * n, 0.5);
  cudaMemcpy(F, hF.data(), 4 * n * 8, cudaMemcpyHostToDevice);
  size_t smemb = (engine_smem_doubles(NT, NE) + 2 * NT * TB) * 8;
  long long* bs; cudaMalloc(&bs, ctas * 16); cudaMemset(bs, 0, ctas * 16);
  cudaFuncSetAttribute(k_engine<AUG, KEEPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemb);
  long long z[64] = {0};
  cudaMemcpyToSymbol(gpf_trace, z, sizeof z);
  k_engine<AUG, KEEPL><<<ctas, NTHR, smemb>>>(n, NT, NE, reps, ws, stride, F, cyc, out, bs);
  long long hb[2]; cudaMemcpy(hb, bs, 16, cudaMemcpyDeviceToHost);
  std::vector<long long> hc(ctas); std::vector<double> ho(ctas);
  cudaMemcpy(hc.data(), cyc, ctas * 8, cudaMemcpyDeviceToHost); cudaMemcpy(ho.data(), out, ctas * 8, cudaMemcpyDeviceToHost);
  long long tr[64]; cudaMemcpyFromSymbol(tr, gpf_trace, sizeof tr);
  printf("%-22s ctas %3d: %7lld cycles/factorisation (logdet %.6f) %s  back_solve<2> %lld\n", name, ctas, hc[0], ho[0], cudaGetErrorString(cudaGetLastError()), hb[0]);
  const char* nm[8] = {"prologue", "wait-top", "phase2", "wait-p2", "look-mma", "phase3", "-", "look-potrf"};
  for (int w = 0; w < 8; w += 1) {
    printf("   warp %d:", w);
    for (int s = 0; s < 8; ++s) printf(" %s %6lld", nm[s], tr[w * 8 + s] / reps);
    printf("\n");
  }
  cudaFree(ws); cudaFree(F); cudaFree(out); cudaFree(cyc);
}

int main() {
  run<AUG_RHS, true>("chol+rhs keepL (fn)", 148, 1);
  run<AUG_RHS, true>("chol+rhs keepL (fn)", 296, 1);
  run<AUG_IDENT, false>("inverse (kinv)", 296, 0);
  run<AUG_IDENT, false>("inverse n=128", 148, 0, 128);
  run<AUG_IDENT, false>("inverse n=128", 296, 0, 128);
  return 0;
}
