// Micro-benchmark of the register-resident Schur evaluations of gpf_schur.cuh (one warp per chain):
//   variant 0 = schur_eval_reg (normalised generator, 8 slots per lane throughout)
//   variant 1 = schur_eval_staged (scaled generator, 8 -> 4 -> 2 -> 1 slots per lane)
// timed alone on the SM (latency: one warp per SM) and at the kernel's occupancy (444 CTAs x 4 warps), values compared.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-extended-lambda -lineinfo -I../gpfanova_b200/csrc -I.. [-DMINB=2|3|4] -o schur_micro schur_micro.cu
#ifndef MINB
#define MINB 3
#endif
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../include/gpfanova_b200.h"
#include "gpf_kernels.cuh"
#include "gpf_schur.cuh"
using namespace gpf;

template <int VAR, int NG>
__global__ void __launch_bounds__(128, MINB) k_micro(CtxDev c, int ng, int reps, int nchain, double* out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wid = __reduce_max_sync(FULLMASK, blockIdx.x * (blockDim.x >> 5) + warp), nw = gridDim.x * (blockDim.x >> 5);
  const int* fns = c.group_fns;
  for (int ch = wid; ch < nchain; ch += nw) {
    const double* Fc = c.F + (size_t)ch * c.f * c.n;
    double ld = 0.0, qd = 0.0;
    int bad = 0;
    for (int rep = 0; rep < reps; ++rep) {
      const double ls = exp10(c.hyper[ch] + 1e-3 * rep);
      SchurOut o;
      if (VAR == 0) o = schur_eval_reg<NG>(c, c.n, ng, fns, Fc, 1.3, ls, lane);
      else o = schur_eval_staged<NG>(c, c.n, ng, fns, Fc, 1.3, ls, lane);
      if (rep == 0) { ld = o.logdet; qd = o.quad; }
      bad += o.ok ? 0 : 1;
    }
    if (lane == 0) { out[3 * ch] = ld; out[3 * ch + 1] = qd; out[3 * ch + 2] = (double)bad; }
  }
}

template <int VAR, int NG>
float run(CtxDev c, int ng, int reps, int nchain, int grid, int block, double* out) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  k_micro<VAR, NG><<<grid, block>>>(c, ng, 1, nchain, out);
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  k_micro<VAR, NG><<<grid, block>>>(c, ng, reps, nchain, out);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); exit(1); }
  return ms;
}

int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 256;
  const int C = 148 * 48, f = 4;      // 4 rounds at 12 warps per SM, 3 at 16
  std::vector<double> d2(n), F((size_t)C * f * n), hy(C);
  for (int k = 0; k < n; ++k) { const double d = 2.0 * k / (n - 1); d2[k] = d * d; }
  srand(1);
  for (int ch = 0; ch < C; ++ch) {
    hy[ch] = -0.8 + 1.6 * (rand() / (double)RAND_MAX);
    for (int j = 0; j < f; ++j)
      for (int t = 0; t < n; ++t)
        F[((size_t)ch * f + j) * n + t] = sin(3.0 * (j + 1) * (-1.0 + 2.0 * t / (n - 1)) + ch) + 1e-3 * (rand() / (double)RAND_MAX);
  }
  std::vector<int> gf = {0, 1, 2, 3};
  double *dd2, *dF, *dh, *dout0, *dout1;
  int* dgf;
  cudaMalloc(&dd2, n * 8); cudaMalloc(&dF, F.size() * 8); cudaMalloc(&dh, C * 8); cudaMalloc(&dgf, 16);
  cudaMalloc(&dout0, C * 24); cudaMalloc(&dout1, C * 24);
  cudaMemcpy(dd2, d2.data(), n * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(dF, F.data(), F.size() * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(dh, hy.data(), C * 8, cudaMemcpyHostToDevice);
  cudaMemcpy(dgf, gf.data(), 16, cudaMemcpyHostToDevice);
  CtxDev c{};
  c.n = n; c.f = f; c.C = C; c.d2 = dd2; c.F = dF; c.hyper = dh; c.group_fns = dgf; c.prior_jitter = 1e-6;
  int dev_clock; cudaDeviceGetAttribute(&dev_clock, cudaDevAttrClockRate, 0);
  const double ghz = dev_clock * 1e-6;
  const int reps = 8;
  auto report = [&](const char* name, int ng, float ms_lat, float ms_thr) {
    // latency: 148 warps each doing reps evaluations; throughput: 4096 chains x reps over the whole GPU
    printf("%-28s ng=%d  alone: %8.0f cycles/eval   full: %7.3f ms for %d evals = %6.1f cycles/eval/SM-warp-slot (%.2f us/eval/GPU)\n",
           name, ng, ms_lat * 1e-3 * ghz * 1e9 / reps, ms_thr, C * reps, ms_thr * 1e-3 * ghz * 1e9 / (C * reps / (148.0 * 4 * MINB)),
           ms_thr * 1e3 / (C * reps));
  };
#define BOTH(NG, ngv)                                                                                         \
  {                                                                                                           \
    float l0 = run<0, NG>(c, ngv, reps, 148, 148, 32, dout0), t0 = run<0, NG>(c, ngv, reps, C, 148 * MINB, 128, dout0); \
    float l1 = run<1, NG>(c, ngv, reps, 148, 148, 32, dout1), t1 = run<1, NG>(c, ngv, reps, C, 148 * MINB, 128, dout1); \
    report("normalised, 8 slots", ngv, l0, t0);                                                               \
    report("scaled, staged", ngv, l1, t1);                                                                    \
    std::vector<double> o0(3 * C), o1(3 * C);                                                                 \
    cudaMemcpy(o0.data(), dout0, C * 24, cudaMemcpyDeviceToHost);                                             \
    cudaMemcpy(o1.data(), dout1, C * 24, cudaMemcpyDeviceToHost);                                             \
    double eld = 0, eq = 0; int bad = 0;                                                                      \
    for (int ch = 0; ch < C; ++ch) {                                                                          \
      eld = fmax(eld, fabs(o0[3 * ch] - o1[3 * ch]) / fmax(1.0, fabs(o0[3 * ch])));                           \
      eq = fmax(eq, fabs(o0[3 * ch + 1] - o1[3 * ch + 1]) / fmax(1e-300, fabs(o0[3 * ch + 1])));              \
      bad += (int)(o0[3 * ch + 2] + o1[3 * ch + 2]);                                                          \
    }                                                                                                         \
    printf("   max rel diff logdet %.3e  quad %.3e  breakdowns %d   (sample: %.10g %.10g | %.10g %.10g)\n", eld, eq, bad, o0[0], o0[1], o1[0], o1[1]); \
  }
  printf("n = %d, clock %.3f GHz\n", n, ghz);
  BOTH(1, 1)
  BOTH(2, 2)
  BOTH(4, 4)
  return 0;
}
