// standalone timing of the engine primitives (cycles per call, 1/4/8 active warps per CTA, 1 CTA per SM)
#include <cstdio>
#include <vector>
#include "gpf_tile.cuh"
using namespace gpf;
__global__ void __launch_bounds__(NTHR, 2) k_prim(const double* A, double* G, double* out, long long* cyc, int reps, int nw, int what) {
  extern __shared__ __align__(16) double sm_[];
  double (*DT)[TILE] = (double (*)[TILE])sm_;
  double (*slot)[TB * PS] = (double (*)[TB * PS])(sm_ + 8 * TILE);
  double (*rinv)[32] = (double (*)[32])(sm_ + 8 * TILE + 8 * TB * PS);
  double (*piv)[32] = (double (*)[32])(sm_ + 8 * TILE + 8 * TB * PS + 256);
  double (*Linv)[256] = (double (*)[256])(sm_ + 8 * TILE + 8 * TB * PS + 512);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tig = lane & 3;
  for (int c = 0; c < 32; ++c) { DT[warp][c * 32 + lane] = A[lane * 32 + c]; slot[warp][lane * PS + c] = 0.01 * (lane + c); }
  rinv[warp][lane] = 1.0 + 0.01 * lane;
  for (int i = lane; i < 256; i += 32) Linv[warp][i] = 0.01 * i;
  __syncthreads();
  double* Gw = G + ((size_t)blockIdx.x * 8 + warp) * 4 * TILE;
  double acc = 0;
  long long t0 = clock64();
  if (warp < nw) {
    for (int r = 0; r < reps; ++r) {
      if (what == 0) {
        for (int c = 0; c < 32; ++c) DT[warp][c * 32 + lane] = A[lane * 32 + c];
        __syncwarp();
        acc += potrf_tile(DT[warp], rinv[warp], piv[warp]) ? 1.0 : 0.0;
      } else if (what == 1) {
        acc += trsm_slot(slot[warp], DT[warp], rinv[warp], false);
      } else if (what == 2) {
        trsm_hybrid(slot[warp], DT[warp], rinv[warp], lane);
      } else if (what == 3) {
        trsm_dmma(slot[warp], DT[warp], Linv[warp], g, tig);
      } else if (what == 4) {   // phase-2 tile pipeline: global rows -> slot -> trsm_hybrid -> global rows
        double a[32];
        rows_load(a, Gw + (size_t)(r & 3) * TILE, lane);
        double* ps = slot[warp] + lane * PS;
        for (int c = 0; c < 32; c += 2) *reinterpret_cast<double2*>(ps + c) = make_double2(a[c], a[c + 1]);
        __syncwarp();
        trsm_hybrid(slot[warp], DT[warp], rinv[warp], lane);
        double* dst = Gw + (size_t)((r + 1) & 3) * TILE + lane * TB;
        for (int c = 0; c < 32; c += 2) *reinterpret_cast<double2*>(dst + c) = *reinterpret_cast<const double2*>(ps + c);
      } else if (what == 5) {   // phase-3 tile: global C -> mma -> global C
        CFrag f;
        cfrag_load(f, Gw + (size_t)(r & 3) * TILE, g, tig, 4);
        tile_mma<true>(f, slot[warp], slot[(warp + 1) & 7], g, tig, 4);
        cfrag_store(f, Gw + (size_t)(r & 3) * TILE, g, tig, 4);
      } else if (what >= 8 && what <= 10) {   // potrf on warp 0 under background work on the other warps
        if (warp == 0) {
          for (int c = 0; c < 32; ++c) DT[warp][c * 32 + lane] = A[lane * 32 + c];
          __syncwarp();
          acc += potrf_tile(DT[warp], rinv[warp], piv[warp]) ? 1.0 : 0.0;
        } else if (what == 8) {               // tile_mma (LDS fragments + DMMA)
          CFrag f; cfrag_zero(f);
          for (int q = 0; q < 3; ++q) tile_mma<true>(f, slot[warp], slot[(warp + 1) & 7], g, tig, 4);
          acc += f.c[0][0][0];
        } else if (what == 9) {               // DMMA only (operands in registers)
          double c0 = 0, c1 = 0;
          for (int q = 0; q < 3 * 128; ++q) dmma884(c0, c1, 1.0 + q, 0.5);
          acc += c0 + c1;
        } else {                              // LDS only
          double s2 = 0;
          for (int q = 0; q < 3 * 64; ++q) s2 += slot[warp][(g * PS + tig + 4 * (q & 7) + 8 * PS * ((q >> 3) & 3))];
          acc += s2;
        }
      } else if (what == 11) {  // coalesced tile load: global (L2) -> slot
        tile_to_slot(Gw + (size_t)(r & 3) * TILE, slot[warp], lane);
        __syncwarp();
        acc += slot[warp][lane * PS + (r & 31)];
      } else if (what == 12) {  // coalesced tile store: slot -> global
        slot_to_tile(slot[warp], Gw + (size_t)(r & 3) * TILE, lane);
      } else if (what == 13) {  // accumulator fragment load + store
        CFrag f;
        cfrag_load(f, Gw + (size_t)(r & 3) * TILE, g, tig, 4);
        f.c[0][0][0] += 1.0;
        cfrag_store(f, Gw + (size_t)((r + 1) & 3) * TILE, g, tig, 4);
      } else if (what == 14) {  // tile load from a tile just written by this warp
        slot_to_tile(slot[warp], Gw + (size_t)(r & 3) * TILE, lane);
        __syncwarp();
        tile_to_slot(Gw + (size_t)(r & 3) * TILE, slot[warp], lane);
        __syncwarp();
        acc += slot[warp][lane * PS + (r & 31)];
      } else if (what >= 15 && what <= 18) {  // global stores followed by a CTA barrier and a shared-memory load
        CFrag f;
        cfrag_zero(f);
        f.c[0][0][0] = acc;
        if (what != 16) cfrag_store(f, Gw + (size_t)(r & 3) * TILE, g, tig, 4);
        if (what == 18) __threadfence_block();
        if (what != 17) asm volatile("bar.sync 1, %0;" ::"r"(nw * 32) : "memory");
        acc += slot[warp][lane * PS + (r & 31)];
      } else if (what == 6) {   // inv8
        inv8_blocks(DT[warp], rinv[warp], Linv[warp], lane);
        __syncwarp();
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
  double *dA, *dout, *dG; long long* dc;
  cudaMalloc(&dA, TILE * 8); cudaMalloc(&dout, 8); cudaMalloc(&dc, 296 * 8); cudaMalloc(&dG, (size_t)296 * 8 * 4 * TILE * 8);
  cudaMemset(dG, 0, (size_t)296 * 8 * 4 * TILE * 8);
  cudaMemcpy(dA, h.data(), TILE * 8, cudaMemcpyHostToDevice);
  long long hc[296];
  const int SMB = (8 * TILE + 8 * TB * PS + 512 + 8 * 256) * 8;
  cudaFuncSetAttribute(k_prim, cudaFuncAttributeMaxDynamicSharedMemorySize, SMB);
  const char* nm[19] = {"potrf_tile", "trsm_slot", "trsm_hybrid", "trsm_dmma", "phase2 tile", "phase3 tile", "inv8_blocks", "-", "potrf|bg mma", "potrf|bg dmma", "potrf|bg lds", "tile->slot", "slot->tile", "cfrag ld+st", "st then ld", "st;bar;lds", "bar;lds", "st;lds", "st;fence;bar;lds"};
  for (int what = 0; what < 19; ++what) {
    if (what == 7) continue;
    printf("%-12s", nm[what]);
    for (int nw : {1, 4, NW}) {
      k_prim<<<148, NTHR, SMB>>>(dA, dG, dout, dc, 20, nw, what); cudaMemcpy(hc, dc, 8 * 148, cudaMemcpyDeviceToHost);
      printf("  %dw: %6lld", nw, hc[0]);
    }
    printf("\n");
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
