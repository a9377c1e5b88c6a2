// Tile engine: blocked right-looking FP64 Cholesky of one n x n matrix per CTA (n <= 512),
// with optional "augmented rows" that ride along the factorisation:
//   AUG_RHS   : rows R appended below the matrix come out as R L^-T, i.e. row v = (L^-1 b_v)^T
//               -> forward solves / quadratic forms for free (slice log-likelihoods, base.py:233-238;
//               b of function.py:32)
//   AUG_IDENT : identity rows come out as U = L^-T; the engine accumulates U U^T = (L L^T)^-1
//               tile by tile -> Kernel.K_inv (kernel/kernel.pyx:51-64).  Fully in place: at step k
//               the live tiles are trailing A(i,j) (i>=j>k) at plt(i,j), pending U(e,j) (e<=k<j) at
//               plt(j,e) (the slot of the dead panel tile L(j,e)) and partial inverse tiles (a,b)
//               (b<=a<=k) at plt(a,b) -- always exactly NT(NT+1)/2 tiles.
// Matrix tiles are 32x32 FP64, row-major, packed lower-triangular by tile (plt), kept in a per-CTA
// global workspace that stays L2-resident (one slot per persistent CTA); the current panel lives in
// shared memory.  Trailing updates are mma.sync m8n8k4 f64 (DMMA.8x8x4 on sm_100a: measured
// 37.1 TFLOP/s on B200, profiles/fp64_peak_r01.json); the 32x32 diagonal factorisation and the panel
// solves are warp-level (registers + broadcast shared-memory reads of the column-major diagonal factor).
// Replaces linalg.jitchol's dpotrf (linalg.py:35-37).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#ifdef GPF_TRACE
// per-warp phase timers of CTA 0 (tools/engine_micro.cu): slot s accumulates the cycles since the previous stamp.
// Accumulators live in registers of lane 0 and are flushed once at the end (GPF_STAMP_FLUSH).
__device__ long long gpf_trace[8 * 8];
#define GPF_STAMP(s)                                          \
  do {                                                        \
    const long long now_ = clock64();                         \
    gpf_tacc[(s)] += now_ - gpf_tprev;                        \
    gpf_tprev = now_;                                         \
  } while (0)
#define GPF_STAMP_INIT long long gpf_tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long gpf_tprev = clock64();
#define GPF_STAMP_FLUSH                                                                          \
  if (blockIdx.x == 0 && (threadIdx.x & 31) == 0)                                                \
    for (int s_ = 0; s_ < 8; ++s_) gpf_trace[(threadIdx.x >> 5) * 8 + s_] += gpf_tacc[s_];
#else
#define GPF_STAMP(s)
#define GPF_STAMP_INIT
#define GPF_STAMP_FLUSH
#endif

// warps per CTA (2 CTAs per SM, set by shared memory): measured with the final engine 8 -> 37.2, 6 -> 36.4, 4 -> 34.4 ms
// per cfg5 sweep.  The engine is bound by the serial pivot / panel chain of each matrix and by warp-level latency
// chains, not by warp-level throughput: fewer warps contend less for the issue slots of the look-ahead warp, and at
// 128 threads per CTA nothing caps the registers per thread (the K_inv halves run as two 2-warp sub-engines).
#ifndef GPF_NW
#define GPF_NW 4
#endif
#ifndef GPF_ENGINE_INLINE
#define GPF_ENGINE_INLINE __forceinline__
#endif

namespace gpf {

constexpr int TB = 32;            // tile edge
constexpr int TILE = TB * TB;     // doubles per tile
constexpr int PS = 36;            // row stride of a panel slot in shared memory (conflict-free DMMA fragment loads)
constexpr int NW = GPF_NW;        // warps per CTA
constexpr int NTHR = NW * 32;
constexpr unsigned FULLMASK = 0xffffffffu;
constexpr int MAX_NT = 16;

enum { AUG_NONE = 0, AUG_RHS = 1, AUG_IDENT = 2 };

__host__ __device__ __forceinline__ int plt(int bi, int bj) { return (bi * (bi + 1)) / 2 + bj; }
__host__ __device__ __forceinline__ int plt_tiles(int NT) { return NT * (NT + 1) / 2; }

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// rows appended below the matrix (AUG_RHS): element (r, c) = base[idx(r) * ld + c] for r < nrows, c < ncols, else 0
struct RowSrc {
  const double* base;
  const int* rowidx;   // may be nullptr (identity)
  int ld, nrows, ncols;
  __device__ __forceinline__ double at(int r, int c) const {
    if (r >= nrows || c >= ncols) return 0.0;
    const int rr = rowidx ? rowidx[r] : r;
    return base[(size_t)rr * ld + c];
  }
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// ---- C fragment (32x32 tile in DMMA accumulator layout): c[ti][tj][e] = C[8ti+g][8tj+2tig+e] ----
struct CFrag { double c[4][4][2]; };

__device__ __forceinline__ void cfrag_zero(CFrag& f) {
#pragma unroll
  for (int ti = 0; ti < 4; ++ti)
#pragma unroll
    for (int tj = 0; tj < 4; ++tj) { f.c[ti][tj][0] = 0.0; f.c[ti][tj][1] = 0.0; }
}
__device__ __forceinline__ void cfrag_load(CFrag& f, const double* tile, int g, int tig, int nti) {
#pragma unroll
  for (int ti = 0; ti < 4; ++ti)
    if (ti < nti) {
#pragma unroll
      for (int tj = 0; tj < 4; ++tj) {
        double2 v = *reinterpret_cast<const double2*>(tile + (8 * ti + g) * TB + 8 * tj + 2 * tig);
        f.c[ti][tj][0] = v.x; f.c[ti][tj][1] = v.y;
      }
    }
}
__device__ __forceinline__ void cfrag_store(const CFrag& f, double* tile, int g, int tig, int nti) {
#pragma unroll
  for (int ti = 0; ti < 4; ++ti)
    if (ti < nti) {
#pragma unroll
      for (int tj = 0; tj < 4; ++tj)
        *reinterpret_cast<double2*>(tile + (8 * ti + g) * TB + 8 * tj + 2 * tig) = make_double2(f.c[ti][tj][0], f.c[ti][tj][1]);
    }
}
// add dv[8ti+g] on the diagonal of a diagonal tile
__device__ __forceinline__ void cfrag_add_diag(CFrag& f, const double* dv, int g, int tig) {
#pragma unroll
  for (int ti = 0; ti < 4; ++ti) {
    // element (8ti+g, 8ti+2tig+e) is diagonal iff g == 2tig+e
    const double d = dv[8 * ti + g];
    if (g == 2 * tig) f.c[ti][ti][0] += d;
    if (g == 2 * tig + 1) f.c[ti][ti][1] += d;
  }
}
__device__ __forceinline__ void cfrag_from_rows(CFrag& f, const RowSrc& s, int r0, int c0, int g, int tig, int nti) {
#pragma unroll
  for (int ti = 0; ti < 4; ++ti)
    if (ti < nti) {
#pragma unroll
      for (int tj = 0; tj < 4; ++tj) {
        f.c[ti][tj][0] = s.at(r0 + 8 * ti + g, c0 + 8 * tj + 2 * tig);
        f.c[ti][tj][1] = s.at(r0 + 8 * ti + g, c0 + 8 * tj + 2 * tig + 1);
      }
    }
}


// C (+/-)= Pa * Pb^T, Pa/Pb panel slots in shared memory (32 x PS), 128 DMMA.8x8x4 for a full tile
template <bool NEG>
__device__ __forceinline__ void tile_mma_real(CFrag& f, const double* __restrict__ Pa, const double* __restrict__ Pb,
                                         int g, int tig, int nti) {
  const double* pa = Pa + g * PS + tig;
  const double* pb = Pb + g * PS + tig;
#pragma unroll
  for (int kk = 0; kk < 8; ++kk) {
    double a[4], b[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      b[t] = pb[(8 * t) * PS + 4 * kk];
      a[t] = 0.0;
      if (t < nti) { double v = pa[(8 * t) * PS + 4 * kk]; a[t] = NEG ? -v : v; }
    }
#pragma unroll
    for (int ti = 0; ti < 4; ++ti)
      if (ti < nti) {
#pragma unroll
        for (int tj = 0; tj < 4; ++tj) dmma884(f.c[ti][tj][0], f.c[ti][tj][1], a[ti], b[tj]);
      }
  }
}

template <bool NEG>
__device__ __forceinline__ void tile_mma(CFrag& f, const double* __restrict__ Pa, const double* __restrict__ Pb, int g, int tig,
                                         int nti) {
#ifdef GPF_DBG_NOMMA
  f.c[0][0][0] += Pa[g] + Pb[tig];
#else
  tile_mma_real<NEG>(f, Pa, Pb, g, tig, nti);
#endif
}

// ---- lane = row helpers ----
__device__ __forceinline__ void rows_load(double (&a)[32], const double* tile, int lane) {
  const double2* p = reinterpret_cast<const double2*>(tile + lane * TB);
#pragma unroll
  for (int c = 0; c < 16; ++c) { double2 v = p[c]; a[2 * c] = v.x; a[2 * c + 1] = v.y; }
}
__device__ __forceinline__ void rows_store(const double (&a)[32], double* tile, int lane) {
  double2* p = reinterpret_cast<double2*>(tile + lane * TB);
#pragma unroll
  for (int c = 0; c < 16; ++c) p[c] = make_double2(a[2 * c], a[2 * c + 1]);
}

// 1/sqrt(x) from the MUFU seed (rel. error e ~ 2^-22) and one third-order step y (1 + e/2 + 3 e^2/8) (error ~ e^3);
// no special-case slow path (callers reject non-positive / non-finite pivots through the running minimum instead).
__device__ __forceinline__ double rsqrt_nr(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(-x * y, y, 1.0);
  const double p = fma(0.375, e, 0.5);
  return fma(e * y, p, y);
}

// ---- explicit shared-memory accesses (the warp-level routines below are separate functions: their pointer
// arguments would otherwise be generic) ----
__device__ __forceinline__ double lds1(unsigned a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ double2 lds2(unsigned a) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts1(unsigned a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
__device__ __forceinline__ void sts2(unsigned a, double x, double y) {
  asm volatile("st.shared.v2.f64 [%0], {%1,%2};" ::"r"(a), "d"(x), "d"(y) : "memory");
}

// Cholesky of a 32x32 tile by one warp, in place in shared memory.
//   in : DT[c*32 + r] = A[r][c] for c <= r (column-major; entries above the diagonal are not referenced)
//   out: DT[c*32 + r] = L[r][c] for c <= r, rinv[c] = 1/L[c][c], piv[c] = pivot before the square root.
// Lane = row, the row lives in registers.  Column c of L is published to DT and read back with broadcast LDS.128 --
// no 64-bit shuffles in the update -- and the next pivot is formed early by its own lane (a[c+1][c+1] - L[c+1][c]^2
// only needs that lane's values), so the dependent chain per column is mul -> fma -> shfl -> rsqrt.  dpotf2's failure
// rule (pivot <= 0 or NaN) is evaluated once from the running minimum.  A separate (noinline) function on purpose:
// it gets its own register allocation instead of competing with the caller's live state (measured 3x faster).
static __device__ __noinline__ bool potrf_tile(double* DT_, double* rinv_, double* piv_) {
  const int lane = threadIdx.x & 31;
  const unsigned DT = smem_u32(DT_), rinv_out = smem_u32(rinv_), piv_out = smem_u32(piv_);
  double a[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) a[c] = lds1(DT + 8u * (c * TB + lane));
  double piv = __shfl_sync(FULLMASK, a[0], 0);
  double dmin = piv, dmax = piv;
  double rinv = rsqrt_nr(piv);
#pragma unroll
  for (int c = 0; c < 32; ++c) {
    const double l = a[c] * rinv;
    sts1(DT + 8u * (c * TB + lane), l);
    if (lane == c) { sts1(rinv_out + 8u * c, rinv); sts1(piv_out + 8u * c, piv); }
    if (c < 31) {
      // next pivot: issue the shuffle now, consume it (rsqrt) only after the column update has been issued
      const double pn = fma(-l, l, a[c + 1]);
      piv = __shfl_sync(FULLMASK, pn, c + 1);
    }
    __syncwarp();
    if (c < 31) {
      if ((c + 1) & 1) a[c + 1] = fma(-l, lds1(DT + 8u * (c * TB + c + 1)), a[c + 1]);
#pragma unroll
      for (int c2 = (c + 2) & ~1; c2 < 32; c2 += 2) {
        const double2 v = lds2(DT + 8u * (c * TB + c2));
        a[c2] = fma(-l, v.x, a[c2]);
        a[c2 + 1] = fma(-l, v.y, a[c2 + 1]);
      }
      dmin = fmin(dmin, piv);
      dmax = fmax(dmax, piv);
      rinv = rsqrt_nr(piv);
    }
  }
  __syncwarp();
  // fmin/fmax drop NaNs: a NaN pivot makes every later pivot NaN, caught by the self-comparison of the last one
  return (dmin > 0.0) && (dmax < 1.0e300) && (piv == piv);
}

// X <- X L^-T for one 32-row block held in a shared-memory panel slot (row stride PS), in place; lane = row.
// DT: L column-major (DT[c*32 + r] = L[r][c]), rinv[c] = 1/L[c][c].  Rows with zero_row set are replaced by zeros.
// Returns the sum of squares of the lane's solved row.  Separate function for the same reason as potrf_tile.
static __device__ __noinline__ double trsm_slot(double* slot_, const double* DT_, const double* rinv_, bool zero_row) {
  const int lane = threadIdx.x & 31;
  const unsigned row = smem_u32(slot_) + 8u * (lane * PS), DT = smem_u32(DT_), rinv = smem_u32(rinv_);
  double a[32];
#pragma unroll
  for (int c = 0; c < 32; c += 2) {
    const double2 v = lds2(row + 8u * c);
    a[c] = zero_row ? 0.0 : v.x;
    a[c + 1] = zero_row ? 0.0 : v.y;
  }
  double ss = 0.0;
#pragma unroll
  for (int c = 0; c < 32; ++c) {
    const double x = a[c] * lds1(rinv + 8u * c);
    a[c] = x;
    ss = fma(x, x, ss);
    if (c < 31) {
      if ((c + 1) & 1) a[c + 1] = fma(-x, lds1(DT + 8u * (c * TB + c + 1)), a[c + 1]);
#pragma unroll
      for (int c2 = (c + 2) & ~1; c2 < 32; c2 += 2) {
        const double2 v = lds2(DT + 8u * (c * TB + c2));
        a[c2] = fma(-x, v.x, a[c2]);
        a[c2 + 1] = fma(-x, v.y, a[c2 + 1]);
      }
    }
  }
#pragma unroll
  for (int c = 0; c < 32; c += 2) sts2(row + 8u * c, a[c], a[c + 1]);
  __syncwarp();
  return ss;
}

// Inverses of the four 8x8 diagonal blocks of the 32x32 factor in DT (column-major), by one warp:
// lane = (block J = lane / 8, column j = lane % 8) solves L_JJ x = e_j by forward substitution.
// Linv[J*64 + i*8 + j] = (L_JJ^-1)[i][j].
__device__ __forceinline__ void inv8_blocks(const double* DT, const double* rinv, double* Linv, int lane) {
  const int J = lane >> 3, j = lane & 7;
  double x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    double s = (i == j) ? 1.0 : 0.0;
#pragma unroll
    for (int q = 0; q < i; ++q) s = fma(-DT[(8 * J + q) * TB + 8 * J + i], x[q], s);   // L[8J+i][8J+q]
    x[i] = s * rinv[8 * J + i];
    Linv[J * 64 + i * 8 + j] = x[i];
  }
}

// X <- X L^-T for one 32-row block held in a shared-memory panel slot S (row stride PS), in place, on the tensor
// pipe: blocked substitution over 8-column blocks,  Y_J = (X_J - sum_{I<J} Y_I L_JI^T) (L_JJ^-1)^T,  every product a
// DMMA.8x8x4 (80 per tile).  The warp-level scalar version (one row per lane, L broadcast from shared memory) is
// bound by shared-memory return bandwidth: 8 bytes per lane per FMA, measured 6.2k cycles per tile with 8 warps.
// DT: L column-major, Linv: inverses of its 8x8 diagonal blocks (inv8_blocks).
__device__ __forceinline__ void trsm_dmma(double* S, const double* __restrict__ DT, const double* __restrict__ Linv, int g,
                                          int tig) {
#pragma unroll
  for (int J = 0; J < 4; ++J) {
    double c[4][2];
#pragma unroll
    for (int ti = 0; ti < 4; ++ti) {
      const double2 v = *reinterpret_cast<const double2*>(S + (8 * ti + g) * PS + 8 * J + 2 * tig);
      c[ti][0] = v.x; c[ti][1] = v.y;
    }
#pragma unroll
    for (int I = 0; I < J; ++I)
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        const double b = DT[(8 * I + 4 * kk + tig) * TB + 8 * J + g];          // L[8J+g][8I+4kk+tig]
#pragma unroll
        for (int ti = 0; ti < 4; ++ti) dmma884(c[ti][0], c[ti][1], -S[(8 * ti + g) * PS + 8 * I + 4 * kk + tig], b);
      }
    // T = X_J - ... goes back to the slot so that it can be re-read in A-operand layout
#pragma unroll
    for (int ti = 0; ti < 4; ++ti)
      *reinterpret_cast<double2*>(S + (8 * ti + g) * PS + 8 * J + 2 * tig) = make_double2(c[ti][0], c[ti][1]);
    __syncwarp();
    double d[4][2];
#pragma unroll
    for (int ti = 0; ti < 4; ++ti) d[ti][0] = d[ti][1] = 0.0;
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
      const double b = Linv[J * 64 + g * 8 + 4 * kk + tig];                      // (L_JJ^-1)[g][4kk+tig]
#pragma unroll
      for (int ti = 0; ti < 4; ++ti) dmma884(d[ti][0], d[ti][1], S[(8 * ti + g) * PS + 8 * J + 4 * kk + tig], b);
    }
    __syncwarp();
#pragma unroll
    for (int ti = 0; ti < 4; ++ti)
      *reinterpret_cast<double2*>(S + (8 * ti + g) * PS + 8 * J + 2 * tig) = make_double2(d[ti][0], d[ti][1]);
    __syncwarp();
  }
}

// X <- X L^-T for one 32-row block in a panel slot, in place: blocked substitution whose off-diagonal block updates
// run on the tensor pipe (DMMA.8x8x4) while each 8x8 diagonal block is solved by plain substitution, one row per lane
// (no explicit inverses: same accuracy as trsm_slot, a third of its shared-memory operand traffic, 144 instead of 496
// dependent multiply-adds per lane).  DT: L column-major, rinv[c] = 1/L[c][c].
__device__ __forceinline__ void trsm_hybrid(double* S, const double* __restrict__ DT, const double* __restrict__ rinv, int lane,
                                            int nti = 4) {
  // nti: number of 8-row strips of the slot that exist (a compact right-hand-side slot holds fewer than 32 rows)
  const int g = lane >> 2, tig = lane & 3;
#pragma unroll
  for (int J = 0; J < 4; ++J) {
    if (J > 0) {
      double c[4][2];
#pragma unroll
      for (int ti = 0; ti < 4; ++ti)
        if (ti < nti) {
          const double2 v = *reinterpret_cast<const double2*>(S + (8 * ti + g) * PS + 8 * J + 2 * tig);
          c[ti][0] = v.x; c[ti][1] = v.y;
        }
#pragma unroll
      for (int I = 0; I < J; ++I)
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
          const double b = DT[(8 * I + 4 * kk + tig) * TB + 8 * J + g];          // L[8J+g][8I+4kk+tig]
#pragma unroll
          for (int ti = 0; ti < 4; ++ti)
            if (ti < nti) dmma884(c[ti][0], c[ti][1], -S[(8 * ti + g) * PS + 8 * I + 4 * kk + tig], b);
        }
#pragma unroll
      for (int ti = 0; ti < 4; ++ti)
        if (ti < nti) *reinterpret_cast<double2*>(S + (8 * ti + g) * PS + 8 * J + 2 * tig) = make_double2(c[ti][0], c[ti][1]);
      __syncwarp();
    }
    // diagonal block: row `lane`, 8 columns, substitution against L_JJ
    if (lane < 8 * nti) {
      double x[8];
      double* row = S + lane * PS + 8 * J;
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        const double2 v = *reinterpret_cast<const double2*>(row + j);
        x[j] = v.x; x[j + 1] = v.y;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        x[j] *= rinv[8 * J + j];
        const double* Lc = DT + (8 * J + j) * TB + 8 * J;      // column 8J+j of L, rows 8J..8J+7
#pragma unroll
        for (int j2 = j + 1; j2 < 8; ++j2) x[j2] = fma(-x[j], Lc[j2], x[j2]);
      }
#pragma unroll
      for (int j = 0; j < 8; j += 2) *reinterpret_cast<double2*>(row + j) = make_double2(x[j], x[j + 1]);
    }
    __syncwarp();
  }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULLMASK, v, o);
  return v;
}

// deterministic CTA-wide sum (fixed order); red: NW doubles of shared memory.  All threads get the result.
__device__ __forceinline__ double block_sum(double v, double* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
#pragma unroll
  for (int w = 0; w < NW; ++w) s += red[w];
  return s;
}

// shared-memory footprint (doubles) of the engine for NT matrix tile rows + NE extra panel slots; the LAST extra slot may be
// compact (rows_last = 8, 16, 24 rows instead of 32: few right-hand sides riding along, e.g. one or two functions)
__host__ __device__ inline size_t engine_slot_doubles(int NT, int NE, int rows_last) {
  return (NE > 0) ? (size_t)(NT + NE - 1) * TB * PS + (size_t)rows_last * PS : (size_t)NT * TB * PS;
}
__host__ __device__ inline size_t engine_smem_doubles(int NT, int NE, int rows_last = TB) {
  const size_t d = (size_t)TILE /*DT*/ + 256 /*Linv*/ + engine_slot_doubles(NT, NE, rows_last) + (size_t)NT * TB /*rinv*/ +
                   (size_t)NT * TB /*dvec*/ + (size_t)TB /*piv*/ + (size_t)NT * TB /*pivall*/ +
                   (size_t)(NE > 0 ? NE : 1) * NT /*quadpart*/ + 16 + (size_t)(NT + NE + 2) / 2 /*slot flags*/;
  return (d + 3) & ~(size_t)3;   // 32-byte granularity: two engines may sit back to back in one CTA
}

struct EngineSmem {
  double* DT;     // 32 x 32, current diagonal factor, column-major: DT[c*32 + r] = L_kk[r][c]
  double* Linv;   // 4 x 64, inverses of the 8x8 diagonal blocks of the current diagonal factor
  double* P;      // (NT+NE) slots of 32 x PS (the last extra slot may be compact)
  double* rinv;   // NT*32 reciprocal diagonal of L
  double* dvec;   // NT*32 additive diagonal applied at first touch (jitter, offset, n_t/sigma_y; 0 on padding)
  double* piv;    // 32 pivots of the current diagonal tile
  double* pivall; // NT*32: every pivot of the factorisation (log-determinant is summed once at the end)
  double* quad;   // NE*NT partial sums of squares of finished augmented rows (deterministic reduction)
  double* scal;   // [0] logdet, [1] fail flag (as double), [2..9] block_sum scratch, [10] tile counter
  int* ready;     // NT+NE: ready[slot] = k+1 once the panel slot holds its solved tile of step k
  __device__ EngineSmem(double* base, int NT, int NE, int rows_last = TB) {
    DT = base;
    Linv = DT + TILE;
    P = Linv + 256;
    rinv = P + engine_slot_doubles(NT, NE, rows_last);
    dvec = rinv + NT * TB;
    piv = dvec + NT * TB;
    pivall = piv + TB;
    quad = pivall + NT * TB;
    scal = quad + (size_t)(NE > 0 ? NE : 1) * NT;
    ready = reinterpret_cast<int*>(scal + 16);
  }
  __device__ double* red() const { return scal + 2; }
};

// 8 KB tile in global memory (row-major, stride 32) <-> panel slot in shared memory (row stride PS), one warp,
// fully coalesced: every instruction moves 512 contiguous bytes (a row per lane would touch 32 cache lines per
// instruction: measured 16.7k cycles per panel tile with 8 warps against 4.6k for one).
__device__ __forceinline__ void tile_to_slot(const double* __restrict__ tile, double* slot, int lane, int nti = 4) {
#pragma unroll
  for (int q = 0; q < 16; ++q) {
    if (q >= 4 * nti) break;                             // rows 2q, 2q+1 (compact right-hand-side slot: 8 nti rows)
    const int idx = q * 32 + lane;                       // 16-byte chunk index: row = idx / 16, chunk in row = idx % 16
    const double2 v = *reinterpret_cast<const double2*>(tile + 2 * idx);
    *reinterpret_cast<double2*>(slot + (idx >> 4) * PS + 2 * (idx & 15)) = v;
  }
}
__device__ __forceinline__ void slot_to_tile(const double* slot, double* __restrict__ tile, int lane) {
#pragma unroll
  for (int q = 0; q < 16; ++q) {
    const int idx = q * 32 + lane;
    *reinterpret_cast<double2*>(tile + 2 * idx) = *reinterpret_cast<const double2*>(slot + (idx >> 4) * PS + 2 * (idx & 15));
  }
}

// ---- first-touch sources of the matrix (step k = 0 reads the matrix from here, later steps from the workspace) ----
// tiles in memory, packed lower-triangular by tile
struct SrcTiles {
  const double* A0;
  // The source matrix is read exactly once: streaming loads (evict-first), so that it does not push the CTAs' workspaces
  // -- which are re-read many times per factorisation -- out of L2.
  __device__ __forceinline__ void rows(double (&a)[32], int bi, int bj, int lane) const {
    const double2* p = reinterpret_cast<const double2*>(A0 + (size_t)plt(bi, bj) * TILE + lane * TB);
#pragma unroll
    for (int c = 0; c < 16; ++c) { const double2 v = __ldcs(p + c); a[2 * c] = v.x; a[2 * c + 1] = v.y; }
  }
  __device__ __forceinline__ void cfrag(CFrag& f, int bi, int bj, int g, int tig) const {
    const double* tile = A0 + (size_t)plt(bi, bj) * TILE;
#pragma unroll
    for (int ti = 0; ti < 4; ++ti)
#pragma unroll
      for (int tj = 0; tj < 4; ++tj) {
        const double2 v = __ldcs(reinterpret_cast<const double2*>(tile + (8 * ti + g) * TB + 8 * tj + 2 * tig));
        f.c[ti][tj][0] = v.x; f.c[ti][tj][1] = v.y;
      }
  }
  __device__ __forceinline__ void fill_slot(double* slot, int bi, int bj, int lane) const {
    const double* tile = A0 + (size_t)plt(bi, bj) * TILE;
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const int idx = q * 32 + lane;
      const double2 v = __ldcs(reinterpret_cast<const double2*>(tile + 2 * idx));
      *reinterpret_cast<double2*>(slot + (idx >> 4) * PS + 2 * (idx & 15)) = v;
    }
  }
};
// stationary kernel on a uniform grid: K[i][j] = tab[|i-j|] (tab in shared memory, sigma folded in); identity padding
struct SrcToeplitz {
  const double* tab;
  int n;
  __device__ __forceinline__ double at(int R, int Cc) const {
    if (R >= n || Cc >= n) return (R == Cc) ? 1.0 : 0.0;
    const int d = R - Cc;
    return tab[d < 0 ? -d : d];
  }
  __device__ __forceinline__ void rows(double (&a)[32], int bi, int bj, int lane) const {
#pragma unroll
    for (int c = 0; c < 32; ++c) a[c] = at(bi * TB + lane, bj * TB + c);
  }
  __device__ __forceinline__ void cfrag(CFrag& f, int bi, int bj, int g, int tig) const {
#pragma unroll
    for (int ti = 0; ti < 4; ++ti)
#pragma unroll
      for (int tj = 0; tj < 4; ++tj) {
        f.c[ti][tj][0] = at(bi * TB + 8 * ti + g, bj * TB + 8 * tj + 2 * tig);
        f.c[ti][tj][1] = at(bi * TB + 8 * ti + g, bj * TB + 8 * tj + 2 * tig + 1);
      }
  }
  __device__ __forceinline__ void fill_slot(double* slot, int bi, int bj, int lane) const {
#pragma unroll 8
    for (int r = 0; r < 32; ++r) slot[r * PS + lane] = at(bi * TB + r, bj * TB + lane);
  }
};

// Symmetric Toeplitz matrices are centrosymmetric: with n = 2h and J the h x h exchange matrix,
//   K = Q^T diag(B+, B-) Q,   B+- = K11 +- K12 J,   Q = [[I, J], [I, -J]] / sqrt(2)   (orthogonal),
// so K^-1 follows from the inverses of two h x h matrices (a quarter of the flops).  This source presents
//   B+-[i][j] = tab[|i-j|] +- tab[n-1-i-j],  i, j < h;  identity padding.
struct SrcCentro {
  const double* tab;
  int n, h;
  double sign;
  __device__ __forceinline__ double at(int R, int Cc) const {
    if (R >= h || Cc >= h) return (R == Cc) ? 1.0 : 0.0;
    const int d = R - Cc;
    return fma(sign, tab[n - 1 - R - Cc], tab[d < 0 ? -d : d]);
  }
  __device__ __forceinline__ void rows(double (&a)[32], int bi, int bj, int lane) const {
#pragma unroll
    for (int c = 0; c < 32; ++c) a[c] = at(bi * TB + lane, bj * TB + c);
  }
  __device__ __forceinline__ void cfrag(CFrag& f, int bi, int bj, int g, int tig) const {
#pragma unroll
    for (int ti = 0; ti < 4; ++ti)
#pragma unroll
      for (int tj = 0; tj < 4; ++tj) {
        f.c[ti][tj][0] = at(bi * TB + 8 * ti + g, bj * TB + 8 * tj + 2 * tig);
        f.c[ti][tj][1] = at(bi * TB + 8 * ti + g, bj * TB + 8 * tj + 2 * tig + 1);
      }
  }
  __device__ __forceinline__ void fill_slot(double* slot, int bi, int bj, int lane) const {
#pragma unroll 8
    for (int r = 0; r < 32; ++r) slot[r * PS + lane] = at(bi * TB + r, bj * TB + lane);
  }
};

struct EngineIO {
  double* W;          // plt_tiles(NT) tiles of workspace; with KEEPL the final L tiles are left there.
                      // AUG_IDENT: the inverse (lower tiles, full diagonal tiles) is left in W
  double* Wx;         // AUG_RHS: NE*NT tiles of in-flight right-hand-side rows
  RowSrc rows;        // AUG_RHS: first-touch source of the extra rows
  double* vout;       // AUG_RHS: finished rows r < rows.nrows are written to vout[r*ldv + col] (may be nullptr)
  int ldv;
};

// factor the diagonal tile staged column-major in sm.DT (warp-level): potrf, log-determinant, optional store of L_kk
template <bool FAST_TRSM, bool KEEPL>
__device__ __forceinline__ void diag_factor(int k, int lane, double* W, const EngineSmem& sm) {
  __syncwarp();
  const bool ok = potrf_tile(sm.DT, sm.rinv + k * TB, sm.piv);
  if (FAST_TRSM) inv8_blocks(sm.DT, sm.rinv + k * TB, sm.Linv, lane);
  sm.pivall[k * TB + lane] = sm.piv[lane];      // the log-determinant is summed after the last step
  if (lane == 0 && !ok) sm.scal[1] = 1.0;
  (void)W;
}

// store the diagonal factor L_kk (column-major in sm.DT) as a row-major tile with a zero upper triangle (KEEPL);
// done by an otherwise idle warp during the panel phase, off the look-ahead's critical path
__device__ __forceinline__ void store_diag_tile(const double* DT, double* tile, int lane) {
  double* dst = tile + lane * TB;
#pragma unroll
  for (int c = 0; c < 32; c += 2) {
    const double v0 = (c <= lane) ? DT[c * TB + lane] : 0.0;
    const double v1 = (c + 1 <= lane) ? DT[(c + 1) * TB + lane] : 0.0;
    *reinterpret_cast<double2*>(dst + c) = make_double2(v0, v1);
  }
}

// named barriers (ids 1, 2; id 0 is __syncthreads): split arrive / wait so that the look-ahead warp and the others do not
// wait for each other's work that they do not depend on
template <int NTE = NTHR>
__device__ __forceinline__ void bar_arrive(int id) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "n"(NTE) : "memory"); }
template <int NTE = NTHR>
__device__ __forceinline__ void bar_sync(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(NTE) : "memory"); }

// panel-slot hand-off inside a CTA: the solving warp publishes, consumers spin (all warps of a CTA are resident)
__device__ __forceinline__ void slot_publish(int* flag, int v, int lane) {
  __syncwarp();
  if (lane == 0) asm volatile("st.release.cta.shared.s32 [%0], %1;" ::"r"(smem_u32(flag)), "r"(v) : "memory");
}
__device__ __forceinline__ void slot_wait(const int* flag, int v) {
  int cur;
  do {
    asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(cur) : "r"(smem_u32(flag)) : "memory");
  } while (cur < v);
  __syncwarp();
}

// prefetch one 8 KB tile into L1/L2 (64 lines of 128 B: two per lane)
__device__ __forceinline__ void prefetch_tile(const double* tile, int lane) {
  asm volatile("prefetch.global.L1 [%0];" ::"l"(tile + lane * 16));
  asm volatile("prefetch.global.L1 [%0];" ::"l"(tile + 512 + lane * 16));
}

// stage an accumulator fragment column-major into the (idle) DT buffer: DT[c*32 + r] = C[r][c]
__device__ __forceinline__ void cfrag_to_DT(const CFrag& f, double* DT, int g, int tig) {
#pragma unroll
  for (int ti = 0; ti < 4; ++ti)
#pragma unroll
    for (int tj = 0; tj < 4; ++tj) {
      DT[(8 * tj + 2 * tig) * TB + 8 * ti + g] = f.c[ti][tj][0];
      DT[(8 * tj + 2 * tig + 1) * TB + 8 * ti + g] = f.c[ti][tj][1];
    }
}
/* The engine.  All NTHR threads of the CTA call it with identical arguments -- or, with NWE < NW, the NWE warps of
 * sub-engine `sub` (threads sub*NWE*32 ...), several independent sub-engines side by side in one CTA, each with its own
 * EngineSmem and its own hardware barriers 1+3*sub .. 3+3*sub.
 *   AUG_RHS  : NE extra tile rows; the last extra tile row uses only nti_last 8-row strips.
 * Returns false if a pivot failed (matrix not PD); logdet (= 2 sum log L_ii) is in sm.scal[0] on success,
 * sm.quad[e*NT+k] hold the sums of squares of the finished augmented rows.
 */
template <int AUG, bool KEEPL, class Src, int NWE = NW>
__device__ GPF_ENGINE_INLINE bool chol_engine(const Src& src, const EngineIO& io, int NT, int NE, int nti_last,
                                            const EngineSmem& sm, int sub = 0) {
  constexpr bool FAST_TRSM = (AUG == AUG_RHS) && !KEEPL;
  constexpr int NTE = NWE * 32;
  const int tid = threadIdx.x - sub * NTE, warp = tid >> 5, lane = tid & 31, g = lane >> 2, tig = lane & 3;
  const int bar0 = (NWE == NW) ? 0 : 1 + 3 * sub;     // bar0: all threads of the (sub-)engine; bar0+1, bar0+2: split barriers
  auto eng_sync = [&]() {
    if constexpr (NWE == NW) __syncthreads();
    else bar_sync<NTE>(bar0);
  };
  double* W = io.W;
  double* Wx = io.Wx;
  int* tilectr2 = reinterpret_cast<int*>(sm.scal + 10);   // two ticket counters, alternating by step parity
  GPF_STAMP_INIT
  if (tid == 0) { sm.scal[0] = 0.0; sm.scal[1] = 0.0; tilectr2[0] = (NT > 1) ? 1 : 0; }
  if (tid < NT + NE) sm.ready[tid] = 0;
  eng_sync();
  // ---------------- prologue: factor diagonal tile 0 (warp 0) ----------------
  if (warp == 0) {
    double a[32];
    src.rows(a, 0, 0, lane);
    const double dv = sm.dvec[lane];
#pragma unroll
    for (int c = 0; c < 32; ++c) sm.DT[c * TB + lane] = a[c] + ((c == lane) ? dv : 0.0);
    diag_factor<FAST_TRSM, KEEPL>(0, lane, W, sm);
  }
  GPF_STAMP(0);
  for (int k = 0; k < NT; ++k) {
    eng_sync();                      // diagonal factor k (DT, rinv[k]) and the step-(k-1) updates are complete
    GPF_STAMP(1);
#ifndef GPF_DBG_NOFAIL
    if (sm.scal[1] != 0.0) return false;
#endif
    // ---------------- phase 2: panel solves  X <- X L_kk^-T, in place in the panel slots ----------------
    const int nreg = NT - 1 - k;
    const int nx = (AUG == AUG_RHS) ? NE : ((AUG == AUG_IDENT) ? (k + 1) : 0);
    // trailing tile 0 = next diagonal tile (look-ahead).  The counter of the NEXT step is initialised here: a warp
    // without a panel tile goes straight to this step's tickets, so this step's counter must be ready at the barrier.
    int* tilectr = tilectr2 + (k & 1);
    if (tid == 0) tilectr2[(k + 1) & 1] = (NT - 2 - k > 0) ? 1 : 0;
    if (KEEPL && warp == NWE - 1) store_diag_tile(sm.DT, W + (size_t)plt(k, k) * TILE, lane);
    for (int t = warp; t < nreg + nx; t += NWE) {
      const int iown = k + 1 + t;
      int slot;
      if (t < nreg) {
        slot = iown;
        if (k == 0) src.fill_slot(sm.P + (size_t)slot * TB * PS, iown, 0, lane);
        else tile_to_slot(W + (size_t)plt(iown, k) * TILE, sm.P + (size_t)slot * TB * PS, lane);
      } else {
        const int e = t - nreg;
        if (AUG == AUG_RHS) {
          slot = NT + e;
          double* sp = sm.P + (size_t)slot * TB * PS;
          const int ntx = (e == NE - 1) ? nti_last : 4;         // the last extra slot only holds 8 ntx rows
          if (k == 0) {
            if (lane < 8 * ntx) {
              double* ps = sp + lane * PS;
#pragma unroll
              for (int c = 0; c < 32; ++c) ps[c] = io.rows.at(e * TB + lane, c);
            }
          } else {
            tile_to_slot(Wx + (size_t)(e * NT + k) * TILE, sp, lane, ntx);
            __syncwarp();
            if (lane < 8 * ntx && e * TB + lane >= io.rows.nrows) {   // unused rows of the last block: keep them exactly zero
              double* ps = sp + lane * PS;
#pragma unroll
              for (int c = 0; c < 32; c += 2) *reinterpret_cast<double2*>(ps + c) = make_double2(0.0, 0.0);
            }
          }
        } else {
          slot = e;
          double* sp = sm.P + (size_t)slot * TB * PS;
          if (e == k) {
#pragma unroll 8
            for (int r = 0; r < 32; ++r) sp[r * PS + lane] = (r == lane) ? 1.0 : 0.0;
          } else tile_to_slot(W + (size_t)plt(k, e) * TILE, sp, lane);
        }
      }
      __syncwarp();
      double* slotp = sm.P + (size_t)slot * TB * PS;
      const int nts = (AUG == AUG_RHS && t >= nreg && t - nreg == NE - 1) ? nti_last : 4;
#ifndef GPF_DBG_NOTRSM
      // Substitution-accurate solves wherever the factor itself is a result (K_inv: cond ~ 1e11; L_A of the function
      // step, held to 1e-10): the explicit 8x8 block inverses of the all-tensor version cost digits there (measured:
      // K_inv residual 0.9 instead of 1e-2; conditional mean 2.3e-10 instead of 7e-11 from the truth).  The dense
      // slice likelihood (general inputs) only needs log-det and quadratic forms and takes the all-tensor version.
      if (FAST_TRSM) trsm_dmma(slotp, sm.DT, sm.Linv, g, tig);
      else trsm_hybrid(slotp, sm.DT, sm.rinv + k * TB, lane, nts);
#endif
      slot_publish(sm.ready + slot, k + 1, lane);     // consumers of this slot need not wait for the whole panel
      if (t < nreg) {
        if (KEEPL) slot_to_tile(slotp, W + (size_t)plt(k + 1 + t, k) * TILE, lane);

      } else if (AUG == AUG_RHS) {
        const int e = t - nreg;
        const double* ps = slotp + lane * PS;
        double ss = 0.0;
        if (lane < 8 * nts) {
#pragma unroll
          for (int c = 0; c < 32; c += 2) {
            const double2 v = *reinterpret_cast<const double2*>(ps + c);
            ss = fma(v.x, v.x, ss);
            ss = fma(v.y, v.y, ss);
          }
        }
        ss = warp_sum(ss);
        if (lane == 0) sm.quad[e * NT + k] = ss;
        const int r = e * TB + lane;
        if (io.vout != nullptr && r < io.rows.nrows) {
#pragma unroll
          for (int c = 0; c < 32; ++c) io.vout[(size_t)r * io.ldv + k * TB + c] = ps[c];
        }
      }
    }
    GPF_STAMP(2);
    // ---------------- phase 3: trailing updates (DMMA), with look-ahead on the next diagonal tile ----------------
    // Warp 0 updates tile (k+1,k+1) -- which only needs the panel slot it has just solved itself -- and factors it at
    // once, so the serial 32-column potrf of step k+1 runs under the other warps' tensor work.  Neither side waits
    // for work it does not depend on: every panel slot carries a ready flag (a consumer waits for exactly the two slots
    // of its tile, not for the whole panel), and one named barrier says "nobody reads the diagonal factor of step k any
    // more" (the others only arrive; warp 0 waits on it before overwriting DT).  The remaining tiles are handed out
    // dynamically; warp 0 joins when done.
    const int T1 = nreg * (nreg + 1) / 2;
    const int T2 = nx * nreg;
    const int T3 = (AUG == AUG_IDENT) ? (k + 1) * (k + 2) / 2 : 0;
    if (nreg > 0) {
      if (warp == 0) {
        const int i = k + 1;
        CFrag f;
        if (k == 0) {
          src.cfrag(f, i, i, g, tig);
          cfrag_add_diag(f, sm.dvec + i * TB, g, tig);
        } else cfrag_load(f, W + (size_t)plt(i, i) * TILE, g, tig, 4);
        tile_mma<true>(f, sm.P + (size_t)i * TB * PS, sm.P + (size_t)i * TB * PS, g, tig, 4);
        GPF_STAMP(4);
        bar_sync<NTE>(bar0 + 2);
        cfrag_to_DT(f, sm.DT, g, tig);
#ifndef GPF_DBG_NOLOOK
        diag_factor<FAST_TRSM, KEEPL>(i, lane, W, sm);
#endif
        GPF_STAMP(7);
      } else {
        bar_arrive<NTE>(bar0 + 2);
      }
    } else eng_sync();
    GPF_STAMP(3);
    // dynamic hand-out of the tiles; the ticket for the next tile is drawn before working on the current one so
    // that its accumulator tile can be prefetched under the tensor work
    const int TT = T1 + T2 + T3;
    auto ticket = [&]() {
      int t = 0;
      if (lane == 0) t = atomicAdd(tilectr, 1);
      return __shfl_sync(FULLMASK, t, 0);
    };
    auto t1_decode = [&](int t, int& i, int& j) {
      int ii = 0, tt = t;
      while (tt > ii) { tt -= ii + 1; ++ii; }
      i = k + 1 + ii; j = k + 1 + tt;
    };
    auto tile_ptr = [&](int t) -> const double* {   // accumulator tile of ticket t in the workspace (nullptr: no load)
      if (t >= TT || k == 0) return nullptr;
      if (t < T1) {
        int i, j;
        t1_decode(t, i, j);
        return W + (size_t)plt(i, j) * TILE;
      }
      if (t < T1 + T2) {
        const int u = t - T1;
        const int e = u / nreg, j = k + 1 + (u % nreg);
        if (AUG == AUG_RHS) return Wx + (size_t)(e * NT + j) * TILE;
        return (k == e) ? nullptr : W + (size_t)plt(j, e) * TILE;
      }
      int aa = 0, tt = t - T1 - T2;
      while (tt > aa) { tt -= aa + 1; ++aa; }
      return (k == aa) ? nullptr : W + (size_t)plt(aa, tt) * TILE;
    };
    int t = ticket();
    while (t < TT) {
      const int tn = ticket();
#ifndef GPF_DBG_NOPREF
      const double* pf = tile_ptr(tn);
      if (pf) prefetch_tile(pf, lane);
#endif
      CFrag f;
      if (t < T1) {
        int i, j;
        t1_decode(t, i, j);
        if (k == 0) {
          src.cfrag(f, i, j, g, tig);
          if (i == j) cfrag_add_diag(f, sm.dvec + i * TB, g, tig);
        } else cfrag_load(f, W + (size_t)plt(i, j) * TILE, g, tig, 4);
        slot_wait(sm.ready + i, k + 1);
        slot_wait(sm.ready + j, k + 1);
        tile_mma<true>(f, sm.P + (size_t)i * TB * PS, sm.P + (size_t)j * TB * PS, g, tig, 4);
        cfrag_store(f, W + (size_t)plt(i, j) * TILE, g, tig, 4);
      } else if (t < T1 + T2) {
        const int u = t - T1;
        const int e = u / nreg, j = k + 1 + (u % nreg);
        double* wt = (AUG == AUG_RHS) ? Wx + (size_t)(e * NT + j) * TILE : W + (size_t)plt(j, e) * TILE;
        if (AUG == AUG_RHS) {
          const int nti = (e == NE - 1) ? nti_last : 4;
          if (k == 0) cfrag_from_rows(f, io.rows, e * TB, j * TB, g, tig, nti);
          else cfrag_load(f, wt, g, tig, nti);
          slot_wait(sm.ready + NT + e, k + 1);
          slot_wait(sm.ready + j, k + 1);
          tile_mma<true>(f, sm.P + (size_t)(NT + e) * TB * PS, sm.P + (size_t)j * TB * PS, g, tig, nti);
          cfrag_store(f, wt, g, tig, nti);
        } else {
          if (k == e) cfrag_zero(f);
          else cfrag_load(f, wt, g, tig, 4);
          slot_wait(sm.ready + e, k + 1);
          slot_wait(sm.ready + j, k + 1);
          tile_mma<true>(f, sm.P + (size_t)e * TB * PS, sm.P + (size_t)j * TB * PS, g, tig, 4);
          cfrag_store(f, wt, g, tig, 4);
        }
      } else {
        // (L L^T)^-1 tile (a,b) += U_ak U_bk^T, first touched at k == a
        int aa = 0, tt = t - T1 - T2;
        while (tt > aa) { tt -= aa + 1; ++aa; }
        const int bb = tt;
        double* kt = W + (size_t)plt(aa, bb) * TILE;
        if (k == aa) cfrag_zero(f);
        else cfrag_load(f, kt, g, tig, 4);
        slot_wait(sm.ready + aa, k + 1);
        slot_wait(sm.ready + bb, k + 1);
        tile_mma<false>(f, sm.P + (size_t)aa * TB * PS, sm.P + (size_t)bb * TB * PS, g, tig, 4);
        cfrag_store(f, kt, g, tig, 4);
      }
      t = tn;
    }
    GPF_STAMP(5);
  }
  eng_sync();
  GPF_STAMP(1);
  GPF_STAMP_FLUSH
  if (sm.scal[1] != 0.0) return false;
  if constexpr (NWE == NW) {
    double lg = 0.0;
    for (int t = tid; t < NT * TB; t += NTHR) lg += log(sm.pivall[t]);
    lg = block_sum(lg, sm.red());
    if (tid == 0) sm.scal[0] = lg;
    __syncthreads();
  } else {
    if (warp == 0) {
      double lg = 0.0;
      for (int t = lane; t < NT * TB; t += 32) lg += log(sm.pivall[t]);
      lg = warp_sum(lg);
      if (lane == 0) sm.scal[0] = lg;
    }
    eng_sync();
  }
  return true;
}

// sum of sm.quad over the NE*NT partials in fixed order (call after the engine; all threads get the value)
__device__ __forceinline__ double engine_quad_total(const EngineSmem& sm, int NT, int NE) {
  double s = 0.0;
  for (int i = 0; i < NE * NT; ++i) s += sm.quad[i];
  return s;
}

/* Back substitution  L^T x = w  for NRHS vectors held in shared memory (xs[v*ldx + i], i < NT*32, in place).
 * L: tiles left in W by the engine (KEEPL); rinv holds 1/L_ii.
 * Right-looking by tile rows with ONE barrier per step: warp 0 solves the 32x32 diagonal block of step k (vectors
 * interleaved for ILP), the barrier publishes x_k, then warp 0 applies the urgent update t_{k-1} -= L(k,k-1)^T x_k
 * itself and goes straight on to the next diagonal solve while the other warps apply the updates of the tile rows
 * further up (each tile by one warp: no cross-warp reduction).
 * Called by the NWE warps of one (sub-)engine (threads sub*NWE*32 ...), which synchronise on the engine's barrier. */
template <int NRHS, int NWE>
__device__ void back_solve_g(const double* W, int NT, double* xs, int ldx, const double* rinv_all, int sub) {
  constexpr int NTE = NWE * 32;
  const int tid = threadIdx.x - sub * NTE, warp = tid >> 5, lane = tid & 31;
  auto gsync = [&]() {
    if constexpr (NWE == NW) __syncthreads();
    else bar_sync<NTE>(1 + 3 * sub);
  };
  // t_i -= L(k,i)^T x_k for one tile (k,i): lane = column of the tile = element of t_i
  auto update = [&](int k, int i) {
    const double* tile = W + (size_t)plt(k, i) * TILE;
    double acc[NRHS];
#pragma unroll
    for (int v = 0; v < NRHS; ++v) acc[v] = 0.0;
#pragma unroll 8
    for (int r = 0; r < 32; ++r) {
      const double l = tile[r * TB + lane];
#pragma unroll
      for (int v = 0; v < NRHS; ++v) acc[v] = fma(l, xs[v * ldx + k * TB + r], acc[v]);
    }
#pragma unroll
    for (int v = 0; v < NRHS; ++v) xs[v * ldx + i * TB + lane] -= acc[v];
  };
  gsync();
  for (int k = NT - 1; k >= 0; --k) {
    if (warp == 0) {
      const double* tile = W + (size_t)plt(k, k) * TILE;
      double a[32];   // a[r] = L[r][lane]
#pragma unroll
      for (int r = 0; r < 32; ++r) a[r] = tile[r * TB + lane];
      if (k > 0) prefetch_tile(W + (size_t)plt(k, k - 1) * TILE, lane);
      const double* rinv = rinv_all + k * TB;
      double t[NRHS], x[NRHS];
#pragma unroll
      for (int v = 0; v < NRHS; ++v) { t[v] = xs[v * ldx + k * TB + lane]; x[v] = 0.0; }
#pragma unroll
      for (int c = 31; c >= 0; --c) {
        const double rc = rinv[c];
#pragma unroll
        for (int v = 0; v < NRHS; ++v) {
          const double xc = __shfl_sync(FULLMASK, t[v], c) * rc;
          if (lane == c) x[v] = xc;
          if (lane < c) t[v] = fma(-a[c], xc, t[v]);
        }
      }
#pragma unroll
      for (int v = 0; v < NRHS; ++v) xs[v * ldx + k * TB + lane] = x[v];
    }
    gsync();                               // x_k is published; every update of earlier steps has been applied
    if (k == 0) break;
    if (NWE == 1) {
      for (int i = k - 1; i >= 0; --i) update(k, i);
      __syncwarp();
    } else if (warp == 0) {
      update(k, k - 1);                    // urgent: the next diagonal solve needs it
      __syncwarp();
    } else {
      for (int i = k - 1 - warp; i >= 0; i -= (NWE - 1)) update(k, i);   // tiles (k, k-2), (k, k-3), ... over warps 1..
    }
  }
  gsync();
}

template <int NRHS>
__device__ void back_solve(const double* W, int NT, double* xs, int ldx, const EngineSmem& sm, double* red) {
  (void)red;
  back_solve_g<NRHS, NW>(W, NT, xs, ldx, sm.rinv, 0);
}

/* Triangular matrix-vector products on the tensor pipe:  y = L x  (TRANS = false)  or  y = L^T x  (TRANS = true)  for up to
 * eight vectors at once.  L: NT x NT lower triangular, packed by tile (plt), tiles row-major 32 x 32 in global memory
 * (L2-resident), diagonal tiles with a zero upper triangle -- the layout the engine leaves with KEEPL.  x, y: shared memory,
 * vector v at x + v*ld (NT*32 entries); only the first R vectors are read and written (the others count as zero).
 * Every 32 x 32 tile is 32 DMMA.8x8x4: the A fragments come straight from the tile in memory (every load instruction
 * touches whole 32-byte sectors), B is the vector block, so one pass over the tiles serves all R vectors.  Output tile
 * rows are dealt to the NWE warps of the calling (sub-)engine; the caller synchronises before and after. */
template <bool TRANS, int NWE>
__device__ __forceinline__ void trimv_dmma(const double* __restrict__ L, int NT, const double* x, double* y, int ld, int R,
                                           int warp, int lane) {
  const int g = lane >> 2, tig = lane & 3;
  for (int o = warp; o < NT; o += NWE) {
    double c[4][2];
#pragma unroll
    for (int t = 0; t < 4; ++t) c[t][0] = c[t][1] = 0.0;
    const int q0 = TRANS ? o : 0, q1 = TRANS ? NT - 1 : o;
    for (int q = q0; q <= q1; ++q) {
      const double* tile = L + (size_t)(TRANS ? plt(q, o) : plt(o, q)) * TILE;
      const double* xv = x + (size_t)g * ld + q * TB + tig;      // B operand: element k = 4kk + tig of vector g
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        const double b = (g < R) ? xv[4 * kk] : 0.0;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const double a = TRANS ? tile[(4 * kk + tig) * TB + 8 * t + g] : tile[(8 * t + g) * TB + 4 * kk + tig];
          dmma884(c[t][0], c[t][1], a, b);
        }
      }
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int row = o * TB + 8 * t + g;
      if (2 * tig < R) y[(size_t)(2 * tig) * ld + row] = c[t][0];
      if (2 * tig + 1 < R) y[(size_t)(2 * tig + 1) * ld + row] = c[t][1];
    }
  }
}

// ---- scaled first-touch sources:  I + S B S  with S = s I (one precision for all time points) or S = diag(sd) ----
// (the additive identity comes through the engine's dvec; padding stays the identity)
template <class Inner>
struct SrcScaled {
  Inner in;
  double s;     // element (R, C) = s * in(R, C)
  int nn;       // matrix size (padding beyond)
  __device__ __forceinline__ double at(int R, int Cc) const {
    if (R >= nn || Cc >= nn) return (R == Cc) ? 1.0 : 0.0;
    return s * in.at(R, Cc);
  }
  __device__ __forceinline__ void rows(double (&a)[32], int bi, int bj, int lane) const {
#pragma unroll
    for (int c = 0; c < 32; ++c) a[c] = at(bi * TB + lane, bj * TB + c);
  }
  __device__ __forceinline__ void cfrag(CFrag& f, int bi, int bj, int g, int tig) const {
#pragma unroll
    for (int ti = 0; ti < 4; ++ti)
#pragma unroll
      for (int tj = 0; tj < 4; ++tj) {
        f.c[ti][tj][0] = at(bi * TB + 8 * ti + g, bj * TB + 8 * tj + 2 * tig);
        f.c[ti][tj][1] = at(bi * TB + 8 * ti + g, bj * TB + 8 * tj + 2 * tig + 1);
      }
  }
  __device__ __forceinline__ void fill_slot(double* slot, int bi, int bj, int lane) const {
#pragma unroll 8
    for (int r = 0; r < 32; ++r) slot[r * PS + lane] = at(bi * TB + r, bj * TB + lane);
  }
};
template <class Inner>
struct SrcDiagScaled {
  Inner in;
  const double* sd;   // element (R, C) = sd[R] * in(R, C) * sd[C]
  int nn;
  __device__ __forceinline__ double at(int R, int Cc) const {
    if (R >= nn || Cc >= nn) return (R == Cc) ? 1.0 : 0.0;
    return sd[R] * in.at(R, Cc) * sd[Cc];
  }
  __device__ __forceinline__ void rows(double (&a)[32], int bi, int bj, int lane) const {
#pragma unroll
    for (int c = 0; c < 32; ++c) a[c] = at(bi * TB + lane, bj * TB + c);
  }
  __device__ __forceinline__ void cfrag(CFrag& f, int bi, int bj, int g, int tig) const {
#pragma unroll
    for (int ti = 0; ti < 4; ++ti)
#pragma unroll
      for (int tj = 0; tj < 4; ++tj) {
        f.c[ti][tj][0] = at(bi * TB + 8 * ti + g, bj * TB + 8 * tj + 2 * tig);
        f.c[ti][tj][1] = at(bi * TB + 8 * ti + g, bj * TB + 8 * tj + 2 * tig + 1);
      }
  }
  __device__ __forceinline__ void fill_slot(double* slot, int bi, int bj, int lane) const {
#pragma unroll 8
    for (int r = 0; r < 32; ++r) slot[r * PS + lane] = at(bi * TB + r, bj * TB + lane);
  }
};

}  // namespace gpf
