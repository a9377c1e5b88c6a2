// Philox4x32-10 keyed stream shared with oracle/gpfanova_oracle.py (Stream class).
// counter = (draw, sampler_id, sweep, chain), key = (seed_lo, seed_hi).
#pragma once
#include <stdint.h>

namespace gpf {

struct Philox4 { uint32_t x, y, z, w; };

__host__ __device__ __forceinline__ void philox_mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
  uint64_t p = (uint64_t)a * (uint64_t)b;
  hi = (uint32_t)(p >> 32);
  lo = (uint32_t)p;
}

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                          uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0, lo0, hi1, lo1;
    philox_mulhilo(0xD2511F53u, c0, hi0, lo0);
    philox_mulhilo(0xCD9E8D57u, c2, hi1, lo1);
    uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  Philox4 o; o.x = c0; o.y = c1; o.z = c2; o.w = c3;
  return o;
}

// 53-bit uniform strictly inside (0,1)
__host__ __device__ __forceinline__ double u01(uint32_t hi, uint32_t lo) {
  return ((double)(hi >> 6) * 134217728.0 + (double)(lo >> 5) + 0.5) * (1.0 / 9007199254740992.0);
}

struct Stream {
  uint32_t k0, k1;
  __host__ __device__ Stream(uint64_t seed) : k0((uint32_t)seed), k1((uint32_t)(seed >> 32)) {}
  __host__ __device__ Stream(uint32_t a, uint32_t b) : k0(a), k1(b) {}
  __device__ __forceinline__ double uniform(uint32_t chain, uint32_t sweep, uint32_t sampler, uint32_t draw) const {
    Philox4 r = philox4x32_10(draw, sampler, sweep, chain, k0, k1);
    return u01(r.x, r.y);
  }
  // Box-Muller pair for counter `pair`: z[2*pair], z[2*pair+1]
  __device__ __forceinline__ void normal_pair(uint32_t chain, uint32_t sweep, uint32_t sampler, uint32_t pair,
                                              double& z0, double& z1) const {
    Philox4 r = philox4x32_10(pair, sampler, sweep, chain, k0, k1);
    double ua = u01(r.x, r.y), ub = u01(r.z, r.w);
    double rad = sqrt(-2.0 * log(ua));
    double s, c;
    sincospi(2.0 * ub, &s, &c);
    z0 = rad * c;
    z1 = rad * s;
  }
};

}  // namespace gpf
