// lds_philox.cuh — counter-based Philox4x32-10 (Salmon et al., SC'11) and the element<->draw mapping.
// Must match oracle/philox.py (the CPU oracle regenerates the same uniforms from it).
//   key     = (seed lo, seed hi)
//   counter = (c0, c1, step lo, (step hi16 << 16) | stream << 12 | sample & 0xfff)
//   edges   : canonical (a,b) = (min,max); c0 = b/2, c1 = a/2; word = 2*(a%2) + (b%2)   (2x2 block per call)
//   dropout : c0 = col/4, c1 = row; word = col%4
//   uniform = (word >> 8) * 2^-24  in [0,1)
#pragma once
#include <stdint.h>

namespace lds {

struct PhiloxKey { uint32_t k0, k1, c2, c3; };

__host__ __device__ inline PhiloxKey philox_key(uint64_t seed, uint64_t step, uint32_t stream, uint32_t sample) {
  PhiloxKey k;
  k.k0 = (uint32_t)(seed & 0xffffffffull);
  k.k1 = (uint32_t)(seed >> 32);
  k.c2 = (uint32_t)(step & 0xffffffffull);
  k.c3 = (uint32_t)((((step >> 32) & 0xffffull) << 16) | ((uint64_t)(stream & 0xfu) << 12) | (sample & 0xfffu));
  return k;
}

__host__ __device__ inline void philox4x32_10(uint32_t c0, uint32_t c1, const PhiloxKey& key, uint32_t out[4]) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  uint32_t c2 = key.c2, c3 = key.c3, k0 = key.k0, k1 = key.k1;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)M0 * c0;
    const uint64_t p1 = (uint64_t)M1 * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    c1 = (uint32_t)p1;
    c3 = (uint32_t)p0;
    c0 = n0;
    c2 = n2;
    k0 += W0;
    k1 += W1;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// Same generator with the 10 round keys precomputed on the host (they depend only on the seed): 4 instructions per
// round (2 x IMAD.WIDE, 2 x LOP3) instead of 6 on the sampling kernel's critical path.
struct PhiloxRounds { uint32_t rk[10][2]; uint32_t c2, c3; };

inline PhiloxRounds philox_rounds(const PhiloxKey& key) {
  PhiloxRounds r;
  uint32_t k0 = key.k0, k1 = key.k1;
  for (int i = 0; i < 10; ++i) { r.rk[i][0] = k0; r.rk[i][1] = k1; k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
  r.c2 = key.c2; r.c3 = key.c3;
  return r;
}

// c2 / c3 (the step / stream / sample words of the counter) are passed explicitly so a kernel can take the step from
// device memory (lds_k1_sample_normalize_dstep); the plain overload below uses the host-prepared ones.
__device__ __forceinline__ void philox4x32_10_rk(uint32_t c0, uint32_t c1, const PhiloxRounds& R, uint32_t c2, uint32_t c3, uint32_t out[4]) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)M0 * c0;
    const uint64_t p1 = (uint64_t)M1 * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ R.rk[r][0];
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ R.rk[r][1];
    c1 = (uint32_t)p1;
    c3 = (uint32_t)p0;
    c0 = n0;
    c2 = n2;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__device__ __forceinline__ void philox4x32_10_rk(uint32_t c0, uint32_t c1, const PhiloxRounds& R, uint32_t out[4]) {
  philox4x32_10_rk(c0, c1, R, R.c2, R.c3, out);
}

// Rolled variant for code that runs once per launch on a few warps (row epilogues): 10x less instruction fetch.
static __device__ __noinline__ void philox4x32_10_rolled(uint32_t c0, uint32_t c1, PhiloxKey key, uint32_t* out) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  uint32_t c2 = key.c2, c3 = key.c3, k0 = key.k0, k1 = key.k1;
#pragma unroll 1
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)M0 * c0;
    const uint64_t p1 = (uint64_t)M1 * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    c1 = (uint32_t)p1;
    c3 = (uint32_t)p0;
    c0 = n0;
    c2 = n2;
    k0 += W0;
    k1 += W1;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__host__ __device__ inline float philox_to_uniform(uint32_t w) { return (float)(w >> 8) * 5.9604644775390625e-8f; }

}  // namespace lds
