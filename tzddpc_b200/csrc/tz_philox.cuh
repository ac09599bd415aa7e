// Philox4x32-10 counter-based generator and the draw convention of the TZDDPC random streams (see tz_rng.cu).
#pragma once
#include "tz_common.cuh"

namespace tz {

__host__ __device__ __forceinline__ void philox4x32_10(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                       uint32_t (&out)[4]) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
    const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// draw number j (0, 1, 2, ...) of (scenario, t, purpose): block j / 2 of the counter, half j % 2 of its output
__device__ __forceinline__ double draw(uint64_t seed, uint64_t scenario, uint32_t t, uint32_t purpose, int j, bool vertex) {
  uint32_t o[4];
  philox4x32_10((uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)scenario, (uint32_t)(scenario >> 32), t,
                (purpose << 16) | (uint32_t)(j >> 1), o);
  const uint64_t bits = (j & 1) ? (((uint64_t)o[3] << 32) | o[2]) : (((uint64_t)o[1] << 32) | o[0]);
  if (vertex) return (bits >> 63) ? 1.0 : -1.0;
  return 2.0 * ((double)(bits >> 11) * (1.0 / 9007199254740992.0)) - 1.0;
}

}  // namespace tz
