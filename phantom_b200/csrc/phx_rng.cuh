// phx_rng.cuh -- device twin of the counter-based RNG contract (oracle/rng.py).
//
// Replaces the reference's process-global draw sites (np.random.randint at
// examples/environments/supply_chain/supply_chain.py:64, np.random.shuffle at
// phantom/resolvers.py:151, ...) with a stateless stream:
//
//   u32(seed, env, episode, step, stream, idx) =
//       Philox4x32-10(key = (seed_lo, seed_hi),
//                     ctr = (env, episode, step, (stream << 16) | (idx >> 2)))[idx & 3]
//
// One Philox block therefore serves four consecutive idx of a stream.
#pragma once
#include <stdint.h>

namespace phx {

struct Philox4 {
  uint32_t w[4];
};

// Philox4x32-10, Random123 constants (Salmon et al., SC'11).
__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2,
                                                 uint32_t c3, uint32_t k0, uint32_t k1) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  constexpr uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    c0 = hi1 ^ c1 ^ k0;
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k1;
    c3 = lo0;
    k0 += W0;
    k1 += W1;
  }
  Philox4 out;
  out.w[0] = c0; out.w[1] = c1; out.w[2] = c2; out.w[3] = c3;
  return out;
}

// The block holding idx = 4*block .. 4*block+3 of `stream`.
__device__ __forceinline__ Philox4 rng_block(uint64_t seed, uint32_t env, uint32_t episode,
                                             uint32_t step, uint32_t stream, uint32_t block) {
  return philox4x32_10(env, episode, step, (stream << 16) | (block & 0xFFFFu),
                       (uint32_t)seed, (uint32_t)(seed >> 32));
}

__device__ __forceinline__ uint32_t rng_u32(uint64_t seed, uint32_t env, uint32_t episode,
                                            uint32_t step, uint32_t stream, uint32_t idx) {
  const Philox4 b = rng_block(seed, env, episode, step, stream, idx >> 2);
  return b.w[idx & 3];
}

// randint(n) := (u32 * n) >> 32   (replaces np.random.randint(n))
__device__ __forceinline__ int rng_randint(uint32_t word, uint32_t n) {
  return (int)__umulhi(word, n);
}

// uniform01() := (u32 >> 8) * 2^-24, float32-exact, in [0, 1)
__device__ __forceinline__ float rng_uniform01(uint32_t word) {
  return (float)(word >> 8) * 5.9604644775390625e-8f;
}

}  // namespace phx
