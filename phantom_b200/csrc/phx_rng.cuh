// phx_rng.cuh -- device twin of the counter-based RNG contract (oracle/rng.py).
//
// Replaces the reference's process-global draw sites (np.random.randint at
// examples/environments/supply_chain/supply_chain.py:64, np.random.shuffle at
// phantom/resolvers.py:151, ...) with a stateless stream of 24-bit draws:
//
//   d24(seed, env, episode, step, stream, idx) = slot (idx % 5) of
//       Philox4x32-10(key = (seed_lo, seed_hi),
//                     ctr = (env, episode, step, (stream << 16) | (idx / 5)))
//
// One 128-bit block (w0..w3) yields FIVE draws: slots 0..3 = w_k >> 8, slot 4 = the low
// bytes of w0, w1, w2.  24 bits = a float32 mantissa (uniform01 is exact), and five draws
// per block is what lets ONE Philox block serve the five customers of the supply chain.
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

// The block holding draws idx = 5*block .. 5*block+4 of `stream`.
__device__ __forceinline__ Philox4 rng_block(uint64_t seed, uint32_t env, uint32_t episode,
                                             uint32_t step, uint32_t stream, uint32_t block) {
  return philox4x32_10(env, episode, step, (stream << 16) | (block & 0xFFFFu),
                       (uint32_t)seed, (uint32_t)(seed >> 32));
}

// Draw `slot` (0..4) of a block, LEFT-ALIGNED in 32 bits (d24 << 8): with the draw in the
// top 24 bits, randint is a single multiply-high.
__device__ __forceinline__ uint32_t rng_slot_hi(const Philox4& b, int slot) {
  if (slot < 4) return b.w[slot] & 0xFFFFFF00u;
  // ((w0 & 0xff) << 16 | (w1 & 0xff) << 8 | (w2 & 0xff)) << 8: two byte-permutes + mask
  const uint32_t t = __byte_perm(b.w[2], b.w[1], 0x7407);  // bytes [x, w2.b0, w1.b0, x]
  return __byte_perm(t, b.w[0], 0x4210) & 0xFFFFFF00u;      // bytes [0, w2.b0, w1.b0, w0.b0]
}

__device__ __forceinline__ uint32_t rng_d24_hi(uint64_t seed, uint32_t env, uint32_t episode,
                                               uint32_t step, uint32_t stream, uint32_t idx) {
  const Philox4 b = rng_block(seed, env, episode, step, stream, idx / 5u);
  const uint32_t slot = idx % 5u;
  uint32_t r = rng_slot_hi(b, 4);
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (slot == (uint32_t)k) r = rng_slot_hi(b, k);
  return r;
}

// randint(n) := (d24 * n) >> 24 == umulhi(d24 << 8, n)   (replaces np.random.randint(n))
__device__ __forceinline__ int rng_randint(uint32_t d24_hi, uint32_t n) {
  return (int)__umulhi(d24_hi, n);
}

// uniform01() := d24 * 2^-24, float32-exact, in [0, 1)
__device__ __forceinline__ float rng_uniform01(uint32_t d24_hi) {
  return (float)(d24_hi >> 8) * 5.9604644775390625e-8f;
}

// ---------------------------------------------------------------------------------------
// Packed small-integer draws (contract v3, oracle/rng.py "packed draws").
//
// A call site that draws K values of randint(n) per step with a small n (the supply chain's
// customers: K = 5, n = 5) does not need 24 bits per draw.  Its draws are the base-n digits of
// 32-bit Philox words, extracted by multiply-high (exact arithmetic decoding of the fraction
// word / 2^32):
//     x_0 = word;   digit_r = (x_r * n) >> 32;   x_{r+1} = (x_r * n) mod 2^32
// A word yields kpw(n) digits, kpw = the largest j with n^j <= 2^16 (so every digit's
// distribution is within 2^-16 relative of uniform; n = 5: 6 digits).  The site consumes
//     W = ceil(K / kpw) words per step;   draw i of step s = digit (i % kpw) of word number
//     g = s * W + i / kpw  of the site's word sequence;
//     word g = w[g & 3] of Philox4x32-10(key = seed,
//                                        ctr = (env, episode, g >> 2, (stream << 16) | 0x8000))
// so consecutive steps share Philox blocks: with K <= kpw one block serves FOUR steps.
__host__ __device__ inline int rng_digits_per_word(uint32_t n) {
  int j = 1;
  uint64_t pw = n;
  while (pw * n <= 65536ull) {
    pw *= n;
    ++j;
  }
  return j;
}

__device__ __forceinline__ Philox4 rng_word_block(uint64_t seed, uint32_t env, uint32_t episode,
                                                  uint32_t block, uint32_t stream) {
  return philox4x32_10(env, episode, block, (stream << 16) | 0x8000u, (uint32_t)seed,
                       (uint32_t)(seed >> 32));
}

// Per-thread cache of the current block of one site's word sequence.
struct PackedWords {
  Philox4 blk;
  uint32_t block = 0xFFFFFFFFu, episode = 0xFFFFFFFFu;

  // word number g of (env, episode, stream)
  __device__ __forceinline__ uint32_t word(uint64_t seed, uint32_t env, uint32_t ep, uint32_t g,
                                           uint32_t stream) {
    const uint32_t b = g >> 2;
    if (b != block || ep != episode) {
      blk = rng_word_block(seed, env, ep, b, stream);
      block = b;
      episode = ep;
    }
    const uint32_t q = g & 3u;
    uint32_t x = blk.w[0];
    if (q == 1u) x = blk.w[1];
    if (q == 2u) x = blk.w[2];
    if (q == 3u) x = blk.w[3];
    return x;
  }
};

// Sequential reader of one site's word sequence for K <= kpw sites (one word per step): the
// four words of a block are handed out in step order and the next block is fetched when they
// run out -- no per-step block compare or word select.  Steps must be consumed consecutively;
// call flush() when the (episode, step) coordinates jump (reset / auto-reset wrap).
struct PackedWordQueue {
  uint32_t w0, w1, w2, w3;
  int left = 0;

  __device__ __forceinline__ void flush() { left = 0; }

  // the word of `step` (word number g = step)
  __device__ __forceinline__ uint32_t take(uint64_t seed, uint32_t env, uint32_t ep, uint32_t step,
                                           uint32_t stream) {
    if (left == 0) {
      const Philox4 b = rng_word_block(seed, env, ep, step >> 2, stream);
      w0 = b.w[0]; w1 = b.w[1]; w2 = b.w[2]; w3 = b.w[3];
      const uint32_t skip = step & 3u;  // 0 except right after a flush
      if (skip >= 1u) { w0 = w1; w1 = w2; w2 = w3; }
      if (skip >= 2u) { w0 = w1; w1 = w2; }
      if (skip >= 3u) { w0 = w1; }
      left = 4 - (int)skip;
    }
    const uint32_t x = w0;
    w0 = w1; w1 = w2; w2 = w3;
    --left;
    return x;
  }
};

// next digit of x: returns it and advances x
__device__ __forceinline__ int rng_next_digit(uint32_t& x, uint32_t n) {
  uint64_t prod;  // one IMAD.WIDE.U32 (plain C++ widens n to 64 bits and adds a zero high part)
  asm("mul.wide.u32 %0, %1, %2;" : "=l"(prod) : "r"(x), "r"(n));
  x = (uint32_t)prod;
  return (int)(prod >> 32);
}

// draw i of a packed site, stand-alone (one block + i % kpw + 1 multiplies); kpw =
// rng_digits_per_word(n) and W = ceil(K / kpw) are precomputed on the host
__device__ __forceinline__ int rng_packed_randint(uint64_t seed, uint32_t env, uint32_t episode,
                                                  uint32_t step, uint32_t stream, uint32_t n,
                                                  uint32_t kpw, uint32_t W, uint32_t i) {
  const uint32_t g = step * W + i / kpw;
  const Philox4 b = rng_word_block(seed, env, episode, g >> 2, stream);
  const uint32_t q = g & 3u;
  uint32_t x = b.w[0];
  if (q == 1u) x = b.w[1];
  if (q == 2u) x = b.w[2];
  if (q == 3u) x = b.w[3];
  int d = 0;
  for (uint32_t r = 0; r <= i % kpw; ++r) d = rng_next_digit(x, n);
  return d;
}

}  // namespace phx
