"""TEST INFRASTRUCTURE (CPU oracle) -- never imported by the product.

The counter-based RNG contract shared by the oracle and the CUDA kernels
(device twin: phantom_b200/csrc/phx_rng.cuh).

The reference draws from process-global generators (`np.random.randint` at
examples/environments/supply_chain/supply_chain.py:64, `np.random.shuffle` at
phantom/resolvers.py:151, ...), which cannot be reproduced across 65 536
concurrently stepped envs.  "Identical seeds" therefore means: both sides consume the
same stateless stream of 24-bit draws

    d24(seed, env, episode, step, stream, idx) = slot (idx % 5) of
        Philox4x32-10( key = (seed & 0xffffffff, seed >> 32),
                       ctr = (env, episode, step, (stream << 16) | (idx // 5)) )

  one 128-bit Philox block (w0..w3) yields FIVE 24-bit draws:
      slot k < 4 :  w_k >> 8
      slot 4     :  (w0 & 0xff) << 16 | (w1 & 0xff) << 8 | (w2 & 0xff)
  (24 bits = a float32 mantissa, so uniform01 is exact; 5 draws per block instead of 4
  32-bit words is what lets one block serve the five customers of the supply chain)

  env     global env index (independent of how envs are sharded over GPUs)
  episode number of resets this env has seen minus one (0 for the first episode)
  step    PhantomEnv.current_step *after* the increment at the top of step()
          (env.py:252), i.e. 1..num_steps; 0 is used for draws made during reset()
  stream  draw-site id chosen by the workload (one per distinct RNG call site)
  idx     index of the draw within that site and step (e.g. the customer index)

Derived distributions:
  randint(n)  := (d24 * n) >> 24          (multiply-shift; replaces np.random.randint(n);
                                           bias <= n / 2^24)
  uniform01() := d24 * 2**-24             (float32-exact, in [0, 1))

Packed small-integer draws (contract v3).  A call site that draws K values of randint(n)
per step with a small n (the supply chain's customers, supply_chain.py:64: K = 5, n = 5)
takes its draws as the base-n digits of 32-bit Philox words, extracted by multiply-high
(exact arithmetic decoding of the fraction word / 2^32):

      x_0 = word;   digit_r = (x_r * n) >> 32;   x_{r+1} = (x_r * n) mod 2^32

  a word yields kpw(n) digits, kpw = the largest j with n^j <= 2^16 (every digit is then
  within 2^-16 relative of uniform -- the same order as the 24-bit draws' n / 2^24; n = 5:
  6 digits).  The site consumes W = ceil(K / kpw) words per step; draw i of step s is digit
  i % kpw of word number g = s * W + i // kpw of the site's word sequence, and

      word g = w[g & 3] of Philox4x32-10( key = seed,
                   ctr = (env, episode, g >> 2, (stream << 16) | 0x8000) )

  so consecutive steps share blocks: with K <= kpw ONE Philox block serves FOUR env steps
  (the device's fused step kernel spent 30 % of its instructions on one block per step).

Philox4x32-10 is Salmon et al., "Parallel random numbers: as easy as 1, 2, 3" (SC'11),
with the standard Random123 constants; checked below against the Random123 known-answer
vectors.
"""
from __future__ import annotations

import numpy as np

M0 = 0xD2511F53
M1 = 0xCD9E8D57
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = 0xFFFFFFFF


def philox4x32(ctr, key, rounds: int = 10):
    """Scalar Philox4x32 on Python ints. ctr = 4 x u32, key = 2 x u32."""
    c0, c1, c2, c3 = (int(c) & MASK for c in ctr)
    k0, k1 = (int(k) & MASK for k in key)
    for _ in range(rounds):
        p0 = M0 * c0
        p1 = M1 * c2
        c0, c1, c2, c3 = (
            ((p1 >> 32) ^ c1 ^ k0) & MASK,
            p1 & MASK,
            ((p0 >> 32) ^ c3 ^ k1) & MASK,
            p0 & MASK,
        )
        k0 = (k0 + W0) & MASK
        k1 = (k1 + W1) & MASK
    return c0, c1, c2, c3


def philox4x32_np(c0, c1, c2, c3, k0, k1, rounds: int = 10):
    """Vectorised Philox4x32 over numpy uint32 arrays (broadcasting)."""
    c0, c1, c2, c3, k0, k1 = (
        np.asarray(x).astype(np.uint64) & MASK for x in (c0, c1, c2, c3, k0, k1)
    )
    c0, c1, c2, c3, k0, k1 = np.broadcast_arrays(c0, c1, c2, c3, k0, k1)
    for _ in range(rounds):
        p0 = np.uint64(M0) * c0
        p1 = np.uint64(M1) * c2
        c0, c1, c2, c3 = (
            ((p1 >> np.uint64(32)) ^ c1 ^ k0) & np.uint64(MASK),
            p1 & np.uint64(MASK),
            ((p0 >> np.uint64(32)) ^ c3 ^ k1) & np.uint64(MASK),
            p0 & np.uint64(MASK),
        )
        k0 = (k0 + np.uint64(W0)) & np.uint64(MASK)
        k1 = (k1 + np.uint64(W1)) & np.uint64(MASK)
    return tuple(x.astype(np.uint32) for x in (c0, c1, c2, c3))


# ------------------------------------------------------------------ the contract proper
def _slots(w):
    """The five 24-bit draws of one block (works on ints and on numpy arrays)."""
    return (w[0] >> 8, w[1] >> 8, w[2] >> 8, w[3] >> 8,
            ((w[0] & 0xFF) << 16) | ((w[1] & 0xFF) << 8) | (w[2] & 0xFF))


def d24(seed: int, env: int, episode: int, step: int, stream: int, idx: int) -> int:
    words = philox4x32(
        (env, episode, step, ((stream & 0xFFFF) << 16) | ((idx // 5) & 0xFFFF)),
        (seed & MASK, (seed >> 32) & MASK),
    )
    return _slots(words)[idx % 5]


def d24_np(seed: int, env, episode, step, stream: int, idx):
    """Vectorised contract draw; env/episode/step/idx broadcast against each other."""
    idx = np.asarray(idx, dtype=np.int64)
    c3 = ((stream & 0xFFFF) << 16) | ((idx // 5) & 0xFFFF)
    w = philox4x32_np(env, episode, step, c3, seed & MASK, (seed >> 32) & MASK)
    slots = _slots(w)
    sel = np.broadcast_to(idx % 5, w[0].shape)
    return np.choose(sel, slots).astype(np.uint32)


def randint(n: int, draw: int) -> int:
    return (int(draw) * int(n)) >> 24


def randint_np(n: int, draws) -> np.ndarray:
    return ((np.asarray(draws).astype(np.uint64) * np.uint64(n)) >> np.uint64(24)).astype(np.int32)


def uniform01(draw: int) -> float:
    return float(np.float32(int(draw)) * np.float32(2.0**-24))


def uniform01_np(draws) -> np.ndarray:
    return np.asarray(draws, dtype=np.uint32).astype(np.float32) * np.float32(2.0**-24)


# ------------------------------------------------------------------ packed draws (v3)
def digits_per_word(n: int) -> int:
    j, pw = 1, int(n)
    while pw * n <= 65536:
        pw *= n
        j += 1
    return j


def packed_word(seed: int, env: int, episode: int, stream: int, g: int) -> int:
    words = philox4x32((env, episode, (g >> 2) & MASK, ((stream & 0xFFFF) << 16) | 0x8000),
                       (seed & MASK, (seed >> 32) & MASK))
    return words[g & 3]


def packed_randint(seed: int, env: int, episode: int, step: int, stream: int, n: int, K: int,
                   i: int) -> int:
    """Draw i (0 <= i < K) of the K randint(n) draws this site makes in `step`."""
    kpw = digits_per_word(n)
    W = (K + kpw - 1) // kpw
    x = packed_word(seed, env, episode, stream, step * W + i // kpw)
    d = 0
    for _ in range(i % kpw + 1):
        prod = x * n
        d, x = prod >> 32, prod & MASK
    return d


def packed_randint_np(seed: int, env, episode, step, stream: int, n: int, K: int) -> np.ndarray:
    """All K draws of a step, vectorised: env/episode/step broadcast -> int32 [..., K]."""
    kpw = digits_per_word(n)
    W = (K + kpw - 1) // kpw
    env, episode, step = np.broadcast_arrays(np.asarray(env, np.int64), np.asarray(episode, np.int64),
                                             np.asarray(step, np.int64))
    out = np.zeros(env.shape + (K,), np.int32)
    for q in range(W):
        g = step * W + q
        w = philox4x32_np(env, episode, (g >> 2) & MASK, ((stream & 0xFFFF) << 16) | 0x8000,
                          seed & MASK, (seed >> 32) & MASK)
        x = np.choose(g & 3, w).astype(np.uint64)
        for r in range(kpw):
            i = q * kpw + r
            if i >= K:
                break
            prod = x * np.uint64(n)
            out[..., i] = (prod >> np.uint64(32)).astype(np.int32)
            x = prod & np.uint64(MASK)
    return out


class PackedStream:
    """Sequential view of a packed site: the i-th randint(n) call after begin() is draw i.
    `n` and the number of draws per step `K` are fixed per site (they define the word layout)."""

    def __init__(self, seed: int, env: int, stream: int, n: int, K: int):
        self.seed, self.env, self.stream, self.n, self.K = seed, env, stream, int(n), int(K)
        self.episode = self.step = self.k = 0

    def begin(self, episode: int, step: int) -> None:
        self.episode, self.step, self.k = episode, step, 0

    def randint(self, n: int) -> int:
        assert int(n) == self.n and self.k < self.K, "packed site: fixed n, at most K draws per step"
        d = packed_randint(self.seed, self.env, self.episode, self.step, self.stream, self.n,
                           self.K, self.k)
        self.k += 1
        return d


# Random123 known-answer vectors for philox4x32-10 (kat_vectors in the Random123
# distribution): (counter, key) -> output.
KAT = [
    ((0x00000000,) * 4, (0x00000000,) * 2,
     (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
    ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2,
     (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
    ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
     (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
]


class StepStream:
    """Sequential view of one (env, episode, step, stream): the k-th call gets idx=k.

    Used to patch the reference's global RNG call sites: within one env step the
    reference consumes its generator sequentially, so "k-th call at this site during
    this step" is well defined (for supply chain, k == customer index because customers
    draw in agent order, env.py:324 / supply_chain.py:64)."""

    def __init__(self, seed: int, env: int, stream: int):
        self.seed, self.env, self.stream = seed, env, stream
        self.episode = 0
        self.step = 0
        self.k = 0

    def begin(self, episode: int, step: int) -> None:
        self.episode, self.step, self.k = episode, step, 0

    def next_d24(self) -> int:
        d = d24(self.seed, self.env, self.episode, self.step, self.stream, self.k)
        self.k += 1
        return d

    def randint(self, n: int) -> int:
        return randint(n, self.next_d24())

    def uniform01(self) -> float:
        return uniform01(self.next_d24())

    def uniform(self, low: float, high: float) -> float:
        return uniform_f64(low, high, self.next_d24())


def uniform_f64(low: float, high: float, draw: int) -> float:
    """Contract version of np.random.uniform(low, high): low + (high - low) * (d24 / 2^24),
    every operation rounded in float64 (device: __dmul_rn / __dadd_rn, no contraction)."""
    return float(low) + (float(high) - float(low)) * (int(draw) / 16777216.0)
