// phx_sc_wire.cpp -- host expansion of the supply chain's compact wire words (phx_sc_wire.h).
//
// Every quotient is the correctly rounded float32 of n / d, which is what the kernel's sc_ratio
// produces and what the reference computes (float64 division cast to float32; the float64
// rounding can never move a float32 rounding decision here).  IEEE float32
// division of two exactly representable integers IS that value, so the vector path is four
// `vdivps`; it streams its results with non-temporal stores (the planes are written once and
// read by somebody else).
#include "phx_sc_wire.h"

#include <immintrin.h>

#include "phx_hostpool.h"

namespace phx {

void sc_wire_expand_scalar(const ScWireParams& p, const uint32_t* wire, size_t n, float* obs,
                           float* reward, uint8_t* all_done) {
  const float ms = (float)p.max_stock, cap = (float)p.cap;
  for (size_t i = 0; i < n; ++i) {
    const uint32_t w = wire[i];
    const int stock = (int)(w & 0xFFFFu) - SCW_STOCK_BIAS;
    const int sales = (int)((w >> 16) & 127u), missed = (int)((w >> 23) & 127u);
    const bool reset = (w >> 31) != 0;
    obs[3 * i + 0] = (float)(reset ? 0 : stock) / ms;
    obs[3 * i + 1] = (float)sales / cap;
    obs[3 * i + 2] = (float)missed / cap;
    reward[i] = (float)(10 * sales - stock) / 10.0f;
    all_done[2 * i + 0] = 0;
    all_done[2 * i + 1] = (uint8_t)((w >> 30) & 1u);
  }
}

namespace {

// 16 env-steps per iteration; obs / reward / all_done destinations 32-byte aligned.
__attribute__((target("avx2"))) void expand_avx2(const ScWireParams& p, const uint32_t* wire,
                                                 size_t n, float* obs, float* reward,
                                                 uint8_t* all_done) {
  const __m256 ms = _mm256_set1_ps((float)p.max_stock), cap = _mm256_set1_ps((float)p.cap);
  const __m256 ten = _mm256_set1_ps(10.0f);
  const __m256i m16 = _mm256_set1_epi32(0xFFFF), m7 = _mm256_set1_epi32(127);
  const __m256i bias = _mm256_set1_epi32(SCW_STOCK_BIAS), ten_i = _mm256_set1_epi32(10);
  // AoS interleave of three 8-float vectors a, b, c -> a0 b0 c0 a1 b1 c1 ...
  const __m256i ia0 = _mm256_setr_epi32(0, 0, 0, 1, 0, 0, 2, 0);  // lanes that take a / b / c
  const __m256i ib0 = _mm256_setr_epi32(0, 0, 0, 0, 1, 0, 0, 2);
  const __m256i ic0 = _mm256_setr_epi32(0, 0, 0, 0, 0, 1, 0, 0);
  const __m256i ia1 = _mm256_setr_epi32(0, 3, 0, 0, 4, 0, 0, 5);
  const __m256i ib1 = _mm256_setr_epi32(2, 0, 3, 0, 0, 4, 0, 0);
  const __m256i ic1 = _mm256_setr_epi32(2, 0, 0, 3, 0, 0, 4, 0);
  const __m256i ia2 = _mm256_setr_epi32(0, 0, 6, 0, 0, 7, 0, 0);
  const __m256i ib2 = _mm256_setr_epi32(5, 0, 0, 6, 0, 0, 7, 0);
  const __m256i ic2 = _mm256_setr_epi32(0, 5, 0, 0, 6, 0, 0, 7);
  for (size_t i = 0; i + 16 <= n; i += 16) {
    __m256i done16[2];
    for (int hft = 0; hft < 2; ++hft) {
      const size_t j = i + 8 * hft;
      const __m256i w = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(wire + j));
      const __m256i stock = _mm256_sub_epi32(_mm256_and_si256(w, m16), bias);
      const __m256i sales = _mm256_and_si256(_mm256_srli_epi32(w, 16), m7);
      const __m256i missed = _mm256_and_si256(_mm256_srli_epi32(w, 23), m7);
      const __m256i reset = _mm256_srai_epi32(w, 31);  // all ones where the env was reset
      const __m256i k = _mm256_sub_epi32(_mm256_mullo_epi32(sales, ten_i), stock);
      const __m256 o0 = _mm256_div_ps(_mm256_cvtepi32_ps(_mm256_andnot_si256(reset, stock)), ms);
      const __m256 o1 = _mm256_div_ps(_mm256_cvtepi32_ps(sales), cap);
      const __m256 o2 = _mm256_div_ps(_mm256_cvtepi32_ps(missed), cap);
      const __m256 rw = _mm256_div_ps(_mm256_cvtepi32_ps(k), ten);
      // three interleaved output vectors
      const __m256 v0 = _mm256_blend_ps(
          _mm256_blend_ps(_mm256_permutevar8x32_ps(o0, ia0), _mm256_permutevar8x32_ps(o1, ib0), 0x92),
          _mm256_permutevar8x32_ps(o2, ic0), 0x24);
      const __m256 v1 = _mm256_blend_ps(
          _mm256_blend_ps(_mm256_permutevar8x32_ps(o0, ia1), _mm256_permutevar8x32_ps(o1, ib1), 0x24),
          _mm256_permutevar8x32_ps(o2, ic1), 0x49);
      const __m256 v2 = _mm256_blend_ps(
          _mm256_blend_ps(_mm256_permutevar8x32_ps(o0, ia2), _mm256_permutevar8x32_ps(o1, ib2), 0x49),
          _mm256_permutevar8x32_ps(o2, ic2), 0x92);
      float* o = obs + 3 * j;
      _mm256_stream_ps(o, v0);
      _mm256_stream_ps(o + 8, v1);
      _mm256_stream_ps(o + 16, v2);
      _mm256_stream_ps(reward + j, rw);
      // all_done pair {0, truncated} as one uint16 = truncated << 8
      done16[hft] = _mm256_slli_epi32(_mm256_and_si256(_mm256_srli_epi32(w, 30), _mm256_set1_epi32(1)), 8);
    }
    // 2 x 8 int32 -> 16 x uint16 (packus works per 128-bit lane: fix the order afterwards)
    const __m256i pk = _mm256_permute4x64_epi64(_mm256_packus_epi32(done16[0], done16[1]), 0xD8);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(all_done + 2 * i), pk);
  }
  _mm_sfence();
}

}  // namespace

void sc_wire_expand(HostPool& pool, const ScWireParams& p, const uint32_t* wire, size_t n,
                    float* obs, float* reward, uint8_t* all_done) {
  const bool aligned = ((uintptr_t)obs % 32 == 0) && ((uintptr_t)reward % 32 == 0) &&
                       ((uintptr_t)all_done % 32 == 0);
  const bool avx2 = aligned && __builtin_cpu_supports("avx2");
  pool.parallel_for(n, 16, [&](size_t b, size_t e) {
    size_t done = 0;
    if (avx2) {
      done = (e - b) / 16 * 16;
      expand_avx2(p, wire + b, done, obs + 3 * b, reward + b, all_done + 2 * b);
    }
    sc_wire_expand_scalar(p, wire + b + done, e - b - done, obs + 3 * (b + done),
                          reward + b + done, all_done + 2 * (b + done));
  });
}

}  // namespace phx
