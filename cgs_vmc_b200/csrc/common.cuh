// Shared device helpers: Philox4x32-10, log-cosh, warp reductions, bit tricks.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#ifndef CGSVMC_MAX_SITES
#define CGSVMC_MAX_WORDS 4      // 64-bit words per walker: n_sites <= 256
#define CGSVMC_MAX_SITES 256
#endif
#define CGSVMC_FULL_MASK 0xffffffffu

namespace cgsvmc {

// ---------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11).  Counter = (c0..c3), key = (k0, k1).
// ---------------------------------------------------------------------------
struct Philox4 { uint32_t x, y, z, w; };

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1,
                                                           uint32_t c2, uint32_t c3,
                                                           uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  const uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)M0 * c0;
    const uint64_t p1 = (uint64_t)M1 * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    const uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  return Philox4{c0, c1, c2, c3};
}

// Random stream of the sampler: one Philox block per (walker, step).
__host__ __device__ __forceinline__ Philox4 walker_step_random(uint64_t seed, uint64_t walker,
                                                                uint64_t step) {
  return philox4x32_10((uint32_t)step, (uint32_t)(step >> 32), (uint32_t)walker,
                       (uint32_t)(walker >> 32), (uint32_t)seed, (uint32_t)(seed >> 32));
}

// [0, 1) with 24 random bits -- exactly representable, never 1.0.
__host__ __device__ __forceinline__ float u32_to_unit(uint32_t r) {
  return (float)(r >> 8) * (1.0f / 16777216.0f);
}

// ---------------------------------------------------------------------------
// math
// ---------------------------------------------------------------------------
// log(cosh(x)) = |x| + ln2 * (log2(1 + 2^(-2 log2(e) |x|)) - 1): 2 MUFU + 4 FP32.
// The reference evaluates tf.log(tf.cosh(x)) (wavefunctions.py:415), which
// overflows float32 for |x| > ~89; this form is equal elsewhere and finite.
__device__ __forceinline__ float log_cosh(float x) {
  const float ax = fabsf(x);
  float e, l;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(ax * -2.885390081777927f));
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(1.0f + e));
  return fmaf(l, 0.6931471805599453f, ax - 0.6931471805599453f);
}

// tanh(x) = sign(x) * (1 - 2 / (1 + e^{2|x|})), accurate to ~2 ulp.
__device__ __forceinline__ float tanh_accurate(float x) {
  const float ax = fabsf(x);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(ax * -2.885390081777927f));  // e^{-2|x|}
  const float t = __fdividef(1.0f - e, 1.0f + e);
  return copysignf(t, x);
}

__device__ __forceinline__ float fast_exp(float x) {
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 1.4426950408889634f));
  return e;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(CGSVMC_FULL_MASK, v, o);
  return v;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(CGSVMC_FULL_MASK, v, o);
  return v;
}

// ---------------------------------------------------------------------------
// bit helpers on a walker held as NW 64-bit words replicated in every lane
// ---------------------------------------------------------------------------
// Index of the k-th (0-based) set bit of the NW-word mask `m`; all lanes of the
// warp must call with identical arguments (uses a ballot).
template <int NW>
__device__ __forceinline__ int select_kth_bit(const uint64_t (&m)[NW], int k, int lane) {
  int word = 0;
  uint64_t w = m[0];
#pragma unroll
  for (int i = 0; i < NW - 1; ++i) {
    const int c = __popcll(w);
    if (word == i && k >= c) { k -= c; word = i + 1; w = m[i + 1]; }
  }
  uint32_t half = (uint32_t)w;
  int base = word * 64;
  const int clo = __popc(half);
  if (k >= clo) { k -= clo; half = (uint32_t)(w >> 32); base += 32; }
  const uint32_t below = half & ((1u << lane) - 1u);
  const bool mine = ((half >> lane) & 1u) && (__popc(below) == k);
  const uint32_t vote = __ballot_sync(CGSVMC_FULL_MASK, mine);
  return base + __ffs(vote) - 1;
}

__device__ __forceinline__ uint64_t valid_mask_word(int n_sites, int word) {
  const int rem = n_sites - word * 64;
  if (rem >= 64) return ~0ull;
  if (rem <= 0) return 0ull;
  return (1ull << rem) - 1ull;
}

template <int NW>
__device__ __forceinline__ int get_bit(const uint64_t (&m)[NW], int site) {
  uint64_t w = m[0];
#pragma unroll
  for (int i = 1; i < NW; ++i) if ((site >> 6) == i) w = m[i];
  return (int)((w >> (site & 63)) & 1ull);
}

template <int NW>
__device__ __forceinline__ void flip_bit(uint64_t (&m)[NW], int site) {
#pragma unroll
  for (int i = 0; i < NW; ++i) if ((site >> 6) == i) m[i] ^= (1ull << (site & 63));
}

}  // namespace cgsvmc
