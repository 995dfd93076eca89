// Layout conversion, initial walker state, bond enumeration and small
// reductions.  All of these are HBM-bound integer / byte work: coalesced,
// one pass, no staging.
#include "common.cuh"
#include "internal.h"

namespace cgsvmc {
namespace {

constexpr int kThreads = 256;

// float32 [B, N] of +-1  ->  uint64 [B, W].  One warp per (walker, word):
// lanes read 32 consecutive floats, a ballot forms half a word.
__global__ void pack_kernel(const float* __restrict__ configs, int64_t B, int N, int W,
                            uint64_t* __restrict__ packed) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t item = warp; item < B * W; item += n_warps) {
    const int64_t b = item / W;
    const int w = (int)(item - b * W);
    const int i0 = w * 64 + lane, i1 = i0 + 32;
    const float v0 = i0 < N ? configs[b * N + i0] : -1.f;
    const float v1 = i1 < N ? configs[b * N + i1] : -1.f;
    const uint32_t lo = __ballot_sync(CGSVMC_FULL_MASK, v0 > 0.f);
    const uint32_t hi = __ballot_sync(CGSVMC_FULL_MASK, v1 > 0.f);
    if (lane == 0) packed[item] = ((uint64_t)hi << 32) | lo;
  }
}

__global__ void unpack_kernel(const uint64_t* __restrict__ packed, int64_t B, int N, int W,
                              float* __restrict__ configs) {
  const int64_t total = B * N;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = e / N;
    const int i = (int)(e - b * N);
    const uint64_t word = packed[b * W + (i >> 6)];
    configs[e] = ((word >> (i & 63)) & 1ull) ? 1.f : -1.f;
  }
}

// utils.random_configurations (utils.py:169-192): N/2 distinct uniformly
// chosen sites are set to -1.  Selection sampling (Knuth 3.4.2 S) draws the
// subset in one ordered pass; Philox counter = (draw index, walker).
__global__ void random_configs_kernel(uint64_t* __restrict__ packed, int64_t B, int N, int W,
                                      uint64_t seed, uint64_t walker0) {
  for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < B;
       b += (int64_t)gridDim.x * blockDim.x) {
    uint64_t words[CGSVMC_MAX_WORDS];
#pragma unroll
    for (int w = 0; w < CGSVMC_MAX_WORDS; ++w) words[w] = valid_mask_word(N, w);
    int need = N / 2;
    Philox4 r = {0, 0, 0, 0};
    for (int i = 0; i < N && need > 0; ++i) {
      if ((i & 3) == 0)
        r = walker_step_random(seed ^ 0x5DEECE66Dull, walker0 + (uint64_t)b, (uint64_t)(i >> 2));
      const uint32_t x = (i & 3) == 0 ? r.x : (i & 3) == 1 ? r.y : (i & 3) == 2 ? r.z : r.w;
      if ((int)__umulhi(x, (uint32_t)(N - i)) < need) {
#pragma unroll
        for (int w = 0; w < CGSVMC_MAX_WORDS; ++w)
          if ((i >> 6) == w) words[w] &= ~(1ull << (i & 63));
        --need;
      }
    }
    for (int w = 0; w < W; ++w) packed[b * W + w] = words[w];
  }
}

// operators.py:154-167 on bits: one thread per (walker, bond).
__global__ void flip_enum_kernel(const int2* __restrict__ ij, int n_bonds,
                                 const uint64_t* __restrict__ packed, int64_t B, int W,
                                 uint64_t* __restrict__ flipped, uint32_t* __restrict__ mask) {
  const int mask_words = (n_bonds + 31) / 32;
  const int lane = threadIdx.x & 31;
  // items are (walker, bond-group-of-32) so that a ballot builds one mask word
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t item = warp; item < B * mask_words; item += n_warps) {
    const int64_t b = item / mask_words;
    const int g = (int)(item - b * mask_words);
    const int k = g * 32 + lane;
    bool anti = false;
    if (k < n_bonds) {
      const int2 bd = ij[k];
      const uint64_t wi = packed[b * W + (bd.x >> 6)], wj = packed[b * W + (bd.y >> 6)];
      anti = (((wi >> (bd.x & 63)) ^ (wj >> (bd.y & 63))) & 1ull) != 0;
      if (flipped != nullptr) {
        for (int w = 0; w < W; ++w) {
          uint64_t word = packed[b * W + w];
          if (anti && (bd.x >> 6) == w) word ^= 1ull << (bd.x & 63);
          if (anti && (bd.y >> 6) == w) word ^= 1ull << (bd.y & 63);
          flipped[(b * n_bonds + k) * W + w] = word;
        }
      }
    }
    const uint32_t vote = __ballot_sync(CGSVMC_FULL_MASK, anti);
    if (lane == 0 && mask != nullptr) mask[item] = vote;
  }
}

__global__ void energy_stats_kernel(const float* __restrict__ e, int64_t B, double* stats) {
  double s1 = 0.0, s2 = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < B;
       i += (int64_t)gridDim.x * blockDim.x) {
    const double v = (double)e[i];
    s1 += v;
    s2 += v * v;
  }
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  __shared__ double sh[2][kThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { sh[0][warp] = s1; sh[1][warp] = s2; }
  __syncthreads();
  if (warp == 0) {
    s1 = lane < kThreads / 32 ? sh[0][lane] : 0.0;
    s2 = lane < kThreads / 32 ? sh[1][lane] : 0.0;
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if (lane == 0) {
      atomicAdd(&stats[0], s1);
      atomicAdd(&stats[1], s2);
      if (blockIdx.x == 0) atomicAdd(&stats[2], (double)B);
    }
  }
}

// out[n] += sum_p partials[p][n]   (fixed order => deterministic).  A block
// reduces 32 consecutive outputs; its 8 warps each take every 8th partial
// (coalesced 128-byte reads), then the 8 slices are added in a fixed order.
__global__ void reduce_partials_kernel(const float* __restrict__ partials, int n_parts, int64_t n,
                                       float* __restrict__ out) {
  __shared__ float sh[8][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t i0 = (int64_t)blockIdx.x * 32; i0 < n; i0 += (int64_t)gridDim.x * 32) {
    const int64_t i = i0 + lane;
    float acc = 0.f;
    if (i < n)
      for (int p = warp; p < n_parts; p += 8) acc += partials[(int64_t)p * n + i];
    sh[warp][lane] = acc;
    __syncthreads();
    if (warp == 0 && i < n) {
      float total = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) total += sh[w][lane];
      out[i] += total;
    }
    __syncthreads();
  }
}

int blocks_for(int64_t work_items, int per_block) {
  const int64_t need = (work_items + per_block - 1) / per_block;
  return (int)std::max<int64_t>(1, std::min<int64_t>(need, 148 * 16));
}

}  // namespace

int launch_pack(const float* configs, int64_t B, int N, uint64_t* packed, cudaStream_t s) {
  if (B == 0) return CGSVMC_OK;
  const int W = n_words(N);
  pack_kernel<<<blocks_for(B * W, kThreads / 32), kThreads, 0, s>>>(configs, B, N, W, packed);
  return cuda_fail(cudaGetLastError(), "pack launch");
}

int launch_unpack(const uint64_t* packed, int64_t B, int N, float* configs, cudaStream_t s) {
  if (B == 0) return CGSVMC_OK;
  unpack_kernel<<<blocks_for(B * N, kThreads), kThreads, 0, s>>>(packed, B, N, n_words(N), configs);
  return cuda_fail(cudaGetLastError(), "unpack launch");
}

int launch_random_configs(uint64_t* packed, int64_t B, int N, uint64_t seed, uint64_t walker0,
                          cudaStream_t s) {
  if (B == 0) return CGSVMC_OK;
  random_configs_kernel<<<blocks_for(B, kThreads), kThreads, 0, s>>>(packed, B, N, n_words(N), seed,
                                                                     walker0);
  return cuda_fail(cudaGetLastError(), "random_configs launch");
}

int launch_flip_enum(const cgsvmc_ham* h, const uint64_t* packed, int64_t B, uint64_t* flipped,
                     uint32_t* mask, cudaStream_t s) {
  if (B == 0 || h->n_bonds == 0) return CGSVMC_OK;
  const int mask_words = (h->n_bonds + 31) / 32;
  flip_enum_kernel<<<blocks_for(B * mask_words, kThreads / 32), kThreads, 0, s>>>(
      h->ij, h->n_bonds, packed, B, n_words(h->n_sites), flipped, mask);
  return cuda_fail(cudaGetLastError(), "flip_enum launch");
}

__global__ void fill_kernel(float* __restrict__ dst, int64_t n, float value) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n;
       e += (int64_t)gridDim.x * blockDim.x) dst[e] = value;
}

__global__ void advance_counter_kernel(uint64_t* counter, uint64_t by) { *counter += by; }

int launch_advance_counter(uint64_t* counter, uint64_t by, cudaStream_t s) {
  advance_counter_kernel<<<1, 1, 0, s>>>(counter, by);
  return cuda_fail(cudaGetLastError(), "advance_counter launch");
}

int launch_fill(float* dst, int64_t n, float value, cudaStream_t s) {
  if (n <= 0) return CGSVMC_OK;
  const int blocks = (int)std::min<int64_t>((n + 255) / 256, 1024);
  fill_kernel<<<blocks, 256, 0, s>>>(dst, n, value);
  return cuda_fail(cudaGetLastError(), "fill launch");
}

int launch_energy_stats(const float* e, int64_t B, double* stats, cudaStream_t s) {
  if (B == 0) return CGSVMC_OK;
  energy_stats_kernel<<<blocks_for(B, kThreads * 8), kThreads, 0, s>>>(e, B, stats);
  return cuda_fail(cudaGetLastError(), "energy_stats launch");
}

int launch_reduce_partials(const float* partials, int n_parts, int64_t n, float* out,
                           cudaStream_t s) {
  reduce_partials_kernel<<<blocks_for(n, 32), kThreads, 0, s>>>(partials, n_parts, n, out);
  return cuda_fail(cudaGetLastError(), "reduce_partials launch");
}

}  // namespace cgsvmc
