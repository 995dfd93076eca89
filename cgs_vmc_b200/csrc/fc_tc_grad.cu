// Weighted gradient sums S_k = sum_b w_kb d z_b / d params of the
// FullyConnectedNetwork (the two tf.gradients of training.py:545-548 and the
// SWO loss gradient of 169-175) on the 5th-generation tensor cores.  Per tile of
// 128 walkers, with h_0 = sigma, h_l = act(h_{l-1} W_l + b_l), z = h_L . w_out:
//
//   forward    D[128, 3H]  = h_{l-1}[128, K] x W_l            (K-major A, K-major B)
//   backward   D[128, 2H]  = g_l[128, H] x W_l^T              (K-major A; B = the SAME weight
//                                                              image read MN-major)
//   weights    D[K + 1, 2H] = [h_{l-1} | 1]^T x (w_k . g_l)   (MN-major A and B: the
//                                                              activation planes a layer wrote as
//                                                              its successor's K-major operand are
//                                                              read transposed; the reduction runs
//                                                              over the 128 walkers of the tile)
//
// with g_L = w_out . act'(h_L), g_{l-1} = (g_l W_l^T) . act'(h_{l-1}).  The
// constant-one feature appended to every activation plane makes the bias
// gradient row K of the same GEMM.  Values are carried as two fp16 planes
// (v = v1 + v2 / S, 22 mantissa bits; the gradient tolerance is 1e-4 of the
// largest entry), the weights keep their three-way split in the forward pass;
// accumulation is fp32 in TMEM.  A tile's weight-gradient accumulator is
// drained to the CTA's slice of `partials` (transposed through shared memory so
// the read-modify-write is coalesced); a deterministic reduction over the CTA
// slices follows (launch_reduce_partials).  Weight images are streamed layer by
// layer through one shared-memory slot (TMA bulk copy, refilled under the
// epilogue of the layer that just consumed it).
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include <cuda_fp16.h>

#include "common.cuh"
#include "internal.h"
#include "tc_common.cuh"

namespace cgsvmc {
namespace {

using namespace tc;

constexpr int kWorkers = 256;             // epilogue threads
constexpr int kThreads = kWorkers + 32;   // + the MMA-issuing warp
constexpr int kTile = 128;
constexpr int kMaxL = 8;
constexpr uint32_t kWgCol = 256;          // TMEM column of the weight-gradient accumulator
constexpr uint32_t kPlane = kTile * 16;   // bytes of one 8-feature chunk of a plane (128 rows x 16 B)

struct GradDesc {
  int N, H, L, act, NW, K1;
  const __half* wimg;      // [K1/8][3H][8] then (L - 1) x [H/8][3H][8]
  const float* consts;     // bias [L][H], w_out [H], b_out
  int64_t w_off[kMaxL + 1], b_off[kMaxL + 1];   // flat offsets of W_1..W_L, w_out and the biases
  int64_t P;
};

struct GradSmem {
  size_t wslot, consts, h0, hbuf, hbuf_each, gbuf, wgbuf, wk, accout, bars, cfg, total;
};

__host__ __device__ inline GradSmem grad_plan(const GradDesc& d, int KW) {
  GradSmem p;
  const int CH = d.H / 8;
  size_t off = 0;
  p.wslot = off; off += (size_t)3 * d.H * (d.K1 > d.H ? d.K1 : d.H) * 2;
  p.consts = off; off += ((size_t)d.L * d.H + d.H + 4) * 4;
  off = (off + 127) / 128 * 128;
  p.h0 = off; off += (size_t)(d.K1 / 8 + 1) * kPlane;
  p.hbuf_each = (size_t)2 * (CH + 1) * kPlane;
  p.hbuf = off; off += (size_t)(d.L - 1) * p.hbuf_each;
  p.gbuf = off; off += (size_t)2 * CH * kPlane;
  {   // w . g planes; also the [rows][H + 4] float staging of a drain (rows <= max(K1, H) + 1)
    const size_t planes = (size_t)2 * CH * kPlane;
    const size_t staging = (size_t)((d.K1 > d.H ? d.K1 : d.H) + 1) * (d.H + 4) * 4;
    p.wgbuf = off; off += ((planes > staging ? planes : staging) + 127) / 128 * 128;
  }
  p.wk = off; off += (size_t)KW * kTile * 4;
  p.accout = off; off += (size_t)4 * KW * (d.H + 4) * 4;     // [lane quarter][k][H + 4]: one writer per slot
  off = (off + 15) / 16 * 16;
  p.bars = off; off += 64;
  p.cfg = off; off += (size_t)kTile * d.NW * 8;
  // The weight-gradient MMAs read their MN-major A operand as M = 128 rows = 16
  // chunks from the start of an activation plane (rows past the real features
  // are junk accumulator rows nobody reads): keep those reads inside the CTA's
  // allocation for narrow layers.
  size_t last_a = p.h0;
  if (d.L > 1) last_a = p.hbuf + (size_t)(d.L - 2) * p.hbuf_each + (size_t)(CH + 1) * kPlane;
  if (last_a + 16 * (size_t)kPlane > off) off = last_a + 16 * (size_t)kPlane;
  p.total = off;
  return p;
}

__device__ __forceinline__ int word_bit(const uint64_t* words, int site) {
  return (int)((words[site >> 6] >> (site & 63)) & 1ull);
}

// d act / d x through the OUTPUT h = act(x) (cos is rejected on the host)
__device__ __noinline__ float act_grad_slow(int act, float h) {
  switch (act) {
    case CGSVMC_ACT_TANH: return 1.f - h * h;
    case CGSVMC_ACT_SIGMOID: return h * (1.f - h);
    case CGSVMC_ACT_IDENTITY: return 1.f;
    case CGSVMC_ACT_SELU: return h < 0.f ? h + 1.7580993408473766f : 1.0507009873554805f;
    case CGSVMC_ACT_EXP: return h;
    default: return 1.f + h * h;   // tan
  }
}
__device__ __noinline__ float activate_slow(int act, float x) {
  if (act == CGSVMC_ACT_SELU) return x > 0.f ? 1.0507009873554805f * x : 1.7580993408473766f * (expf(x) - 1.f);
  return tc_activate(act, x);
}

// v = h1 + h2 / S for two values at once (packed conversions)
__device__ __forceinline__ void split2_pair(float a, float b, __half2& h1, __half2& h2) {
  h1 = __floats2half2_rn(a, b);
  const float2 f1 = __half22float2(h1);
  h2 = __floats2half2_rn((a - f1.x) * kSplitScale, (b - f1.y) * kSplitScale);
}

// 8 values -> the two 16-byte plane entries of their 8-feature chunk
__device__ __forceinline__ void store_split2(char* plane1, char* plane2, const float (&v)[8]) {
  __half2 a[4], b[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) split2_pair(v[2 * k], v[2 * k + 1], a[k], b[k]);
  *reinterpret_cast<uint4*>(plane1) = make_uint4(*reinterpret_cast<uint32_t*>(&a[0]), *reinterpret_cast<uint32_t*>(&a[1]),
                                                 *reinterpret_cast<uint32_t*>(&a[2]), *reinterpret_cast<uint32_t*>(&a[3]));
  *reinterpret_cast<uint4*>(plane2) = make_uint4(*reinterpret_cast<uint32_t*>(&b[0]), *reinterpret_cast<uint32_t*>(&b[1]),
                                                 *reinterpret_cast<uint32_t*>(&b[2]), *reinterpret_cast<uint32_t*>(&b[3]));
}

// the 8 values of a chunk back from its two plane entries
__device__ __forceinline__ void load_split2(const char* plane1, const char* plane2, float (&v)[8]) {
  const uint4 a = *reinterpret_cast<const uint4*>(plane1);
  const uint4 b = *reinterpret_cast<const uint4*>(plane2);
  const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 fa = __half22float2(*reinterpret_cast<const __half2*>(&aw[k]));
    const float2 fb = __half22float2(*reinterpret_cast<const __half2*>(&bw[k]));
    v[2 * k] = fmaf(fb.x, 1.f / kSplitScale, fa.x);
    v[2 * k + 1] = fmaf(fb.y, 1.f / kSplitScale, fa.y);
  }
}

// Column sums over the 32 lanes of a warp for 8 values per lane: lane l returns
// the total of column (l >> 2) & 7 (9 shuffles instead of 40).
__device__ __forceinline__ float colsum8(const float (&v)[8], int lane) {
  float a[4], b[2];
  const bool u16 = (lane & 16) != 0, u8 = (lane & 8) != 0, u4 = (lane & 4) != 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float send = u16 ? v[i] : v[i + 4], keep = u16 ? v[i + 4] : v[i];
    a[i] = keep + __shfl_xor_sync(CGSVMC_FULL_MASK, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float send = u8 ? a[i] : a[i + 2], keep = u8 ? a[i + 2] : a[i];
    b[i] = keep + __shfl_xor_sync(CGSVMC_FULL_MASK, send, 8);
  }
  float c = (u4 ? b[1] : b[0]) + __shfl_xor_sync(CGSVMC_FULL_MASK, u4 ? b[0] : b[1], 4);
  c += __shfl_xor_sync(CGSVMC_FULL_MASK, c, 2);
  c += __shfl_xor_sync(CGSVMC_FULL_MASK, c, 1);
  return c;
}

template <int H, int KW>
__global__ void __launch_bounds__(kThreads, 1)
fc_grad_kernel(GradDesc d, const uint64_t* __restrict__ packed, const float* __restrict__ weights, int64_t B,
               int64_t per_cta, float* __restrict__ partials) {
  constexpr int CH = H / 8, HH = H / 2, HS = H + 4;
  extern __shared__ __align__(1024) char smem[];
  const GradSmem pl = grad_plan(d, KW);
  char* wslot = smem + pl.wslot;
  float* bias_s = reinterpret_cast<float*>(smem + pl.consts);
  float* wout_s = bias_s + d.L * H;
  char* h0 = smem + pl.h0;
  char* gbuf = smem + pl.gbuf;
  char* wgbuf = smem + pl.wgbuf;
  float* stage = reinterpret_cast<float*>(wgbuf);
  float* wk_s = reinterpret_cast<float*>(smem + pl.wk);
  float* accout = reinterpret_cast<float*>(smem + pl.accout);
  uint64_t* mma_bar = reinterpret_cast<uint64_t*>(smem + pl.bars);
  uint64_t* wg_bar = mma_bar + 1;
  uint64_t* wbar = mma_bar + 2;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(mma_bar + 3);
  uint64_t* cfg = reinterpret_cast<uint64_t*>(smem + pl.cfg);
  auto hbuf = [&](int l) { return smem + pl.hbuf + (size_t)(l - 1) * pl.hbuf_each; };   // h_l, l = 1 .. L-1

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool issuer = warp == kWorkers / 32;
  const int q = warp & 3, half = (warp >> 2) & 1;
  const int r = 32 * q + lane;              // tile row (epilogues) / accumulator row (drain)
  const int act = d.act;
  const bool relu = act == CGSVMC_ACT_RELU;

  // ---- one-time setup ----
  for (int e = threadIdx.x; e < d.L * H + H + 1; e += kThreads) bias_s[e] = d.consts[e];
  for (int e = threadIdx.x; e < 4 * KW * (H + 4); e += kThreads) accout[e] = 0.f;
  // constant-one feature behind the real ones: chunk K1/8 of h_0, chunk CH of the
  // first plane of h_1 .. h_{L-1} (second planes: zero)
  for (int e = threadIdx.x; e < kTile; e += kThreads) {
    const uint4 one = make_uint4(0x00003c00u, 0u, 0u, 0u), zero = make_uint4(0u, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(h0 + (size_t)(d.K1 / 8) * kPlane + e * 16) = one;
    for (int l = 1; l < d.L; ++l) {
      *reinterpret_cast<uint4*>(hbuf(l) + (size_t)CH * kPlane + e * 16) = one;
      *reinterpret_cast<uint4*>(hbuf(l) + (size_t)(CH + 1 + CH) * kPlane + e * 16) = zero;
    }
  }
  float* part = partials + (size_t)blockIdx.x * KW * d.P;
  for (int64_t e = threadIdx.x; e < (int64_t)KW * d.P; e += kThreads) part[e] = 0.f;
  if (threadIdx.x == 0) {
    mbar_init(mma_bar, 1);
    mbar_init(wg_bar, 1);
    mbar_init(wbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_holder)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_holder;
  uint32_t mma_ph = 0, wg_ph = 0, w_ph = 0;

  auto load_w = [&](int layer) {          // weight image of layer `layer` (0-based) into the slot
    if (threadIdx.x == 0) {
      const size_t off = layer == 0 ? 0 : ((size_t)d.K1 + (size_t)(layer - 1) * H) * 3 * H;
      const uint32_t bytes = (uint32_t)((layer == 0 ? d.K1 : H) * 3 * H * 2);
      bulk_load_async(wslot, d.wimg + off, bytes, wbar);
    }
  };
  auto wait_w = [&]() { mbar_wait(wbar, w_ph); w_ph ^= 1u; };
  auto commit = [&](uint64_t* bar) {
    if (elect_one())
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                   : "memory");
    __syncwarp();
  };

  // descriptors: K-major operands: LBO = stride between the two 8-element K chunks,
  // SBO = 8 units (8 rows of 16 B); MN-major operands: LBO = 8 units (8 K rows),
  // SBO = stride between 8-element M / N chunks.  Units of 16 bytes; version 1.
  const uint32_t idesc0 = (1u << 4) | (8u << 24);                  // D = F32, A = B = F16, M = 128
  const uint32_t hi_k = 8u | (1u << 14);                           // K-major: SBO = 8
  const uint32_t hi_mn_plane = (uint32_t)kTile | (1u << 14);       // MN-major over a plane: chunk stride 128 units
  const uint32_t hi_mn_w = (uint32_t)(3 * H) | (1u << 14);         // MN-major over a weight image: chunk stride 3H units
  const uint32_t w_units = smem_u32(wslot) >> 4;

  const int64_t b_begin = (int64_t)blockIdx.x * per_cta, b_end = min(B, b_begin + per_cta);
  for (int64_t b0 = b_begin; b0 < b_end; b0 += kTile) {
    const int n_items = (int)min((int64_t)kTile, b_end - b0);
    load_w(0);
    for (int e = threadIdx.x; e < kTile * d.NW; e += kThreads)
      cfg[e] = e < n_items * d.NW ? packed[b0 * d.NW + e] : 0ull;
    for (int e = threadIdx.x; e < KW * kTile; e += kThreads) {
      const int k = e / kTile, t = e - k * kTile;
      wk_s[e] = t < n_items ? weights[(int64_t)k * B + b0 + t] : 0.f;
    }
    __syncthreads();
    // input plane: spins as fp16 +-1, [K1/8 chunks][128 rows][8]
    for (int e = threadIdx.x; e < (d.K1 / 8) * kTile; e += kThreads) {
      const int c = e / kTile, row = e - c * kTile;
      uint32_t w[4] = {0u, 0u, 0u, 0u};
      if (row < n_items) {
        const uint64_t* words = cfg + (size_t)row * d.NW;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int site = 8 * c + k;
          uint32_t hbits = 0u;
          if (site < d.N) hbits = word_bit(words, site) ? 0x3c00u : 0xbc00u;
          w[k >> 1] |= hbits << (16 * (k & 1));
        }
      }
      *reinterpret_cast<uint4*>(h0 + (size_t)e * 16) = make_uint4(w[0], w[1], w[2], w[3]);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();

    // ================= forward =================
    for (int l = 0; l < d.L; ++l) {
      wait_w();
      if (issuer) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t b_lbo = (uint32_t)(3 * H) << 16, a_lbo = (uint32_t)kTile << 16;
        if (l == 0) {
          const uint32_t a_units = smem_u32(h0) >> 4;
          for (int ks = 0; ks < d.K1 / 16; ++ks) {
            const uint64_t a = ((uint64_t)hi_k << 32) | ((a_units + (uint32_t)(2 * ks * kTile)) | a_lbo);
            const uint64_t b = ((uint64_t)hi_k << 32) | ((w_units + (uint32_t)(2 * ks * 3 * H)) | b_lbo);
            if (elect_one()) mma_f16(tmem, a, b, idesc0 | ((uint32_t)(3 * H >> 3) << 17), ks > 0 ? 1u : 0u);
          }
        } else {
          const uint32_t a_units = smem_u32(hbuf(l)) >> 4;
#pragma unroll
          for (int ks = 0; ks < H / 16; ++ks) {
            const uint32_t a_lo = (a_units + (uint32_t)(2 * ks * kTile)) | a_lbo;
            const uint64_t a1 = ((uint64_t)hi_k << 32) | a_lo;
            const uint64_t a2 = ((uint64_t)hi_k << 32) | (a_lo + (uint32_t)(CH + 1) * kTile);
            const uint64_t b = ((uint64_t)hi_k << 32) | ((w_units + (uint32_t)(2 * ks * 3 * H)) | b_lbo);
            if (elect_one()) {
              mma_f16(tmem, a1, b, idesc0 | ((uint32_t)(3 * H >> 3) << 17), ks > 0 ? 1u : 0u);   // [P0 P1 P2] += A1 [b1 b2 b3]
              mma_f16(tmem + H, a2, b, idesc0 | ((uint32_t)(2 * H >> 3) << 17), 1u);             // [P1 P2]    += A2 [b1 b2]
            }
          }
        }
        commit(mma_bar);
      }
      mbar_wait(mma_bar, mma_ph);
      mma_ph ^= 1u;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (l + 1 < d.L) load_w(l + 1);          // the slot is free: next layer's weights under the epilogue
      if (!issuer) {
        const bool last = l == d.L - 1;
        const float* bj = bias_s + l * H + half * HH;
        const uint32_t trow = tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(half * HH);
        float db_acc[KW];
#pragma unroll
        for (int k = 0; k < KW; ++k) db_acc[k] = 0.f;
#pragma unroll 1
        for (int c8 = 0; c8 < HH / 8; ++c8) {
          uint32_t p0[8], p1[8], p2[8];
          tmem_ld8_nowait(trow + (uint32_t)(8 * c8), p0);
          tmem_ld8_nowait(trow + (uint32_t)(H + 8 * c8), p1);
          tmem_ld8_nowait(trow + (uint32_t)(2 * H + 8 * c8), p2);
          tmem_ld_wait();
          float v[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float pre = fmaf(fmaf(__uint_as_float(p2[k]), 1.f / kSplitScale, __uint_as_float(p1[k])),
                                   1.f / kSplitScale, __uint_as_float(p0[k])) + bj[8 * c8 + k];
            v[k] = relu ? fmaxf(pre, 0.f) : activate_slow(act, pre);
          }
          const int kc = (half * HH) / 8 + c8;
          if (!last) {
            char* hb = hbuf(l + 1);
            store_split2(hb + ((size_t)kc * kTile + r) * 16, hb + ((size_t)(CH + 1 + kc) * kTile + r) * 16, v);
          } else {
            // d z / d w_out = h_L: weighted column sums; g_L = w_out . act'(h_L)
#pragma unroll
            for (int k = 0; k < KW; ++k) {
              const float wkr = wk_s[k * kTile + r];
              float wv[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) wv[e] = wkr * v[e];
              const float cs = colsum8(wv, lane);
              // slot (lane quarter q, k, column): written by this warp only -- deterministic, no atomics
              if ((lane & 3) == 0) accout[(q * KW + k) * (H + 4) + half * HH + 8 * c8 + ((lane >> 2) & 7)] += cs;
            }
            float g[8];
#pragma unroll
            for (int e = 0; e < 8; ++e)
              g[e] = wout_s[half * HH + 8 * c8 + e] * (relu ? (v[e] > 0.f ? 1.f : 0.f) : act_grad_slow(act, v[e]));
            store_split2(gbuf + ((size_t)kc * kTile + r) * 16, gbuf + ((size_t)(CH + kc) * kTile + r) * 16, g);
          }
        }
        if (last && half == 0) {                // d z / d b_out = 1
#pragma unroll
          for (int k = 0; k < KW; ++k) {
            const float s = warp_sum(wk_s[k * kTile + r]);
            if (lane == 0) accout[(q * KW + k) * (H + 4) + H] += s;
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
    }

    // ================= backward =================
    // gbuf holds g_l = d z / d pre_l of layer image l (0-based), l = L-1 .. 0
    for (int l = d.L - 1; l >= 0; --l) {
      const int in_dim = l == 0 ? d.N : H;
      const int ones_row = l == 0 ? d.K1 : H;               // row of the constant-one feature = bias gradient
      char* a_plane = l == 0 ? h0 : hbuf(l);
      const uint32_t a_split = (uint32_t)(CH + 1) * kTile;  // second plane of h_l (units)
      if (l >= 1) {
        if (l < d.L - 1) wait_w();                           // image l streamed back in
        if (issuer) {
          // raw = g_l W_l^T: P0 = g1 b1^T, P1 = g1 b2^T + g2 b1^T  (B: the forward image, MN-major)
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t g_units = smem_u32(gbuf) >> 4;
          const uint32_t idesc = idesc0 | (1u << 16) | ((uint32_t)(H >> 3) << 17);
#pragma unroll
          for (int ks = 0; ks < H / 16; ++ks) {
            const uint32_t a_lo = (g_units + (uint32_t)(2 * ks * kTile)) | ((uint32_t)kTile << 16);
            const uint64_t a1 = ((uint64_t)hi_k << 32) | a_lo;
            const uint64_t a2 = ((uint64_t)hi_k << 32) | (a_lo + (uint32_t)CH * kTile);
            const uint64_t b1 = ((uint64_t)hi_mn_w << 32) | ((w_units + (uint32_t)(16 * ks)) | (8u << 16));
            const uint64_t b2 = ((uint64_t)hi_mn_w << 32) | ((w_units + (uint32_t)(H + 16 * ks)) | (8u << 16));
            if (elect_one()) {
              mma_f16(tmem, a1, b1, idesc, ks > 0 ? 1u : 0u);
              mma_f16(tmem + H, a1, b2, idesc, ks > 0 ? 1u : 0u);
              mma_f16(tmem + H, a2, b1, idesc, 1u);
            }
          }
          commit(mma_bar);
        }
      }
      // weight gradients of layer l for every weight column
      for (int k = 0; k < KW; ++k) {
        if (!issuer) {                                       // w_k . g_l as MN-major B operand planes
#pragma unroll 1
          for (int c8 = 0; c8 < HH / 8; ++c8) {
            const int kc = (half * HH) / 8 + c8;
            float g[8];
            load_split2(gbuf + ((size_t)kc * kTile + r) * 16, gbuf + ((size_t)(CH + kc) * kTile + r) * 16, g);
            const float wkr = wk_s[k * kTile + r];
#pragma unroll
            for (int e = 0; e < 8; ++e) g[e] *= wkr;
            store_split2(wgbuf + ((size_t)kc * kTile + r) * 16, wgbuf + ((size_t)(CH + kc) * kTile + r) * 16, g);
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (issuer) {
          // Q0 = h1^T d1, Q1 = h1^T d2 + h2^T d1   (A and B MN-major; K = the 128 tile rows)
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a_units = smem_u32(a_plane) >> 4, wg_units = smem_u32(wgbuf) >> 4;
          const uint32_t idesc_mn = idesc0 | (1u << 15) | (1u << 16);
          for (int ks = 0; ks < kTile / 16; ++ks) {
            const uint64_t a1 = ((uint64_t)hi_mn_plane << 32) | ((a_units + (uint32_t)(16 * ks)) | (8u << 16));
            const uint64_t a2 = ((uint64_t)hi_mn_plane << 32) | ((a_units + a_split + (uint32_t)(16 * ks)) | (8u << 16));
            const uint64_t bb = ((uint64_t)hi_mn_plane << 32) | ((wg_units + (uint32_t)(16 * ks)) | (8u << 16));
            if (elect_one()) {
              mma_f16(tmem + kWgCol, a1, bb, idesc_mn | ((uint32_t)(2 * H >> 3) << 17), ks > 0 ? 1u : 0u);
              if (l > 0) mma_f16(tmem + kWgCol + H, a2, bb, idesc_mn | ((uint32_t)(H >> 3) << 17), 1u);
            }
          }
          commit(wg_bar);
        }
        mbar_wait(wg_bar, wg_ph);
        wg_ph ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // drain: accumulator row r = input feature, columns = output features -> staging [row][H + 4]
        // (TMEM loads are warp-collective: every lane loads, rows past the last needed one do not store)
        if (!issuer) {
          const uint32_t trow = tmem + ((uint32_t)(32 * q) << 16) + kWgCol + (uint32_t)(half * HH);
          const bool keep = r <= ones_row;
#pragma unroll 1
          for (int c8 = 0; c8 < HH / 8; ++c8) {
            uint32_t q0[8], q1[8];
            tmem_ld8_nowait(trow + (uint32_t)(8 * c8), q0);
            tmem_ld8_nowait(trow + (uint32_t)(H + 8 * c8), q1);
            tmem_ld_wait();
            float4 lo, hi;
            lo.x = fmaf(__uint_as_float(q1[0]), 1.f / kSplitScale, __uint_as_float(q0[0]));
            lo.y = fmaf(__uint_as_float(q1[1]), 1.f / kSplitScale, __uint_as_float(q0[1]));
            lo.z = fmaf(__uint_as_float(q1[2]), 1.f / kSplitScale, __uint_as_float(q0[2]));
            lo.w = fmaf(__uint_as_float(q1[3]), 1.f / kSplitScale, __uint_as_float(q0[3]));
            hi.x = fmaf(__uint_as_float(q1[4]), 1.f / kSplitScale, __uint_as_float(q0[4]));
            hi.y = fmaf(__uint_as_float(q1[5]), 1.f / kSplitScale, __uint_as_float(q0[5]));
            hi.z = fmaf(__uint_as_float(q1[6]), 1.f / kSplitScale, __uint_as_float(q0[6]));
            hi.w = fmaf(__uint_as_float(q1[7]), 1.f / kSplitScale, __uint_as_float(q0[7]));
            if (keep) {
              float* dst = stage + (size_t)r * HS + half * HH + 8 * c8;
              *reinterpret_cast<float4*>(dst) = lo;
              *reinterpret_cast<float4*>(dst + 4) = hi;
            }
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        {   // coalesced read-modify-write of this CTA's (L2-resident) slice, eight
            // independent loads in flight per thread (a load-add-store chain per
            // element would pay the L2 latency in_dim * H / 288 times)
          float* pw = part + (size_t)k * d.P + d.w_off[l];
          const int n = in_dim * H;
          for (int e0 = threadIdx.x; e0 < n; e0 += 8 * kThreads) {
            float old[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const int e = e0 + u * kThreads;
              old[u] = e < n ? __ldcg(pw + e) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const int e = e0 + u * kThreads;
              if (e < n) {
                const int i = e / H, j = e - i * H;
                pw[e] = old[u] + stage[(size_t)i * HS + j];
              }
            }
          }
          float* pb = part + (size_t)k * d.P + d.b_off[l];
          for (int j = threadIdx.x; j < H; j += kThreads) pb[j] += stage[(size_t)ones_row * HS + j];
        }
        __syncthreads();
      }
      if (l >= 1) {
        mbar_wait(mma_bar, mma_ph);
        mma_ph ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (l - 1 >= 1) load_w(l - 1);                       // weights of the next backward layer
        if (!issuer) {                                       // g_{l-1} = raw . act'(h_{l-1})
          const uint32_t trow = tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(half * HH);
          const char* hb = hbuf(l);                          // h of the layer below = input of image l
#pragma unroll 1
          for (int c8 = 0; c8 < HH / 8; ++c8) {
            uint32_t p0[8], p1[8];
            tmem_ld8_nowait(trow + (uint32_t)(8 * c8), p0);
            tmem_ld8_nowait(trow + (uint32_t)(H + 8 * c8), p1);
            tmem_ld_wait();
            const int kc = (half * HH) / 8 + c8;
            float hv[8], g[8];
            load_split2(hb + ((size_t)kc * kTile + r) * 16, hb + ((size_t)(CH + 1 + kc) * kTile + r) * 16, hv);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float raw = fmaf(__uint_as_float(p1[e]), 1.f / kSplitScale, __uint_as_float(p0[e]));
              g[e] = raw * (relu ? (hv[e] > 0.f ? 1.f : 0.f) : act_grad_slow(act, hv[e]));
            }
            store_split2(gbuf + ((size_t)kc * kTile + r) * 16, gbuf + ((size_t)(CH + kc) * kTile + r) * 16, g);
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
      }
    }
  }

  // output layer: d z / d w_out, d z / d b_out
  __syncthreads();
  for (int e = threadIdx.x; e < KW * (H + 1); e += kThreads) {
    const int k = e / (H + 1), j = e - k * (H + 1);
    float total = 0.f;
#pragma unroll
    for (int qq = 0; qq < 4; ++qq) total += accout[(qq * KW + k) * (H + 4) + j];
    part[(size_t)k * d.P + (j < H ? d.w_off[d.L] + j : d.b_off[d.L])] += total;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

bool grad_enabled() {
  const char* e = getenv("CGSVMC_FC_TC_GRAD");
  if (e != nullptr && atoi(e) == 0) return false;
  const char* f = getenv("CGSVMC_FC_TC");
  return f == nullptr || atoi(f) != 0;
}

bool make_grad_desc(const cgsvmc_ansatz* a, GradDesc* out) {
  const cgsvmc_ansatz_desc& s = a->desc;
  if (s.kind != CGSVMC_ANSATZ_FULLY_CONNECTED) return false;
  const int H = s.layer_size;
  if (s.num_layers < 1 || s.num_layers > kMaxL) return false;
  if (H != 16 && H != 32 && H != 48 && H != 64 && H != 80) return false;
  if (s.nonlinearity == CGSVMC_ACT_COS) return false;          // no gradient through the output value
  GradDesc d;
  memset(&d, 0, sizeof(d));
  d.N = s.n_sites; d.H = H; d.L = s.num_layers; d.act = s.nonlinearity;
  d.NW = n_words(s.n_sites);
  d.K1 = (s.n_sites + 15) / 16 * 16;
  if (d.K1 + 8 > 128) return false;                            // input features + the one row must fit M = 128
  d.P = a->n_params;
  for (int l = 0; l <= d.L; ++l) { d.w_off[l] = a->offsets[2 * l]; d.b_off[l] = a->offsets[2 * l + 1]; }
  if (grad_plan(d, 2).total + 1024 > (size_t)a->max_smem_optin) return false;
  *out = d;
  return true;
}

template <typename F>
int opt_in(F kernel, size_t bytes) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(smem)");
  return CGSVMC_OK;
}

template <int KW>
int launch_grad(const GradDesc& d, int grid, size_t smem, cudaStream_t st, const uint64_t* packed,
                const float* w, int64_t B, int64_t per_cta, float* partials) {
#define CGSVMC_FCG(HV)                                                                    \
  case HV:                                                                                \
    if (int rc = opt_in(fc_grad_kernel<HV, KW>, smem)) return rc;                         \
    fc_grad_kernel<HV, KW><<<grid, kThreads, smem, st>>>(d, packed, w, B, per_cta, partials); \
    break;
  switch (d.H) {
    CGSVMC_FCG(16) CGSVMC_FCG(32) CGSVMC_FCG(48) CGSVMC_FCG(64)
    default:
      if (int rc = opt_in(fc_grad_kernel<80, KW>, smem)) return rc;
      fc_grad_kernel<80, KW><<<grid, kThreads, smem, st>>>(d, packed, w, B, per_cta, partials);
      break;
  }
#undef CGSVMC_FCG
  return cuda_fail(cudaGetLastError(), "fc_tc grad launch");
}

}  // namespace

bool fc_tc_grad_supported(const cgsvmc_ansatz* a) {
  if (!grad_enabled()) return false;
  GradDesc d;
  return make_grad_desc(a, &d);
}

int fc_tc_grad(cgsvmc_ansatz* a, const uint64_t* packed, const float* weights, int64_t B, int K,
               float* out, cudaStream_t st) {
  GradDesc d;
  if (!make_grad_desc(a, &d)) { set_error("fc_tc grad: unsupported network"); return CGSVMC_ERR_UNSUPPORTED; }
  const void* wimg = nullptr;
  const float* consts = nullptr;
  if (int rc = fc_tc_image(a, &wimg, &consts, st)) return rc;
  d.wimg = reinterpret_cast<const __half*>(wimg);
  d.consts = consts;
  const int64_t tiles = (B + kTile - 1) / kTile;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(tiles, a->num_sms));
  const int64_t per_cta = ((tiles + grid - 1) / grid) * kTile;
  const int64_t P = d.P;
  if (int rc = ensure_scratch(a, (size_t)grid * 2 * P * sizeof(float))) return rc;
  float* partials = a->scratch;
  for (int k0 = 0; k0 < K; k0 += 2) {
    const int kk = std::min(2, K - k0);
    const float* w = weights + (int64_t)k0 * B;
    const size_t smem = grad_plan(d, kk).total;
    const int used = (int)((B + per_cta - 1) / per_cta);
    if (kk == 2) { if (int rc = launch_grad<2>(d, used, smem, st, packed, w, B, per_cta, partials)) return rc; }
    else { if (int rc = launch_grad<1>(d, used, smem, st, packed, w, B, per_cta, partials)) return rc; }
    if (int rc = launch_reduce_partials(partials, used, (int64_t)kk * P, out + (int64_t)k0 * P, st)) return rc;
  }
  return CGSVMC_OK;
}

}  // namespace cgsvmc
