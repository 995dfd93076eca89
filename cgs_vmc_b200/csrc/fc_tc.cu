// FullyConnectedNetwork (wavefunctions.py:328-388) on the 5th-generation tensor
// cores: z = w_out . act(W_L^T ... act(W_1^T sigma + b_1) ... + b_L) + b_out for
// tiles of 128 configurations, every layer one tcgen05 GEMM
//
//     D[128, 3H] (TMEM, fp32) = A[128, K] (shared, fp16 planes) x B[K, 3H] (shared)
//
// with float32-grade products from fp16 MMAs: a value v is split as
// v = v1 + v2 / S + v3 / S^2 (fp16 each, S = 2^11, 33 mantissa bits) and the
// three weight parts are concatenated along N, so per 16 input features three
// MMAs cover the six significant products while each activation plane is read
// from shared memory once:
//     A1 x [b1 b2 b3] -> columns [P0 P1 P2],  A2 x [b1 b2] -> [P1 P2],  A3 x [b1] -> [P2]
//     h = P0 + P1 / S + P2 / S^2.
// The input layer needs one plane (spins are +-1, exact in fp16).  The epilogue
// (tcgen05.ld -> bias -> nonlinearity -> split -> K-major operand planes of the
// next layer) is run by all eight warps; the last hidden layer's epilogue takes
// the dot product with w_out instead (the [H, 1] output layer never becomes a
// GEMM).  All weight images stay resident in shared memory for the lifetime of
// the CTA (one TMA bulk copy); two tiles are in flight per CTA when they fit
// (two TMEM accumulators, two activation buffers), so the MMAs of one tile
// overlap the epilogue of the other.  Warp specialisation: a ninth warp only
// issues MMAs -- it waits on a "planes ready" mbarrier that the 256 epilogue
// threads arrive on when a tile's next operand planes are written and its
// accumulator is drained -- so issuing never sits on the epilogue warps'
// critical path (measured before: 1.3 k cycles of issue per 2.6 k of epilogue).
//
// Used by log_amp, the sampler (one forward per proposal) and the local energy
// (tiles of (walker, antiparallel bond) items); the gradient stays on the SIMT
// tile path of net.cu.  CGSVMC_FC_TC=0 routes everything to net.cu.
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

// Development aid (-DCGSVMC_RBM2_TIMING build, profiles/run_fc_tc_phases.py):
// thread 0 of every CTA accumulates the cycles it waits for the MMAs of a layer,
// spends in the epilogue, and in whole forward passes.
#ifdef CGSVMC_RBM2_TIMING
__device__ unsigned long long g_fc_phase[6];   // wait cycles, epilogue cycles, epilogues, forward cycles, forwards, issue cycles
#define FC_CLOCK(NAME) const long long NAME = clock64()
#define FC_ADD(IDX, VAL) do { if (threadIdx.x == 0) atomicAdd(&g_fc_phase[IDX], (unsigned long long)(VAL)); } while (0)
#else
#define FC_CLOCK(NAME) do {} while (0)
#define FC_ADD(IDX, VAL) do {} while (0)
#endif

constexpr int kWorkers = 256;        // epilogue threads: 8 warps = 4 TMEM lane quarters x 2 column halves
constexpr int kThreads = kWorkers + 32;   // + one warp that only issues MMAs (runs ahead of the epilogues)
constexpr int kWarps = kThreads / 32;
constexpr int kTile = 128;          // configurations per tile (UMMA M)
constexpr int kMaxItems = 1024;     // local-energy items per walker chunk

struct FcDesc {
  int N, H, L, act, NW;
  int K1;                  // input features padded to a multiple of 16
  int n_hh;                // hidden -> hidden layers (L - 1)
  int nbuf;                // tiles in flight per CTA (1 or 2)
  int tmem_cols;
  int act_bytes;           // one activation buffer
  const __half* wimg;      // [K1/8][3H][8] then n_hh x [H/8][3H][8]
  const float* consts;     // bias [L][H], w_out [H], b_out
};

struct FcSmem {
  size_t w, consts, act, zpart, z, bars, total;
};

__host__ __device__ inline size_t w_bytes(const FcDesc& d) {
  return ((size_t)d.K1 + (size_t)d.n_hh * d.H) * 3 * d.H * 2;
}

__host__ __device__ inline FcSmem fc_plan(const FcDesc& d) {
  FcSmem p;
  size_t off = 0;
  p.w = off; off += w_bytes(d);
  p.consts = off; off += ((size_t)d.L * d.H + d.H + 4) * 4;
  off = (off + 127) / 128 * 128;
  p.act = off; off += (size_t)d.nbuf * d.act_bytes;
  p.zpart = off; off += (size_t)d.nbuf * 2 * kTile * 4;
  p.z = off; off += (size_t)d.nbuf * kTile * 4;
  p.bars = off; off += 64;
  p.total = off;
  return p;
}

__device__ __forceinline__ int word_bit(const uint64_t* words, int site) {
  return (int)((words[site >> 6] >> (site & 63)) & 1ull);
}

// The nonlinearities other than relu as ONE out-of-line function: inlined into
// the unrolled epilogue they multiplied its code size by the number of
// activation kinds times the elements per thread (336 KB of SASS for H = 80),
// and the relu path crawled through instruction-cache misses (measured: 10.6 k
// cycles per epilogue).
__device__ __noinline__ float activate_slow(int act, float x) { return tc_activate(act, x); }

// ---------------------------------------------------------------------------
// The forward engine.  All 256 threads of the CTA call every method.
// ---------------------------------------------------------------------------
template <int H>
struct FcEngine {
  static constexpr int HH = H / 2;          // features per epilogue warp
  static constexpr int CH = H / 8;          // 8-feature chunks per split plane
  const FcDesc& d;
  __device__ explicit FcEngine(const FcDesc& desc) : d(desc) {}
  char* wimg_s;
  float *bias_s, *wout_s;
  char* act_s;                 // [nbuf] activation buffers; the input plane aliases the start of each
  float *zpart_s, *z_s;
  uint64_t *mma_bar, *wbar;    // mma_bar[2] (MMAs of a tile-layer complete), wbar (weights landed)
  uint64_t* ready_bar;         // ready_bar[2]: a tile's operand planes written + accumulator drained (kWorkers arrivals)
  uint32_t* tmem_holder;
  uint32_t tmem;
  uint32_t phase[2];           // workers: mma_bar phases; issuer: ready_bar phases

  __device__ void setup(char* smem) {
    const FcSmem p = fc_plan(d);
    wimg_s = smem + p.w;
    bias_s = reinterpret_cast<float*>(smem + p.consts);
    wout_s = bias_s + d.L * H;
    act_s = smem + p.act;
    zpart_s = reinterpret_cast<float*>(smem + p.zpart);
    z_s = reinterpret_cast<float*>(smem + p.z);
    mma_bar = reinterpret_cast<uint64_t*>(smem + p.bars);
    wbar = mma_bar + 2;
    ready_bar = mma_bar + 3;
    tmem_holder = reinterpret_cast<uint32_t*>(mma_bar + 5);
    phase[0] = phase[1] = 0;
    for (int e = threadIdx.x; e < d.L * H + H + 1; e += kThreads) bias_s[e] = d.consts[e];
    if (threadIdx.x == 0) {
      mbar_init(mma_bar, 1);
      mbar_init(mma_bar + 1, 1);
      mbar_init(wbar, 1);
      mbar_init(ready_bar, kWorkers);
      mbar_init(ready_bar + 1, kWorkers);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                       smem_u32(tmem_holder)),
                   "r"((uint32_t)d.tmem_cols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    tmem = *tmem_holder;
    if (threadIdx.x == 0) bulk_load_async(wimg_s, d.wimg, (uint32_t)w_bytes(d), wbar);
    mbar_wait(wbar, 0);          // every thread: the weight images have landed
  }

  __device__ void teardown() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32)
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem),
                   "r"((uint32_t)d.tmem_cols)
                   : "memory");
  }

  __device__ __forceinline__ char* act_buf(int buf) const { return act_s + (size_t)buf * d.act_bytes; }
  __device__ __forceinline__ float* z_of(int buf) const { return z_s + buf * kTile; }

  // Input plane of tile `buf`: row r holds the spins of configuration r as fp16
  // +-1, K-major planes [K1/8 chunks][128 rows][8]; sites beyond N and rows
  // beyond n_items are zero.  cfg: [n_items][NW] packed words in shared memory.
  __device__ void write_input(int buf, const uint64_t* cfg, int n_items) {
    uint4* plane = reinterpret_cast<uint4*>(act_buf(buf));
    const int chunks = d.K1 / 8;
    for (int e = threadIdx.x; e < chunks * kTile; e += kThreads) {
      const int c = e / kTile, r = e - c * kTile;
      uint32_t w[4] = {0u, 0u, 0u, 0u};
      if (r < n_items) {
        const uint64_t* words = cfg + (size_t)r * d.NW;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int site = 8 * c + k;
          uint32_t hbits = 0u;
          if (site < d.N) hbits = word_bit(words, site) ? 0x3c00u : 0xbc00u;   // +1 / -1
          w[k >> 1] |= hbits << (16 * (k & 1));
        }
      }
      plane[e] = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }

  // MMAs of network layer `layer` (0 = input layer) on tile `buf`, issued by one
  // elected lane of warp 0; completion arrives on mma_bar[buf].
  __device__ void issue(int buf, int layer) {
    FC_CLOCK(t_i0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // instruction descriptor: D = F32, A = B = F16, K-major, M = 128; N per MMA
    const uint32_t idesc0 = (1u << 4) | (8u << 24);
    const uint32_t idesc3 = idesc0 | ((uint32_t)(3 * H >> 3) << 17);
    const uint32_t idesc2 = idesc0 | ((uint32_t)(2 * H >> 3) << 17);
    const uint32_t idesc1 = idesc0 | ((uint32_t)(H >> 3) << 17);
    const uint32_t desc_hi = 8u | (1u << 14);              // SBO = 8 units (8 rows of 16 B), version 1
    const uint32_t a_units = smem_u32(act_buf(buf)) >> 4;
    const uint32_t a_lbo = (uint32_t)kTile << 16;          // chunk stride: 128 rows of 16 B
    const uint32_t b_lbo = (uint32_t)(3 * H) << 16;        // chunk stride of the weight image
    const uint32_t d_tmem = tmem + (uint32_t)(buf * 3 * H);
    if (layer == 0) {
      const uint32_t b_units = smem_u32(wimg_s) >> 4;
      for (int ks = 0; ks < d.K1 / 16; ++ks) {
        const uint64_t a = ((uint64_t)desc_hi << 32) | ((a_units + (uint32_t)(2 * ks * kTile)) | a_lbo);
        const uint64_t b = ((uint64_t)desc_hi << 32) | ((b_units + (uint32_t)(2 * ks * 3 * H)) | b_lbo);
        if (elect_one()) mma_f16(d_tmem, a, b, idesc3, ks > 0 ? 1u : 0u);
      }
    } else {
      const uint32_t b_units =
          (smem_u32(wimg_s) >> 4) + (uint32_t)(d.K1 / 8) * 3 * H + (uint32_t)(layer - 1) * CH * 3 * H;
      const uint32_t split_units = (uint32_t)CH * kTile;
#pragma unroll
      for (int ks = 0; ks < H / 16; ++ks) {
        const uint32_t a_lo = (a_units + (uint32_t)(2 * ks * kTile)) | a_lbo;
        const uint64_t a1 = ((uint64_t)desc_hi << 32) | a_lo;
        const uint64_t a2 = ((uint64_t)desc_hi << 32) | (a_lo + split_units);
        const uint64_t a3 = ((uint64_t)desc_hi << 32) | (a_lo + 2 * split_units);
        const uint64_t b = ((uint64_t)desc_hi << 32) | ((b_units + (uint32_t)(2 * ks * 3 * H)) | b_lbo);
        if (elect_one()) {
          mma_f16(d_tmem, a1, b, idesc3, ks > 0 ? 1u : 0u);   // [P0 P1 P2] += A1 [b1 b2 b3]
          mma_f16(d_tmem + H, a2, b, idesc2, 1u);             // [P1 P2]    += A2 [b1 b2]
          mma_f16(d_tmem + 2 * H, a3, b, idesc1, 1u);         // [P2]       += A3 [b1]
        }
      }
    }
    if (elect_one())
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                       smem_u32(mma_bar + buf))
                   : "memory");
    __syncwarp();
#ifdef CGSVMC_RBM2_TIMING
    if (threadIdx.x == kWorkers) atomicAdd(&g_fc_phase[5], (unsigned long long)(clock64() - t_i0));
#endif
  }

  // Epilogue of layer `layer` on tile `buf`: TMEM -> bias -> nonlinearity ->
  // split planes of the next layer, or (last hidden layer) the dot product with
  // w_out.  Ends with a CTA barrier: afterwards the next layer may be issued /
  // z_of(buf) may be read.
  __device__ void epilogue(int buf, int layer) {
    if (d.act == CGSVMC_ACT_RELU) epilogue_impl<true>(buf, layer);
    else epilogue_impl<false>(buf, layer);
  }

  template <bool RELU>
  __device__ __forceinline__ void epilogue_impl(int buf, int layer) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    FC_CLOCK(t_in);
    mbar_wait(mma_bar + buf, phase[buf]);
    FC_CLOCK(t_go);
    FC_ADD(0, t_go - t_in);
    phase[buf] ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int q = warp & 3, half = warp >> 2;
    const int r = 32 * q + lane;
    const bool last = layer == d.L - 1;
    const float* bj = bias_s + layer * H + half * HH;
    const float* wo = wout_s + half * HH;
    const uint32_t trow = tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(buf * 3 * H + half * HH);
    char* abuf = act_buf(buf);
    const int act = d.act;
    float zacc = 0.f;
    // all TMEM loads of this thread's columns in flight, one wait (the math
    // below then has the whole tile row to schedule from)
    uint32_t pa[HH / 8][8], pb[HH / 8][8], pc[HH / 8][8];
#pragma unroll
    for (int c8 = 0; c8 < HH / 8; ++c8) {
      tmem_ld8_nowait(trow + (uint32_t)(8 * c8), pa[c8]);
      tmem_ld8_nowait(trow + (uint32_t)(H + 8 * c8), pb[c8]);
      tmem_ld8_nowait(trow + (uint32_t)(2 * H + 8 * c8), pc[c8]);
    }
    tmem_ld_wait();
#pragma unroll
    for (int c8 = 0; c8 < HH / 8; ++c8) {
      const uint32_t (&p0)[8] = pa[c8];
      const uint32_t (&p1)[8] = pb[c8];
      const uint32_t (&p2)[8] = pc[c8];
      float v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float pre = fmaf(fmaf(__uint_as_float(p2[k]), 1.f / kSplitScale, __uint_as_float(p1[k])),
                               1.f / kSplitScale, __uint_as_float(p0[k])) + bj[8 * c8 + k];
        v[k] = RELU ? fmaxf(pre, 0.f) : activate_slow(act, pre);
      }
      if (last) {
#pragma unroll
        for (int k = 0; k < 8; ++k) zacc = fmaf(v[k], wo[8 * c8 + k], zacc);
      } else {
        __half2 hs[3][4];
#pragma unroll
        for (int k = 0; k < 4; ++k) split3_pair(v[2 * k], v[2 * k + 1], hs[0][k], hs[1][k], hs[2][k]);
        const int kc = (half * HH) / 8 + c8;           // 8-feature chunk of the next layer's K
#pragma unroll
        for (int sp = 0; sp < 3; ++sp) {
          uint4 pk;
          pk.x = *reinterpret_cast<const uint32_t*>(&hs[sp][0]);
          pk.y = *reinterpret_cast<const uint32_t*>(&hs[sp][1]);
          pk.z = *reinterpret_cast<const uint32_t*>(&hs[sp][2]);
          pk.w = *reinterpret_cast<const uint32_t*>(&hs[sp][3]);
          *reinterpret_cast<uint4*>(abuf + ((size_t)(sp * CH + kc) * kTile + r) * 16) = pk;
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (!last) {
      // this thread's share of the next operand planes is written and its TMEM
      // reads are done: tell the issuer (release; it issues after all 256 arrived)
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(ready_bar + buf)) : "memory");
    } else {
      zpart_s[(buf * 2 + half) * kTile + r] = zacc;
      asm volatile("bar.sync 1, %0;" ::"n"(kWorkers) : "memory");      // the epilogue warps only
      if (threadIdx.x < kTile)
        z_s[buf * kTile + threadIdx.x] = (zpart_s[(buf * 2) * kTile + threadIdx.x] +
                                          zpart_s[(buf * 2 + 1) * kTile + threadIdx.x]) + wout_s[H];
    }
    FC_CLOCK(t_out);
    FC_ADD(1, t_out - t_go);
    FC_ADD(2, 1);
  }

  // Forward pass of the tiles whose input planes have been written
  // (write_input; n_tiles = 1 or 2): z_of(buf)[r] for every row.  With two
  // tiles the MMAs of one overlap the epilogue of the other.
  __device__ __noinline__ void forward(int n_tiles) {
    FC_CLOCK(t_f0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if ((threadIdx.x >> 5) == kWorkers / 32) {
      // ---- MMA issuer: tile 0 layer 0, tile 1 layer 0, tile 0 layer 1 (once its planes are ready), ...
      for (int layer = 0; layer < d.L; ++layer)
        for (int buf = 0; buf < n_tiles; ++buf) {
          if (layer > 0) {
            mbar_wait(ready_bar + buf, phase[buf]);
            phase[buf] ^= 1u;
          }
          issue(buf, layer);
        }
    } else {
      // ---- epilogue warps
      for (int layer = 0; layer < d.L; ++layer)
        for (int buf = 0; buf < n_tiles; ++buf) epilogue(buf, layer);
    }
    __syncthreads();          // z_of(*) complete, every MMA consumed
    FC_CLOCK(t_f1);
    FC_ADD(3, t_f1 - t_f0);
    FC_ADD(4, 1);
  }
};

// ---------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------
template <int H>
__global__ void __launch_bounds__(kThreads, 1)
fc_log_amp_kernel(FcDesc d, const uint64_t* __restrict__ packed, int64_t B, float* __restrict__ out) {
  extern __shared__ __align__(1024) char smem[];
  FcEngine<H> eng(d);
  eng.setup(smem);
  uint64_t* cfg = reinterpret_cast<uint64_t*>(smem + fc_plan(d).total);    // [nbuf][128][NW]
  const int per_cta = d.nbuf * kTile;
  const int64_t n_groups = (B + per_cta - 1) / per_cta;
  for (int64_t grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
    const int64_t b0 = grp * per_cta;
    const int n_items = (int)min((int64_t)per_cta, B - b0);
    const int n_tiles = (n_items + kTile - 1) / kTile;
    for (int e = threadIdx.x; e < n_items * d.NW; e += kThreads) cfg[e] = packed[b0 * d.NW + e];
    __syncthreads();
    for (int t = 0; t < n_tiles; ++t)
      eng.write_input(t, cfg + (size_t)t * kTile * d.NW, min(kTile, n_items - t * kTile));
    eng.forward(n_tiles);
    for (int e = threadIdx.x; e < n_items; e += kThreads) out[b0 + e] = eng.z_s[e];
    __syncthreads();
  }
  eng.teardown();
}

__device__ __forceinline__ int kth_set_bit(const uint64_t* words, int nw, int k) {
  for (int w = 0; w < nw; ++w) {
    uint64_t m = words[w];
    const int c = __popcll(m);
    if (k >= c) { k -= c; continue; }
    for (int qd = 0; qd < k; ++qd) m &= m - 1;
    return w * 64 + __ffsll((long long)m) - 1;
  }
  return 0;
}

// graph_builders.py:54-89 x n_steps: a CTA owns up to nbuf x 128 walkers for all
// steps, one forward pass per proposal (the reference runs two), z of the
// current configuration cached.  Same Philox stream as every other sampler.
template <int H>
__global__ void __launch_bounds__(kThreads, 1)
fc_mc_kernel(FcDesc d, uint64_t* __restrict__ packed, int64_t B, int per_cta, int n_steps, uint64_t seed,
             uint64_t walker0, uint64_t step0, const uint64_t* __restrict__ step0_dev,
             unsigned long long* accept_count, float* __restrict__ log_amp_out) {
  if (step0_dev != nullptr) step0 += *step0_dev;
  extern __shared__ __align__(1024) char smem[];
  FcEngine<H> eng(d);
  eng.setup(smem);
  char* p = smem + fc_plan(d).total;
  uint64_t* cfg = reinterpret_cast<uint64_t*>(p); p += (size_t)d.nbuf * kTile * d.NW * 8;   // proposals
  uint64_t* cur = reinterpret_cast<uint64_t*>(p); p += (size_t)d.nbuf * kTile * d.NW * 8;
  float* z_cur = reinterpret_cast<float*>(p); p += (size_t)d.nbuf * kTile * 4;
  float* u_acc = reinterpret_cast<float*>(p);
  __shared__ unsigned int n_acc_s;
  if (threadIdx.x == 0) n_acc_s = 0;
  const int64_t n_groups = (B + per_cta - 1) / per_cta;
  for (int64_t grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
    const int64_t b0 = grp * per_cta;
    const int n_items = (int)min((int64_t)per_cta, B - b0);
    const int n_tiles = (n_items + kTile - 1) / kTile;
    for (int e = threadIdx.x; e < n_items * d.NW; e += kThreads) {
      const uint64_t w = packed[b0 * d.NW + e];
      cur[e] = w;
      cfg[e] = w;
    }
    __syncthreads();
    for (int t = 0; t < n_tiles; ++t)
      eng.write_input(t, cfg + (size_t)t * kTile * d.NW, min(kTile, n_items - t * kTile));
    eng.forward(n_tiles);
    for (int g = threadIdx.x; g < n_items; g += kThreads) z_cur[g] = eng.z_s[g];
    __syncthreads();
    for (int step = 0; step < n_steps; ++step) {
      for (int g = threadIdx.x; g < n_items; g += kThreads) {
        uint64_t s[CGSVMC_MAX_WORDS], dn[CGSVMC_MAX_WORDS];
        int n_up = 0;
        for (int w = 0; w < d.NW; ++w) {
          s[w] = cur[g * d.NW + w];
          dn[w] = ~s[w] & valid_mask_word(d.N, w);
          n_up += __popcll(s[w]);
        }
        const int n_dn = d.N - n_up;
        float u = 2.f;     // > any probability: never accepted
        if (n_up > 0 && n_dn > 0) {
          const Philox4 r = walker_step_random(seed, walker0 + (uint64_t)(b0 + g), step0 + (uint64_t)step);
          const int up = kth_set_bit(s, d.NW, (int)__umulhi(r.x, (uint32_t)n_up));
          const int dns = kth_set_bit(dn, d.NW, (int)__umulhi(r.y, (uint32_t)n_dn));
          s[up >> 6] ^= 1ull << (up & 63);
          s[dns >> 6] ^= 1ull << (dns & 63);
          u = u32_to_unit(r.z);
        }
        for (int w = 0; w < d.NW; ++w) cfg[g * d.NW + w] = s[w];
        u_acc[g] = u;
      }
      __syncthreads();
      for (int t = 0; t < n_tiles; ++t)
        eng.write_input(t, cfg + (size_t)t * kTile * d.NW, min(kTile, n_items - t * kTile));
      eng.forward(n_tiles);
      for (int g = threadIdx.x; g < n_items; g += kThreads) {
        const float zn = eng.z_s[g];
        const float prob = fast_exp(2.f * (zn - z_cur[g]));
        if (prob > u_acc[g]) {   // strict; NaN rejects (graph_builders.py:75-79)
          for (int w = 0; w < d.NW; ++w) cur[g * d.NW + w] = cfg[g * d.NW + w];
          z_cur[g] = zn;
          atomicAdd(&n_acc_s, 1u);
        }
      }
      __syncthreads();
    }
    for (int e = threadIdx.x; e < n_items * d.NW; e += kThreads) packed[b0 * d.NW + e] = cur[e];
    if (log_amp_out != nullptr)
      for (int g = threadIdx.x; g < n_items; g += kThreads) log_amp_out[b0 + g] = z_cur[g];
    __syncthreads();
  }
  if (threadIdx.x == 0 && accept_count != nullptr && n_acc_s)
    atomicAdd(accept_count, (unsigned long long)n_acc_s);
  eng.teardown();
}

// operators.py:227-259: a CTA takes `wch` walkers at a time, lists (walker,
// antiparallel bond) items behind each walker's base configuration and
// evaluates them in tiles of 128.
template <int H>
__global__ void __launch_bounds__(kThreads, 1)
fc_eloc_kernel(FcDesc d, const int2* __restrict__ bonds_ij, const float* __restrict__ bonds_jx,
               const float* __restrict__ bonds_jz, int n_bonds, const uint64_t* __restrict__ packed,
               int64_t B, int wch, float* __restrict__ e_loc, float* __restrict__ log_amp_out,
               float* __restrict__ diag_out, float* __restrict__ off_out) {
  extern __shared__ __align__(1024) char smem[];
  FcEngine<H> eng(d);
  eng.setup(smem);
  char* p = smem + fc_plan(d).total;
  uint64_t* cfg = reinterpret_cast<uint64_t*>(p); p += (size_t)d.nbuf * kTile * d.NW * 8;   // item configurations
  uint64_t* base = reinterpret_cast<uint64_t*>(p); p += (size_t)wch * d.NW * 8;             // walker configurations
  float* z_item = reinterpret_cast<float*>(p); p += (size_t)kMaxItems * 4;
  uint32_t* item = reinterpret_cast<uint32_t*>(p); p += (size_t)kMaxItems * 4;              // walker << 16 | bond + 1
  int* first = reinterpret_cast<int*>(p); p += (size_t)(wch + 1) * 4;                       // first item of a walker
  float* diag_s = reinterpret_cast<float*>(p);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int per_pass = d.nbuf * kTile;
  const int64_t n_chunks = (B + wch - 1) / wch;
  for (int64_t chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
    const int64_t b0 = chunk * wch;
    const int n_w = (int)min((int64_t)wch, B - b0);
    for (int e = threadIdx.x; e < n_w * d.NW; e += kThreads) base[e] = packed[b0 * d.NW + e];
    __syncthreads();
    if (warp == 0) {   // ordered item list (walker by walker, base first) + diagonal terms
      int count = 0;
      for (int w = 0; w < n_w; ++w) {
        const uint64_t* s = base + (size_t)w * d.NW;
        if (lane == 0) { first[w] = count; item[count] = (uint32_t)w << 16; }
        ++count;
        float diag = 0.f;
        for (int k0 = 0; k0 < n_bonds; k0 += 32) {
          const int k = k0 + lane;
          bool anti = false;
          if (k < n_bonds) {
            const int2 bd = bonds_ij[k];
            anti = word_bit(s, bd.x) != word_bit(s, bd.y);
            diag += (anti ? -0.25f : 0.25f) * bonds_jz[k];            // operators.py:165,169
          }
          const uint32_t vote = __ballot_sync(CGSVMC_FULL_MASK, anti);
          if (anti) item[count + __popc(vote & ((1u << lane) - 1u))] = ((uint32_t)w << 16) | (uint32_t)(k + 1);
          count += __popc(vote);
        }
        diag = warp_sum(diag);
        if (lane == 0) diag_s[w] = diag;
      }
      if (lane == 0) first[n_w] = count;
    }
    __syncthreads();
    const int n_items = first[n_w];
    for (int i0 = 0; i0 < n_items; i0 += per_pass) {
      const int n_pass = min(per_pass, n_items - i0);
      const int n_tiles = (n_pass + kTile - 1) / kTile;
      for (int e = threadIdx.x; e < n_pass * d.NW; e += kThreads) {
        const int g = e / d.NW, w = e - g * d.NW;
        const uint32_t it = item[i0 + g];
        uint64_t word = base[(size_t)(it >> 16) * d.NW + w];
        const int bond = (int)(it & 0xffffu);
        if (bond > 0) {                                               // operators.py:158-164
          const int2 bd = bonds_ij[bond - 1];
          if ((bd.x >> 6) == w) word ^= 1ull << (bd.x & 63);
          if ((bd.y >> 6) == w) word ^= 1ull << (bd.y & 63);
        }
        cfg[e] = word;
      }
      __syncthreads();
      for (int t = 0; t < n_tiles; ++t)
        eng.write_input(t, cfg + (size_t)t * kTile * d.NW, min(kTile, n_pass - t * kTile));
      eng.forward(n_tiles);
      for (int g = threadIdx.x; g < n_pass; g += kThreads) z_item[i0 + g] = eng.z_s[g];
      __syncthreads();
    }
    // E_loc = diag + sum_active jx/2 exp(z' - z)   (operators.py:168-169, 259): one warp per walker
    for (int w = warp; w < n_w; w += kWarps) {
      const int i_first = first[w], i_end = first[w + 1];
      const float z0 = z_item[i_first];
      float off = 0.f;
      for (int it = i_first + 1 + lane; it < i_end; it += 32)
        off = fmaf(0.5f * bonds_jx[(item[it] & 0xffffu) - 1], fast_exp(z_item[it] - z0), off);
      off = warp_sum(off);
      if (lane == 0) {
        const int64_t b = b0 + w;
        e_loc[b] = diag_s[w] + off;
        if (log_amp_out) log_amp_out[b] = z0;
        if (diag_out) diag_out[b] = diag_s[w];
        if (off_out) off_out[b] = off;
      }
    }
    __syncthreads();
  }
  eng.teardown();
}

// ---------------------------------------------------------------------------
// parameter image: split B-operand planes of every layer, biases, output layer
// ---------------------------------------------------------------------------
__device__ __forceinline__ __half split_part(float w, int split) {
  __half h1, h2, h3;
  split3(w, h1, h2, h3);
  return split == 0 ? h1 : split == 1 ? h2 : h3;
}

// w_off / b_off: flat offsets of the L + 1 weight matrices and biases.
__global__ void fc_prep_kernel(int N, int H, int L, int K1, const float* __restrict__ params,
                               const int64_t* __restrict__ w_off, const int64_t* __restrict__ b_off,
                               __half* __restrict__ wimg, int64_t wimg_halfs, float* __restrict__ consts) {
  const int64_t first_halfs = (int64_t)K1 * 3 * H;
  const int64_t per_layer = (int64_t)H * 3 * H;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < wimg_halfs;
       e += (int64_t)gridDim.x * blockDim.x) {
    // [chunk][n = split * H + j][8 input features]
    int layer = 0;
    int64_t r = e;
    if (e >= first_halfs) { layer = 1 + (int)((e - first_halfs) / per_layer); r = (e - first_halfs) % per_layer; }
    const int chunk = (int)(r / (3 * H * 8)); r -= (int64_t)chunk * 3 * H * 8;
    const int n = (int)(r / 8), el = (int)(r - 8 * n);
    const int split = n / H, j = n - split * H;
    const int k = chunk * 8 + el;
    const int k_in = layer == 0 ? N : H;
    const float w = k < k_in ? params[w_off[layer] + (int64_t)k * H + j] : 0.f;     // W_l[k][j]
    wimg[e] = split_part(w, split);
  }
  if (blockIdx.x == 0) {
    for (int e = threadIdx.x; e < L * H; e += blockDim.x) consts[e] = params[b_off[e / H] + e % H];
    for (int j = threadIdx.x; j < H; j += blockDim.x) consts[L * H + j] = params[w_off[L] + j];   // w_out [H, 1]
    if (threadIdx.x == 0) consts[L * H + H] = params[b_off[L]];
  }
}

bool fc_tc_enabled() {
  const char* e = getenv("CGSVMC_FC_TC");
  return e == nullptr || atoi(e) != 0;
}

// Geometry and shared-memory plan; false when the network is outside the
// tensor-core path (the SIMT tile kernels of net.cu take over).  extra(nbuf) is
// the kernel's own shared memory behind the engine's plan.
template <typename F>
bool make_fc_desc(const cgsvmc_ansatz* a, F extra, FcDesc* out) {
  const cgsvmc_ansatz_desc& s = a->desc;
  if (s.kind != CGSVMC_ANSATZ_FULLY_CONNECTED) return false;
  const int H = s.layer_size;
  if (s.num_layers < 1 || s.num_layers > 8) return false;
  if (H != 16 && H != 32 && H != 48 && H != 64 && H != 80) return false;      // 3H <= 256 (UMMA N), H % 16 == 0 (UMMA K)
  FcDesc d;
  memset(&d, 0, sizeof(d));
  d.N = s.n_sites; d.H = H; d.L = s.num_layers; d.act = s.nonlinearity;
  d.NW = n_words(s.n_sites);
  d.K1 = (s.n_sites + 15) / 16 * 16;
  d.n_hh = d.L - 1;
  d.act_bytes = std::max(3 * H, d.K1) * kTile * 2;
  for (int nbuf = 2; nbuf >= 1; --nbuf) {
    d.nbuf = nbuf;
    int cols = 32;
    while (cols < nbuf * 3 * H) cols *= 2;
    d.tmem_cols = cols;
    if (fc_plan(d).total + extra(d) + 1024 <= (size_t)a->max_smem_optin) { *out = d; return true; }
  }
  return false;
}

size_t log_amp_extra(const FcDesc& d) { return (size_t)d.nbuf * kTile * d.NW * 8 + 16; }
size_t mc_extra(const FcDesc& d) { return (size_t)d.nbuf * kTile * (d.NW * 16 + 8) + 16; }
struct ElocExtra {
  int wch;
  size_t operator()(const FcDesc& d) const {
    return (size_t)d.nbuf * kTile * d.NW * 8 + (size_t)wch * d.NW * 8 + (size_t)kMaxItems * 8 +
           (size_t)(wch + 1) * 4 + (size_t)wch * 4 + 16;
  }
};

int build_fc_image(cgsvmc_ansatz* a, FcDesc* d, cudaStream_t st) {
  const int64_t wimg_halfs = ((int64_t)d->K1 + (int64_t)d->n_hh * d->H) * 3 * d->H;
  const size_t off_bytes = (size_t)2 * (d->L + 1) * sizeof(int64_t);
  const size_t off_pad = (off_bytes + 255) / 256 * 256;
  const size_t const_bytes = (((size_t)d->L * d->H + d->H + 4) * 4 + 255) / 256 * 256;
  const size_t bytes = off_pad + const_bytes + (size_t)wimg_halfs * 2;
  if (a->tables_bytes < bytes) {
    if (a->tables != nullptr) {
      if (int rc = cuda_fail(cudaDeviceSynchronize(), "tables sync")) return rc;
      cudaFree(a->tables);
      a->tables = nullptr;
      a->tables_bytes = 0;
    }
    if (int rc = cuda_fail(cudaMalloc(&a->tables, bytes), "tables alloc")) return rc;
    a->tables_bytes = bytes;
    a->tables_valid = false;
    // flat offsets of every layer's weights and biases (layout of include/cgsvmc.h: w, b per layer)
    std::vector<int64_t> offs(2 * (d->L + 1));
    for (int l = 0; l <= d->L; ++l) { offs[l] = a->offsets[2 * l]; offs[d->L + 1 + l] = a->offsets[2 * l + 1]; }
    if (int rc = cuda_fail(cudaMemcpy(a->tables, offs.data(), off_bytes, cudaMemcpyHostToDevice),
                           "tables offsets"))
      return rc;
  }
  char* base = reinterpret_cast<char*>(a->tables);
  const int64_t* w_off = reinterpret_cast<const int64_t*>(base);
  const int64_t* b_off = w_off + d->L + 1;
  float* consts = reinterpret_cast<float*>(base + off_pad);
  __half* wimg = reinterpret_cast<__half*>(base + off_pad + const_bytes);
  d->wimg = wimg;
  d->consts = consts;
  if (a->track_params && a->tables_valid) return CGSVMC_OK;
  const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((wimg_halfs + 255) / 256, 592));
  fc_prep_kernel<<<blocks, 256, 0, st>>>(d->N, d->H, d->L, d->K1, a->params, w_off, b_off, wimg, wimg_halfs,
                                         consts);
  a->tables_valid = true;
  return cuda_fail(cudaGetLastError(), "fc_tc prep launch");
}

template <typename F>
int opt_in(F kernel, size_t bytes) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(smem)");
  return CGSVMC_OK;
}

// Launches KERNEL<H> for the plan `d`.
#define CGSVMC_FC_LAUNCH(KERNEL, grid, smem, st, ...)                       \
  do {                                                                      \
    switch (d.H) {                                                          \
      case 16: if (int rc = opt_in(KERNEL<16>, smem)) return rc;            \
               KERNEL<16><<<grid, kThreads, smem, st>>>(__VA_ARGS__); break; \
      case 32: if (int rc = opt_in(KERNEL<32>, smem)) return rc;            \
               KERNEL<32><<<grid, kThreads, smem, st>>>(__VA_ARGS__); break; \
      case 48: if (int rc = opt_in(KERNEL<48>, smem)) return rc;            \
               KERNEL<48><<<grid, kThreads, smem, st>>>(__VA_ARGS__); break; \
      case 64: if (int rc = opt_in(KERNEL<64>, smem)) return rc;            \
               KERNEL<64><<<grid, kThreads, smem, st>>>(__VA_ARGS__); break; \
      default: if (int rc = opt_in(KERNEL<80>, smem)) return rc;            \
               KERNEL<80><<<grid, kThreads, smem, st>>>(__VA_ARGS__); break; \
    }                                                                       \
  } while (0)

// walkers per local-energy chunk: enough CTAs to fill the device, at most
// kMaxItems items (every bond of every walker active) per chunk
int eloc_chunk(const cgsvmc_ansatz* a, const cgsvmc_ham* h, int64_t B) {
  const int64_t cap = std::max<int64_t>(1, kMaxItems / (h->n_bonds + 1));
  const int64_t want = std::max<int64_t>(1, (B + a->num_sms - 1) / a->num_sms);
  return (int)std::min<int64_t>(std::min<int64_t>(cap, want), 64);
}

}  // namespace

// The split weight image and constants of a supported network (built when the
// parameters changed), for the gradient kernel of fc_tc_grad.cu.
int fc_tc_image(cgsvmc_ansatz* a, const void** wimg, const float** consts, cudaStream_t st) {
  FcDesc d;
  if (!make_fc_desc(a, mc_extra, &d)) { set_error("fc_tc: unsupported network"); return CGSVMC_ERR_UNSUPPORTED; }
  if (int rc = build_fc_image(a, &d, st)) return rc;
  *wimg = d.wimg;
  *consts = d.consts;
  return CGSVMC_OK;
}

bool fc_tc_supported(const cgsvmc_ansatz* a, const cgsvmc_ham* h) {
  if (!fc_tc_enabled()) return false;
  FcDesc d;
  if (h != nullptr) {
    if (h->n_bonds + 1 > kMaxItems || h->n_bonds >= 65535) return false;
    return make_fc_desc(a, ElocExtra{64}, &d);
  }
  return make_fc_desc(a, mc_extra, &d);
}

int fc_tc_log_amp(cgsvmc_ansatz* a, const uint64_t* packed, int64_t B, float* out, cudaStream_t st) {
  FcDesc d;
  if (!make_fc_desc(a, log_amp_extra, &d)) { set_error("fc_tc: unsupported network"); return CGSVMC_ERR_UNSUPPORTED; }
  if (B <= (int64_t)kTile * a->num_sms) d.nbuf = 1;      // few walkers: one tile per CTA, more CTAs
  if (int rc = build_fc_image(a, &d, st)) return rc;
  const size_t smem = fc_plan(d).total + log_amp_extra(d);
  const int per_cta = d.nbuf * kTile;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((B + per_cta - 1) / per_cta, a->num_sms));
  CGSVMC_FC_LAUNCH(fc_log_amp_kernel, grid, smem, st, d, packed, B, out);
  return cuda_fail(cudaGetLastError(), "fc_tc log_amp launch");
}

int fc_tc_mc_steps(cgsvmc_ansatz* a, uint64_t* packed, int64_t B, int n_steps, uint64_t seed,
                   uint64_t walker0, uint64_t step0, unsigned long long* accept_count,
                   float* log_amp_out, cudaStream_t st) {
  FcDesc d;
  if (!make_fc_desc(a, mc_extra, &d)) { set_error("fc_tc: unsupported network"); return CGSVMC_ERR_UNSUPPORTED; }
  if (B <= (int64_t)kTile * a->num_sms) d.nbuf = 1;
  if (int rc = build_fc_image(a, &d, st)) return rc;
  const size_t smem = fc_plan(d).total + mc_extra(d);
  // spread few walkers over the SMs (a tile need not be full)
  int per_cta = d.nbuf * kTile;
  if (d.nbuf == 1) per_cta = (int)std::min<int64_t>(kTile, std::max<int64_t>(8, (B + a->num_sms - 1) / a->num_sms));
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((B + per_cta - 1) / per_cta, a->num_sms));
  CGSVMC_FC_LAUNCH(fc_mc_kernel, grid, smem, st, d, packed, B, per_cta, n_steps, seed, walker0, step0,
                   a->step_counter_dev, accept_count, log_amp_out);
  return cuda_fail(cudaGetLastError(), "fc_tc mc launch");
}

int fc_tc_local_energy(cgsvmc_ansatz* a, const cgsvmc_ham* h, const uint64_t* packed, int64_t B,
                       float* e_loc, float* log_amp_out, float* diag_out, float* off_out,
                       cudaStream_t st) {
  FcDesc d;
  const int wch = eloc_chunk(a, h, B);
  if (!make_fc_desc(a, ElocExtra{wch}, &d)) { set_error("fc_tc: unsupported network"); return CGSVMC_ERR_UNSUPPORTED; }
  // expected items per chunk ~ wch (1 + n_bonds / 2): one tile in flight when they fit one
  if ((int64_t)wch * (1 + h->n_bonds / 2) <= kTile) d.nbuf = 1;
  if (int rc = build_fc_image(a, &d, st)) return rc;
  const size_t smem = fc_plan(d).total + ElocExtra{wch}(d);
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((B + wch - 1) / wch, a->num_sms));
  CGSVMC_FC_LAUNCH(fc_eloc_kernel, grid, smem, st, d, h->ij, h->jx, h->jz, h->n_bonds, packed, B, wch, e_loc,
                   log_amp_out, diag_out, off_out);
  return cuda_fail(cudaGetLastError(), "fc_tc local_energy launch");
}

}  // namespace cgsvmc

#ifdef CGSVMC_RBM2_TIMING
// Development build only: reads (and clears) the fc_tc phase counters.
extern "C" int cgsvmc_debug_fc_tc_phases(unsigned long long* host_out) {
  cudaError_t e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaMemcpyFromSymbol(host_out, cgsvmc::g_fc_phase, 6 * sizeof(unsigned long long));
  unsigned long long zero[6] = {0, 0, 0, 0, 0, 0};
  if (e == cudaSuccess) e = cudaMemcpyToSymbol(cgsvmc::g_fc_phase, zero, sizeof(zero));
  return e == cudaSuccess ? 0 : -2;
}
#endif
