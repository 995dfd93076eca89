// Metropolis sampler of the fully connected ansatz (wavefunctions.py:345-371;
// graph_builders.py:54-89) for SMALL walker batches: one warp per walker.
//
// The tensor-core sampler of fc_tc.cu evaluates tiles of 128 walkers and needs
// four layer round trips (operand planes -> MMA -> TMEM -> epilogue) of ~2.5 us
// per Metropolis step whatever the batch; at the BASELINE C1 size (1,024
// walkers = 8 tiles' worth of rows) the sweep is bound by that latency, 10 us
// per step.  Here every walker is one warp with no CTA-wide barrier on its
// path: all parameters sit in shared memory (C1: 59 KB), lane l owns the hidden
// units l, l + 32, l + 64 of every layer (conflict-free weight reads), the
// previous layer's activations are broadcast from a warp-private shared-memory
// row (one LDS.128 per four inputs), and the proposal -- Philox block, k-th
// set-bit selection by ballot -- is computed redundantly by all lanes.  FP32
// FMA chains in a fixed order: float32 grade like net.cu.  Chosen by
// cgsvmc_mc_steps for batches of at most 2,048 walkers (CGSVMC_FC_WARP=0 / 1
// forces it off / on at any size): above that the tensor-core tiles win.
// Same proposal stream and acceptance rule as every other sampler of the
// library (Philox keyed by global walker id and step; accept iff
// exp(2 (z' - z)) > u, strict).
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "internal.h"

namespace cgsvmc {
namespace {

constexpr int kWarps = 8;
constexpr int kThreads = 32 * kWarps;
constexpr int kMaxUnits = 3;          // hidden units per lane: H <= 96
constexpr int kMaxLayers = 8;
constexpr int kHPad = 96;

struct FwDesc {
  int N, L, H, act, words;
  int64_t P;
  int64_t w_off[kMaxLayers + 1], b_off[kMaxLayers + 1];   // flat offsets; entry L = output layer
};

__device__ __forceinline__ float fw_act(int act, float x) {
  switch (act) {
    case CGSVMC_ACT_RELU: return fmaxf(x, 0.f);
    case CGSVMC_ACT_TANH: return tanh_accurate(x);
    case CGSVMC_ACT_SIGMOID: return 1.f / (1.f + expf(-x));
    case CGSVMC_ACT_IDENTITY: return x;
    case CGSVMC_ACT_COS: return cosf(x);
    case CGSVMC_ACT_EXP: return expf(x);
    case CGSVMC_ACT_SELU:
      return x > 0.f ? 1.0507009873554805f * x : 1.7580993408473766f * (expf(x) - 1.f);
    default: return tanf(x);
  }
}

// z(sigma) of the WPW configurations of a warp.  `par`: the flat parameter
// vector in shared memory; `hbuf`: this warp's activation rows [2][WPW][kHPad].
// Every weight read serves the WPW walkers of the warp (the kernel is bound by
// shared-memory wavefronts: one per weight and warp).
template <int NW, int WPW>
__device__ __forceinline__ void forward(const FwDesc& d, const float* __restrict__ par, float* hbuf,
                                        const uint64_t (&s)[WPW][NW], int lane, float (&z)[WPW]) {
  const int H = d.H;
  const bool on[kMaxUnits] = {lane < H, lane + 32 < H, lane + 64 < H};
  float h[WPW][kMaxUnits];
  {   // input layer: +- the rows of W_0, two accumulator chains per unit
    const float* w = par + d.w_off[0];
    const float* b = par + d.b_off[0];
    float acc0[WPW][kMaxUnits], acc1[WPW][kMaxUnits];
#pragma unroll
    for (int q = 0; q < WPW; ++q)
#pragma unroll
      for (int u = 0; u < kMaxUnits; ++u) {
        acc0[q][u] = on[u] ? b[lane + 32 * u] : 0.f;
        acc1[q][u] = 0.f;
      }
    int i = 0;
#pragma unroll 2
    for (; i + 2 <= d.N; i += 2) {
      const float* wr = w + i * H + lane;
      float w0[kMaxUnits], w1[kMaxUnits];
#pragma unroll
      for (int u = 0; u < kMaxUnits; ++u) {
        w0[u] = on[u] ? wr[32 * u] : 0.f;
        w1[u] = on[u] ? wr[H + 32 * u] : 0.f;
      }
#pragma unroll
      for (int q = 0; q < WPW; ++q) {
        const float sg0 = get_bit<NW>(s[q], i) ? 1.f : -1.f, sg1 = get_bit<NW>(s[q], i + 1) ? 1.f : -1.f;
#pragma unroll
        for (int u = 0; u < kMaxUnits; ++u) {
          acc0[q][u] = fmaf(sg0, w0[u], acc0[q][u]);
          acc1[q][u] = fmaf(sg1, w1[u], acc1[q][u]);
        }
      }
    }
    if (i < d.N) {
#pragma unroll
      for (int u = 0; u < kMaxUnits; ++u) {
        const float w0 = on[u] ? w[i * H + lane + 32 * u] : 0.f;
#pragma unroll
        for (int q = 0; q < WPW; ++q) acc0[q][u] = fmaf(get_bit<NW>(s[q], i) ? 1.f : -1.f, w0, acc0[q][u]);
      }
    }
#pragma unroll
    for (int q = 0; q < WPW; ++q)
#pragma unroll
      for (int u = 0; u < kMaxUnits; ++u) h[q][u] = fw_act(d.act, acc0[q][u] + acc1[q][u]);
  }
  for (int l = 1; l < d.L; ++l) {
    float* rows = hbuf + (l & 1) * WPW * kHPad;
#pragma unroll
    for (int q = 0; q < WPW; ++q)
#pragma unroll
      for (int u = 0; u < kMaxUnits; ++u)
        if (on[u]) rows[q * kHPad + lane + 32 * u] = h[q][u];
    __syncwarp();
    const float* w = par + d.w_off[l];
    const float* b = par + d.b_off[l];
    float acc0[WPW][kMaxUnits], acc1[WPW][kMaxUnits];
#pragma unroll
    for (int q = 0; q < WPW; ++q)
#pragma unroll
      for (int u = 0; u < kMaxUnits; ++u) {
        acc0[q][u] = on[u] ? b[lane + 32 * u] : 0.f;
        acc1[q][u] = 0.f;
      }
#pragma unroll 2
    for (int i = 0; i < H; i += 4) {
      const float* wr = w + i * H + lane;
      float wv[4][kMaxUnits];
#pragma unroll
      for (int e = 0; e < 4; ++e)
#pragma unroll
        for (int u = 0; u < kMaxUnits; ++u) wv[e][u] = on[u] ? wr[e * H + 32 * u] : 0.f;
#pragma unroll
      for (int q = 0; q < WPW; ++q) {
        const float4 x = *reinterpret_cast<const float4*>(rows + q * kHPad + i);
#pragma unroll
        for (int u = 0; u < kMaxUnits; ++u) {
          acc0[q][u] = fmaf(x.x, wv[0][u], acc0[q][u]);
          acc1[q][u] = fmaf(x.y, wv[1][u], acc1[q][u]);
          acc0[q][u] = fmaf(x.z, wv[2][u], acc0[q][u]);
          acc1[q][u] = fmaf(x.w, wv[3][u], acc1[q][u]);
        }
      }
    }
#pragma unroll
    for (int q = 0; q < WPW; ++q)
#pragma unroll
      for (int u = 0; u < kMaxUnits; ++u) h[q][u] = fw_act(d.act, acc0[q][u] + acc1[q][u]);
  }
  // output layer [H, 1]
  const float* wo = par + d.w_off[d.L];
  const float bo = par[d.b_off[d.L]];
#pragma unroll
  for (int q = 0; q < WPW; ++q) {
    float zq = 0.f;
#pragma unroll
    for (int u = 0; u < kMaxUnits; ++u)
      if (on[u]) zq = fmaf(h[q][u], wo[lane + 32 * u], zq);
    z[q] = warp_sum(zq) + bo;
  }
}

template <int NW, int WPW>
__global__ void __launch_bounds__(kThreads)
fc_warp_mc_kernel(FwDesc d, const float* __restrict__ params, uint64_t* __restrict__ packed, int64_t B,
                  int n_steps, uint64_t seed, uint64_t walker0, uint64_t step0,
                  const uint64_t* __restrict__ step0_dev, unsigned long long* accept_count,
                  float* __restrict__ log_amp_out) {
  extern __shared__ __align__(16) float smem[];
  float* par = smem;
  const int64_t p_pad = (d.P + 3) / 4 * 4;
  float* hbuf = smem + p_pad + (threadIdx.x >> 5) * 2 * WPW * kHPad;
  for (int64_t e = threadIdx.x; e < d.P; e += kThreads) par[e] = params[e];
  if (step0_dev != nullptr) step0 += *step0_dev;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned int n_acc = 0;
  for (int64_t b0 = ((int64_t)blockIdx.x * kWarps + warp) * WPW; b0 < B; b0 += (int64_t)gridDim.x * kWarps * WPW) {
    uint64_t s[WPW][NW];
    bool live[WPW];
#pragma unroll
    for (int q = 0; q < WPW; ++q) {
      live[q] = b0 + q < B;
#pragma unroll
      for (int w = 0; w < NW; ++w) s[q][w] = (live[q] && w < d.words) ? packed[(b0 + q) * d.words + w] : 0ull;
    }
    float z_cur[WPW];
    forward<NW, WPW>(d, par, hbuf, s, lane, z_cur);
    for (int step = 0; step < n_steps; ++step) {
      uint64_t prop[WPW][NW];
      float u_acc[WPW];
#pragma unroll
      for (int q = 0; q < WPW; ++q) {
        uint64_t dn[NW];
        int n_up = 0;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
          dn[w] = ~s[q][w] & valid_mask_word(d.N, w);
          n_up += __popcll(s[q][w]);
          prop[q][w] = s[q][w];
        }
        const int n_dn = d.N - n_up;
        u_acc[q] = 2.f;                 // > any probability: padding / frozen walkers never accept
        if (live[q] && n_up > 0 && n_dn > 0) {      // (uniform over the warp)
          const Philox4 r = walker_step_random(seed, walker0 + (uint64_t)(b0 + q), step0 + (uint64_t)step);
          const int up = select_kth_bit<NW>(s[q], (int)__umulhi(r.x, (uint32_t)n_up), lane);
          const int dnsite = select_kth_bit<NW>(dn, (int)__umulhi(r.y, (uint32_t)n_dn), lane);
          flip_bit<NW>(prop[q], up);
          flip_bit<NW>(prop[q], dnsite);
          u_acc[q] = u32_to_unit(r.z);
        }
      }
      float z_new[WPW];
      forward<NW, WPW>(d, par, hbuf, prop, lane, z_new);
      // accept iff |psi'/psi|^2 > u  (graph_builders.py:75-79 squared; strict, NaN rejects)
#pragma unroll
      for (int q = 0; q < WPW; ++q)
        if (fast_exp(2.f * (z_new[q] - z_cur[q])) > u_acc[q]) {
#pragma unroll
          for (int w = 0; w < NW; ++w) s[q][w] = prop[q][w];
          z_cur[q] = z_new[q];
          ++n_acc;
        }
    }
    if (lane == 0) {
#pragma unroll
      for (int q = 0; q < WPW; ++q) {
        if (!live[q]) continue;
#pragma unroll
        for (int w = 0; w < NW; ++w) if (w < d.words) packed[(b0 + q) * d.words + w] = s[q][w];
        if (log_amp_out != nullptr) log_amp_out[b0 + q] = z_cur[q];
      }
    }
  }
  if (lane == 0 && accept_count != nullptr && n_acc) atomicAdd(accept_count, (unsigned long long)n_acc);
}

// Walkers per warp.  Measured at C1 (1,024 walkers, 20 steps): 1 -> 126-132 us, 2 -> 151 us (the
// second walker halves the shared-memory wavefronts per walker but the warp's
// dependent chain gets longer and half as many warps hide it); fc_tc.cu: 199 us.
constexpr int kWalkersPerWarp = 1;

size_t smem_bytes(const cgsvmc_ansatz* a) {
  return ((size_t)(a->n_params + 3) / 4 * 4 + (size_t)kWarps * 2 * kWalkersPerWarp * kHPad) * sizeof(float);
}

}  // namespace

bool fc_warp_supported(const cgsvmc_ansatz* a, int64_t B) {
  const cgsvmc_ansatz_desc& s = a->desc;
  if (s.kind != CGSVMC_ANSATZ_FULLY_CONNECTED) return false;
  if (s.num_layers < 1 || s.num_layers > kMaxLayers) return false;
  if (s.layer_size < 4 || s.layer_size > 32 * kMaxUnits || s.layer_size % 4 != 0) return false;
  if (s.n_sites > CGSVMC_MAX_SITES) return false;
  if (smem_bytes(a) > (size_t)a->max_smem_optin) return false;
  const char* e = getenv("CGSVMC_FC_WARP");        // read at every call: the tests compare the samplers
  const int forced = e != nullptr ? atoi(e) : -1;
  if (forced == 0) return false;
  if (forced == 1) return true;
  return B <= 2048;
}

int fc_warp_mc_steps(const cgsvmc_ansatz* a, uint64_t* packed, int64_t B, int n_steps, uint64_t seed,
                     uint64_t walker0, uint64_t step0, unsigned long long* accept_count, float* log_amp_out,
                     cudaStream_t st) {
  const cgsvmc_ansatz_desc& s = a->desc;
  FwDesc d;
  d.N = s.n_sites; d.L = s.num_layers; d.H = s.layer_size; d.act = s.nonlinearity; d.P = a->n_params;
  for (int l = 0; l <= s.num_layers; ++l) {
    d.w_off[l] = a->offsets[2 * l];
    d.b_off[l] = a->offsets[2 * l + 1];
  }
  const size_t smem = smem_bytes(a);
  const int per_cta = kWarps * kWalkersPerWarp;
  const int grid = (int)std::min<int64_t>((B + per_cta - 1) / per_cta, (int64_t)a->num_sms * 8);
  const int nw = n_words(s.n_sites);
  d.words = nw;
#define FC_WARP_LAUNCH(NW)                                                                                   \
  do {                                                                                                       \
    auto kern = fc_warp_mc_kernel<NW, kWalkersPerWarp>;                                                      \
    if (smem > 48 * 1024)                                                                                    \
      if (int rc = cuda_fail(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), \
                             "fc_warp smem opt-in"))                                                         \
        return rc;                                                                                           \
    kern<<<grid, kThreads, smem, st>>>(d, a->params, packed, B, n_steps, seed, walker0, step0,               \
                                       a->step_counter_dev, accept_count, log_amp_out);                      \
  } while (0)
  if (nw == 1) FC_WARP_LAUNCH(1);
  else if (nw == 2) FC_WARP_LAUNCH(2);
  else FC_WARP_LAUNCH(4);
#undef FC_WARP_LAUNCH
  return cuda_fail(cudaGetLastError(), "fc_warp_mc_steps launch");
}

}  // namespace cgsvmc
