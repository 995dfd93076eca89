// Generic "tile" networks: fully_connected (wavefunctions.py:328-388), rbm with
// hidden layers (391-452), conv_1d (454-528) and conv_2d (531-615) with the
// periodic padding of layers.py:24-160.
//
// A CTA evaluates a tile of T configurations at a time with all activations in
// shared memory (they never touch HBM); weights stream through L1.  The same
// tile forward pass serves four kernels:
//   log_amp      tiles = consecutive walkers
//   mc_steps     a CTA owns T walkers for all n_steps: propose -> forward ->
//                accept, cached z (one forward per step, the reference runs two)
//   local_energy tiles = (walker, antiparallel bond) pairs generated on the fly
//                from the bit-packed state; only active bonds are evaluated
//   grad         forward + backward per tile, parameter-gradient entries owned
//                by threads, two-stage deterministic reduction
#include "common.cuh"
#include "internal.h"

namespace cgsvmc {
namespace {

constexpr int kThreads = 256;
constexpr int kWarps = 8;
constexpr int kMaxLayers = 16;
constexpr int NWMAX = CGSVMC_MAX_WORDS;

struct NetDesc {
  int kind, N, L, act;
  int D;                      // mlp: hidden width (fc_layer_size)
  int X, Y, C, kx, ky;        // conv: lattice, filters, kernel extents
  int pad_x, pad_y;           // conv: wrap padding placed before the data
  int NW;                     // 64-bit words per configuration
  const float* w[kMaxLayers + 2];   // weight tensor of layer l (head last)
  const float* b[kMaxLayers + 2];   // bias tensor of layer l
  const float* wt[kMaxLayers + 2];  // transposed weights (grad only)
  int64_t w_off[kMaxLayers + 2], b_off[kMaxLayers + 2];   // flat offsets (grad)
  const float* rbm_a; const float* rbm_a0;                // rbm onsite
  int resnet;                 // conv family: layer 0 plain, then (selu conv, conv + skip) blocks
  int stage_weights;          // mlp forward: copy each layer's weights into shared memory first
  int resident_weights;       // mlp forward: ALL layers fit: copied once per kernel (prepare())
};

__device__ __forceinline__ float activate(int act, float x) {
  switch (act) {
    case CGSVMC_ACT_RELU: return fmaxf(x, 0.f);
    case CGSVMC_ACT_TANH: return tanh_accurate(x);
    case CGSVMC_ACT_SIGMOID: return 1.f / (1.f + expf(-x));
    case CGSVMC_ACT_IDENTITY: return x;
    case CGSVMC_ACT_COS: return cosf(x);
    case CGSVMC_ACT_EXP: return expf(x);
    case CGSVMC_ACT_SELU:   // tf.nn.selu: scale * (x > 0 ? x : alpha * (exp(x) - 1))
      return x > 0.f ? 1.0507009873554805f * x : 1.7580993408473766f * (expf(x) - 1.f);
    default: return tanf(x);
  }
}

// d act / d x expressed through the OUTPUT h = act(x) (cos is not expressible
// this way and is rejected on the host for gradients).
__device__ __forceinline__ float activate_grad(int act, float h) {
  switch (act) {
    case CGSVMC_ACT_RELU: return h > 0.f ? 1.f : 0.f;
    case CGSVMC_ACT_TANH: return 1.f - h * h;
    case CGSVMC_ACT_SIGMOID: return h * (1.f - h);
    case CGSVMC_ACT_IDENTITY: return 1.f;
    case CGSVMC_ACT_SELU: return h < 0.f ? h + 1.7580993408473766f : 1.0507009873554805f;
    case CGSVMC_ACT_EXP: return h;
    default: return 1.f + h * h;   // tan
  }
}

__device__ __forceinline__ int word_bit(const uint64_t* words, int site) {
  return (int)((words[site >> 6] >> (site & 63)) & 1ull);
}

// ===========================================================================
// MLP family: fully_connected and rbm (any number of hidden layers)
// ===========================================================================
// Activations live in shared memory as [feature][TS] with TS = T + 4 floats:
// the walkers of a warp are contiguous (one broadcast LDS.128 feeds 4 FMAs
// per weight) and the +4 skew makes the transposed epilogue stores
// conflict-free.
template <int TW>
struct Mlp {
  static constexpr int T = 8 * TW;      // configurations per tile
  static constexpr int TS = T + 4;

  // floats of shared memory for `n_buf` activation buffers
  __host__ __device__ static size_t act_floats(const NetDesc& d) {
    const int dmax = d.D > d.N ? d.D : d.N;
    return (size_t)dmax * TS;
  }
  // one layer's weight matrix [din][dout] staged in shared memory (forward of
  // the non-gradient kernels): the 8 warps of the CTA would otherwise each
  // stream it through L1 (ncu, C1: long-scoreboard stalls 6.3 per issue)
  __host__ __device__ static size_t resident_floats(const NetDesc& d) {
    // layers 0 .. L-1 (and the rbm head matrix): [N][D], then [D][D] each
    const int n_mat = d.L + (d.kind == CGSVMC_ANSATZ_RBM ? 1 : 0);
    if (n_mat == 0) return 0;
    return ((size_t)d.N * d.D + (size_t)(n_mat - 1) * d.D * d.D + 3) / 4 * 4;
  }
  __host__ __device__ static size_t stage_floats(const NetDesc& d) {
    if (d.resident_weights) return resident_floats(d);
    if (!d.stage_weights || TW < 4) return 0;      // small tiles: the copy would outweigh the layer
    const int dmax = d.D > d.N ? d.D : d.N;
    return ((size_t)dmax * d.D + 3) / 4 * 4;
  }
  // offset of matrix l inside the resident copy
  __host__ __device__ static size_t resident_offset(const NetDesc& d, int l) {
    return l == 0 ? 0 : (size_t)d.N * d.D + (size_t)(l - 1) * d.D * d.D;
  }
  // once per kernel, all threads: every weight matrix into shared memory
  __device__ static void prepare(const NetDesc& d, float* smem) {
    if (!d.resident_weights) return;
    float* wall = smem + 2 * act_floats(d);
    const int n_mat = d.L + (d.kind == CGSVMC_ANSATZ_RBM ? 1 : 0);
    for (int l = 0; l < n_mat; ++l) {
      const int n = (l == 0 ? d.N : d.D) * d.D;
      float* dst = wall + resident_offset(d, l);
      for (int e = threadIdx.x; e < n; e += kThreads) dst[e] = __ldg(d.w[l] + e);
    }
    __syncthreads();
  }
  __host__ static size_t forward_smem_bytes(const NetDesc& d) {
    return (2 * act_floats(d) + T + stage_floats(d)) * sizeof(float);
  }
  // CTA-wide copy of W[din * dout] into `wbuf` (all threads call)
  __device__ static const float* stage(const float* __restrict__ W, int din, int dout, float* wbuf) {
    if (wbuf == nullptr) return W;
    __syncthreads();                       // the previous layer may still read wbuf
    const int n = din * dout;
    if ((n & 3) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0) {
      for (int e = threadIdx.x; e < n / 4; e += kThreads)
        reinterpret_cast<float4*>(wbuf)[e] = __ldg(reinterpret_cast<const float4*>(W) + e);
    } else {
      for (int e = threadIdx.x; e < n; e += kThreads) wbuf[e] = __ldg(W + e);
    }
    __syncthreads();
    return wbuf;
  }

  // spins of T configurations -> buf[i][t] = +-1
  __device__ static void load_spins(const NetDesc& d, const uint64_t* cfg /*[T][NW] smem*/,
                                    float* buf) {
    for (int e = threadIdx.x; e < d.N * T; e += kThreads) {
      const int i = e / T, t = e - i * T;
      buf[i * TS + t] = word_bit(cfg + t * d.NW, i) ? 1.f : -1.f;
    }
  }

  // out[j][t] = act(b[j] + sum_i in[i][t] W[i][j]) for the warp's TW walkers.
  // MODE 0: store activations; MODE 1: rbm head (sum_j log cosh -> zacc);
  // MODE 3: rbm head for the gradient, store tanh(pre) = d z / d theta;
  // MODE 4: backward-data with transposed weights, no bias, the result is
  //         multiplied by act'(.) evaluated from the stored output `aux`.
  // STAGED: W points into shared memory (see stage()); otherwise the weights
  // are read through the read-only global path.
  template <int MODE, bool STAGED = false>
  __device__ static void layer(const NetDesc& d, const float* __restrict__ W,
                               const float* __restrict__ bias, int din, int dout, int act,
                               const float* in, float* out, float* zacc,
                               const float* aux = nullptr) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int t0 = warp * TW;
    float zpart[TW];
#pragma unroll
    for (int t = 0; t < TW; ++t) zpart[t] = 0.f;
    for (int j0 = 0; j0 < dout; j0 += 128) {
      float acc[TW][4];
#pragma unroll
      for (int t = 0; t < TW; ++t)
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[t][k] = 0.f;
      const int jbase = j0 + lane;
      for (int i = 0; i < din; ++i) {
        float hv[TW];
        if (TW == 4) {
          const float4 h4 = *reinterpret_cast<const float4*>(in + i * TS + t0);
          hv[0] = h4.x; hv[1 % TW] = h4.y; hv[2 % TW] = h4.z; hv[3 % TW] = h4.w;
        } else {
#pragma unroll
          for (int t = 0; t < TW; ++t) hv[t] = in[i * TS + t0 + t];
        }
        const float* wrow = W + (size_t)i * dout;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int j = jbase + 32 * k;
          const float wv = j < dout ? (STAGED ? wrow[j] : __ldg(wrow + j)) : 0.f;
#pragma unroll
          for (int t = 0; t < TW; ++t) acc[t][k] = fmaf(hv[t], wv, acc[t][k]);
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int j = jbase + 32 * k;
        if (j < dout) {
          const float bj = MODE == 4 ? 0.f : __ldg(bias + j);
#pragma unroll
          for (int t = 0; t < TW; ++t) {
            const float pre = acc[t][k] + bj;
            if (MODE == 1) zpart[t] += log_cosh(pre);
            else if (MODE == 3) out[j * TS + t0 + t] = tanh_accurate(pre);
            else if (MODE == 4) out[j * TS + t0 + t] = pre * activate_grad(act, aux[j * TS + t0 + t]);
            else out[j * TS + t0 + t] = activate(act, pre);
          }
        }
      }
    }
    if (MODE == 1) {
#pragma unroll
      for (int t = 0; t < TW; ++t) {
        const float s = warp_sum(zpart[t]);
        if (lane == 0) zacc[t0 + t] += s;
      }
    }
  }

  // z[t] (+)= sum_i in[i][t] * v[i] + v0  for the warp's walkers
  __device__ static void dot_head(const float* __restrict__ v, const float* __restrict__ v0,
                                  int din, const float* in, float* z, bool accumulate) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int t0 = warp * TW;
    float part[TW];
#pragma unroll
    for (int t = 0; t < TW; ++t) part[t] = 0.f;
    for (int i = lane; i < din; i += 32) {
      const float vi = __ldg(v + i);
#pragma unroll
      for (int t = 0; t < TW; ++t) part[t] = fmaf(in[i * TS + t0 + t], vi, part[t]);
    }
#pragma unroll
    for (int t = 0; t < TW; ++t) {
      const float s = warp_sum(part[t]) + __ldg(v0);
      if (lane == 0) z[t0 + t] = accumulate ? z[t0 + t] + s : s;
    }
  }

  // Full forward pass of the tile.  smem: buf0, buf1 (act_floats each).
  // cfg: [T][NW] words in shared memory.  z: [T] in shared memory.
  __device__ static void forward(const NetDesc& d, const uint64_t* cfg, float* smem, float* z) {
    float* buf0 = smem;
    float* buf1 = smem + act_floats(d);
    load_spins(d, cfg, buf0);
    __syncthreads();
    const bool rbm = d.kind == CGSVMC_ANSATZ_RBM;
    if (rbm) dot_head(d.rbm_a, d.rbm_a0, d.N, buf0, z, false);   // onsite term
    float* in = buf0;
    float* out = buf1;
    float* wbuf = stage_floats(d) != 0 ? smem + 2 * act_floats(d) : nullptr;
    const bool resident = d.resident_weights != 0;     // copied by prepare()
    int din = d.N;
    for (int l = 0; l < d.L; ++l) {
      if (resident)
        layer<0, true>(d, wbuf + resident_offset(d, l), d.b[l], din, d.D, d.act, in, out, nullptr);
      else if (wbuf != nullptr)
        layer<0, true>(d, stage(d.w[l], din, d.D, wbuf), d.b[l], din, d.D, d.act, in, out, nullptr);
      else
        layer<0>(d, d.w[l], d.b[l], din, d.D, d.act, in, out, nullptr);
      __syncthreads();
      float* tmp = in; in = out; out = tmp;
      din = d.D;
    }
    if (rbm) {
      __syncwarp();
      if (resident)
        layer<1, true>(d, wbuf + resident_offset(d, d.L), d.b[d.L], din, d.D, 0, in, nullptr, z);
      else if (wbuf != nullptr)
        layer<1, true>(d, stage(d.w[d.L], din, d.D, wbuf), d.b[d.L], din, d.D, 0, in, nullptr, z);
      else
        layer<1>(d, d.w[d.L], d.b[d.L], din, d.D, 0, in, nullptr, z);
    } else {
      dot_head(d.w[d.L], d.b[d.L], din, in, z, false);
    }
    __syncthreads();
  }
};

// ===========================================================================
// Convolutional family: conv_1d (Y = 1) and conv_2d, periodic
// ===========================================================================
// Activations: [channel][X*Y] per configuration (channel-major so that the 32
// positions handled by a warp read consecutive addresses).  One thread owns
// one (configuration, position) and all output channels in chunks of 16.
struct Conv {
  __host__ __device__ static size_t act_floats(const NetDesc& d, int T) {
    return (size_t)T * d.C * d.N;
  }
  __host__ static size_t forward_smem_bytes(const NetDesc& d, int T) {
    return (2 * act_floats(d, T) + T) * sizeof(float) + (size_t)(d.X * d.kx + d.Y * d.ky) * sizeof(int);
  }

  // wrap tables: xi[x * kx + dx] = ((x + dx - pad_x) mod X) * Y ; yi similarly
  __device__ static void build_tables(const NetDesc& d, int* xi, int* yi, bool backward) {
    const int px = backward ? d.kx - 1 - d.pad_x : d.pad_x;
    const int py = backward ? d.ky - 1 - d.pad_y : d.pad_y;
    for (int e = threadIdx.x; e < d.X * d.kx; e += kThreads) {
      const int x = e / d.kx, dx = e - x * d.kx;
      xi[e] = (((x + dx - px) % d.X + d.X) % d.X) * d.Y;
    }
    for (int e = threadIdx.x; e < d.Y * d.ky; e += kThreads) {
      const int y = e / d.ky, dy = e - y * d.ky;
      yi[e] = ((y + dy - py) % d.Y + d.Y) % d.Y;
    }
  }

  // One convolution layer over T configurations.
  //   out[t][co][pos] = post(bias[co] + sum_{tap,ci} in[t][ci][wrap(pos + tap)] * W[tap][ci][co])
  // W layout [kx*ky][cin][cout] (snt.Conv2D w:[kh,kw,in,out]).
  // MODE 0: post = act (or identity when last).  MODE 1 (last layer of the
  // forward pass): no store, z[t] += sum of outputs.  MODE 2: backward-data,
  // flipped taps (weights must be the [tap][cout][cin] transpose), post =
  // multiply with activate_grad(h_prev) read from `aux`.
  // RES (MODE 0 and 2): `res`, a tensor of the output's shape, is added to the
  // result -- the skip connection of a residual block (may alias `out`: every
  // element is read and written by the same thread).  ACTGRAD = false: MODE 2
  // without the activation-gradient factor.
  // RES / ACTGRAD are compile-time so that the plain conv ansaetze keep their
  // original inner loops.
  template <int MODE, bool RES = false, bool ACTGRAD = true>
  __device__ static void layer(const NetDesc& d, const float* __restrict__ W,
                               const float* __restrict__ bias, int cin, int cout, int act,
                               bool apply_act, int T, const float* in, float* out, float* z,
                               const int* xi, const int* yi, const float* aux,
                               const float* res = nullptr) {
    const int npos = d.N;
    const int taps = d.kx * d.ky;
    for (int item = threadIdx.x; item < T * npos; item += kThreads) {
      const int t = item / npos, pos = item - t * npos;
      const int x = pos / d.Y, y = pos - x * d.Y;
      const float* in_t = in + (size_t)t * cin * npos;
      float zsum = 0.f;
      for (int c0 = 0; c0 < cout; c0 += 16) {
        float acc[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) acc[c] = 0.f;
        for (int tap = 0; tap < taps; ++tap) {
          const int dx = tap / d.ky, dy = tap - dx * d.ky;
          // backward-data walks the taps mirrored: in[pos - tap + pad]
          const int sx = MODE == 2 ? xi[x * d.kx + (d.kx - 1 - dx)] : xi[x * d.kx + dx];
          const int sy = MODE == 2 ? yi[y * d.ky + (d.ky - 1 - dy)] : yi[y * d.ky + dy];
          const float* src = in_t + sx + sy;
          const float* wtap = W + ((size_t)tap * cin) * cout + c0;
          for (int ci = 0; ci < cin; ++ci) {
            const float v = src[(size_t)ci * npos];
            const float* wp = wtap + (size_t)ci * cout;
            if (cout - c0 >= 16 && (cout & 3) == 0) {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float4 w4 = __ldg(reinterpret_cast<const float4*>(wp) + q);
                acc[4 * q + 0] = fmaf(v, w4.x, acc[4 * q + 0]);
                acc[4 * q + 1] = fmaf(v, w4.y, acc[4 * q + 1]);
                acc[4 * q + 2] = fmaf(v, w4.z, acc[4 * q + 2]);
                acc[4 * q + 3] = fmaf(v, w4.w, acc[4 * q + 3]);
              }
            } else {
#pragma unroll
              for (int c = 0; c < 16; ++c)
                if (c0 + c < cout) acc[c] = fmaf(v, __ldg(wp + c), acc[c]);
            }
          }
        }
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          if (c0 + c < cout) {
            float v = acc[c];
            if (MODE != 2) v += __ldg(bias + c0 + c);
            if (MODE == 1) {
              zsum += v;
            } else if (MODE == 2) {
              const size_t o = ((size_t)t * cout + c0 + c) * npos + pos;
              if (ACTGRAD) v *= activate_grad(act, aux[o]);
              out[o] = RES ? v + res[o] : v;
            } else {
              const size_t o = ((size_t)t * cout + c0 + c) * npos + pos;
              if (apply_act) v = activate(act, v);
              out[o] = RES ? v + res[o] : v;
            }
          }
        }
      }
      if (MODE == 1) out[item] = zsum;   // reduced in fixed order by reduce_z()
    }
  }

  // z[t] = sum over positions of the per-position sums left in `part` by the
  // last layer (fixed order: deterministic amplitudes).
  __device__ static void reduce_z(const NetDesc& d, int T, const float* part, float* z) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int t = warp; t < T; t += kWarps) {
      float s = 0.f;
      for (int p = lane; p < d.N; p += 32) s += part[t * d.N + p];
      s = warp_sum(s);
      if (lane == 0) z[t] = s;
    }
  }

  __device__ static void load_spins(const NetDesc& d, const uint64_t* cfg, int T, float* buf) {
    for (int e = threadIdx.x; e < T * d.N; e += kThreads) {
      const int t = e / d.N, i = e - t * d.N;
      buf[e] = word_bit(cfg + t * d.NW, i) ? 1.f : -1.f;
    }
  }

  __device__ static void forward(const NetDesc& d, const uint64_t* cfg, int T, float* smem,
                                 float* z) {
    float* buf0 = smem;
    float* buf1 = smem + act_floats(d, T);
    int* xi = reinterpret_cast<int*>(smem + 2 * act_floats(d, T));
    int* yi = xi + d.X * d.kx;
    build_tables(d, xi, yi, false);
    load_spins(d, cfg, T, buf0);
    __syncthreads();
    if (d.resnet) {
      // x = conv0(sigma) in buf1; per block h = selu(conv1(x)) in buf0, then
      // x += conv2(h) in place (wavefunctions.py:651-671, layers.py:203-228)
      float* x = buf1;
      float* h = buf0;
      layer<0>(d, d.w[0], d.b[0], 1, d.C, d.act, false, T, buf0, x, z, xi, yi, nullptr);
      __syncthreads();
      for (int l = 1; l < d.L; l += 2) {
        layer<0>(d, d.w[l], d.b[l], d.C, d.C, d.act, true, T, x, h, z, xi, yi, nullptr);
        __syncthreads();
        layer<0, true>(d, d.w[l + 1], d.b[l + 1], d.C, d.C, d.act, false, T, h, x, z, xi, yi, nullptr, x);
        __syncthreads();
      }
      // per-position channel sums (into h, free now), then the fixed-order reduction
      for (int item = threadIdx.x; item < T * d.N; item += kThreads) {
        const int t = item / d.N, pos = item - t * d.N;
        float sum = 0.f;
        for (int c = 0; c < d.C; ++c) sum += x[((size_t)t * d.C + c) * d.N + pos];
        h[item] = sum;
      }
      __syncthreads();
      reduce_z(d, T, h, z);
      __syncthreads();
      return;
    }
    float* in = buf0;
    float* out = buf1;
    int cin = 1;
    for (int l = 0; l < d.L; ++l) {
      if (l + 1 < d.L) layer<0>(d, d.w[l], d.b[l], cin, d.C, d.act, true, T, in, out, z, xi, yi, nullptr);
      else layer<1>(d, d.w[l], d.b[l], cin, d.C, d.act, false, T, in, out, z, xi, yi, nullptr);
      __syncthreads();
      float* tmp = in; in = out; out = tmp;
      cin = d.C;
    }
    reduce_z(d, T, in, z);     // `in` is the buffer the last layer wrote
    __syncthreads();
  }
};

// ---------------------------------------------------------------------------
// Uniform front end over the two families
// ---------------------------------------------------------------------------
template <int TW>
struct MlpNet {
  __host__ __device__ static int tile(const NetDesc&, int) { return Mlp<TW>::T; }
  __host__ static size_t fwd_smem(const NetDesc& d, int) { return Mlp<TW>::forward_smem_bytes(d); }
  __device__ static void forward(const NetDesc& d, const uint64_t* cfg, int, float* smem, float* z) {
    Mlp<TW>::forward(d, cfg, smem, z);
  }
  __device__ static void prepare(const NetDesc& d, float* smem) { Mlp<TW>::prepare(d, smem); }
  static constexpr int kMinCtas = 0;  // residency left to ptxas
};
struct ConvNet {
  __host__ static size_t fwd_smem(const NetDesc& d, int T) { return Conv::forward_smem_bytes(d, T); }
  __device__ static void forward(const NetDesc& d, const uint64_t* cfg, int T, float* smem, float* z) {
    Conv::forward(d, cfg, T, smem, z);
  }
  __device__ static void prepare(const NetDesc&, float*) {}
  static constexpr int kMinCtas = 2;  // conv_tile() budgets 96 KB of shared memory per CTA
};

// Dynamic shared memory layout of the non-grad kernels:
//   [ forward scratch (fwd_smem) | cfg words T*NW u64 | z tile T f32 | kernel extras ]
// Bump allocator over the dynamic shared memory behind the forward scratch;
// every block is 16-byte aligned (host sizes include the slack).
struct Carver {
  char* p;
  __device__ Carver(float* smem, size_t fwd_bytes) : p(reinterpret_cast<char*>(smem) + fwd_bytes) {}
  template <typename Tp>
  __device__ Tp* take(size_t n) {
    Tp* r = reinterpret_cast<Tp*>(p);
    p += (n * sizeof(Tp) + 15) / 16 * 16;
    return r;
  }
};
__host__ inline size_t carve_bytes(size_t n, size_t elem) { return (n * elem + 15) / 16 * 16; }

// ---------------------------------------------------------------------------
// K1: log-amplitudes of consecutive walkers
// ---------------------------------------------------------------------------
template <class NET>
__global__ void __launch_bounds__(kThreads, NET::kMinCtas)
net_log_amp_kernel(NetDesc d, int T, size_t fwd_bytes, const uint64_t* __restrict__ packed,
                   int64_t B, float* __restrict__ out) {
  extern __shared__ __align__(16) float smem[];
  NET::prepare(d, smem);
  Carver carve(smem, fwd_bytes);
  uint64_t* cfg = carve.take<uint64_t>((size_t)T * d.NW);
  float* z = carve.take<float>(T);
  for (int64_t b0 = (int64_t)blockIdx.x * T; b0 < B; b0 += (int64_t)gridDim.x * T) {
    for (int e = threadIdx.x; e < T * d.NW; e += kThreads) {
      const int64_t b = b0 + e / d.NW;
      cfg[e] = b < B ? packed[b * d.NW + e % d.NW] : 0ull;
    }
    __syncthreads();
    NET::forward(d, cfg, T, smem, z);
    for (int t = threadIdx.x; t < T; t += kThreads)
      if (b0 + t < B) out[b0 + t] = z[t];
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------
// K2: Metropolis sampler.  A CTA owns T walkers for all steps.
// ---------------------------------------------------------------------------
__device__ __forceinline__ int kth_set_bit_serial(const uint64_t* words, int nw, int k) {
  for (int w = 0; w < nw; ++w) {
    uint64_t m = words[w];
    const int c = __popcll(m);
    if (k >= c) { k -= c; continue; }
    for (int q = 0; q < k; ++q) m &= m - 1;     // drop k lowest set bits
    return w * 64 + __ffsll((long long)m) - 1;
  }
  return 0;
}

template <class NET>
__global__ void __launch_bounds__(kThreads, NET::kMinCtas)
net_mc_kernel(NetDesc d, int T, size_t fwd_bytes, uint64_t* __restrict__ packed, int64_t B,
              int n_steps, uint64_t seed, uint64_t walker0, uint64_t step0,
              const uint64_t* __restrict__ step0_dev, unsigned long long* accept_count,
              float* __restrict__ log_amp_out) {
  if (step0_dev != nullptr) step0 += *step0_dev;
  extern __shared__ __align__(16) float smem[];
  NET::prepare(d, smem);
  Carver carve(smem, fwd_bytes);
  uint64_t* prop = carve.take<uint64_t>((size_t)T * d.NW);   // proposed configs [T][NW]
  uint64_t* cur = carve.take<uint64_t>((size_t)T * d.NW);    // current configs [T][NW]
  float* z_new = carve.take<float>(T);
  float* z_cur = carve.take<float>(T);
  float* u_acc = carve.take<float>(T);
  __shared__ unsigned int n_acc_s;
  if (threadIdx.x == 0) n_acc_s = 0;
  for (int64_t b0 = (int64_t)blockIdx.x * T; b0 < B; b0 += (int64_t)gridDim.x * T) {
    for (int e = threadIdx.x; e < T * d.NW; e += kThreads) {
      const int64_t b = b0 + e / d.NW;
      const uint64_t wv = b < B ? packed[b * d.NW + e % d.NW] : 0ull;
      cur[e] = wv;
      prop[e] = wv;
    }
    __syncthreads();
    NET::forward(d, prop, T, smem, z_new);
    for (int t = threadIdx.x; t < T; t += kThreads) z_cur[t] = z_new[t];
    __syncthreads();
    for (int step = 0; step < n_steps; ++step) {
      // propose: one thread per walker (uniformly random up and down site)
      for (int t = threadIdx.x; t < T; t += kThreads) {
        const int64_t b = b0 + t;
        uint64_t s[NWMAX], dn[NWMAX];
        int n_up = 0;
        for (int w = 0; w < d.NW; ++w) {
          s[w] = cur[t * d.NW + w];
          dn[w] = ~s[w] & valid_mask_word(d.N, w);
          n_up += __popcll(s[w]);
        }
        const int n_dn = d.N - n_up;
        float u = 2.f;     // > any probability: padding walkers never accept
        if (b < B && n_up > 0 && n_dn > 0) {
          const Philox4 r = walker_step_random(seed, walker0 + (uint64_t)b, step0 + (uint64_t)step);
          const int up = kth_set_bit_serial(s, d.NW, (int)__umulhi(r.x, (uint32_t)n_up));
          const int dnsite = kth_set_bit_serial(dn, d.NW, (int)__umulhi(r.y, (uint32_t)n_dn));
          s[up >> 6] ^= 1ull << (up & 63);
          s[dnsite >> 6] ^= 1ull << (dnsite & 63);
          u = u32_to_unit(r.z);
        }
        for (int w = 0; w < d.NW; ++w) prop[t * d.NW + w] = s[w];
        u_acc[t] = u;
      }
      __syncthreads();
      NET::forward(d, prop, T, smem, z_new);
      // accept iff |psi'/psi|^2 > u  (graph_builders.py:75-79 squared; strict)
      for (int t = threadIdx.x; t < T; t += kThreads) {
        const float prob = fast_exp(2.f * (z_new[t] - z_cur[t]));
        if (prob > u_acc[t]) {
          for (int w = 0; w < d.NW; ++w) cur[t * d.NW + w] = prop[t * d.NW + w];
          z_cur[t] = z_new[t];
          atomicAdd(&n_acc_s, 1u);
        }
      }
      __syncthreads();
    }
    for (int e = threadIdx.x; e < T * d.NW; e += kThreads) {
      const int64_t b = b0 + e / d.NW;
      if (b < B) packed[b * d.NW + e % d.NW] = cur[e];
    }
    if (log_amp_out != nullptr)
      for (int t = threadIdx.x; t < T; t += kThreads)
        if (b0 + t < B) log_amp_out[b0 + t] = z_cur[t];
    __syncthreads();
  }
  if (threadIdx.x == 0 && accept_count != nullptr && n_acc_s)
    atomicAdd(accept_count, (unsigned long long)n_acc_s);
}

// Replay step: uniforms supplied by the caller (graph_builders.py:59-79).
template <class NET>
__global__ void __launch_bounds__(kThreads, NET::kMinCtas)
net_replay_kernel(NetDesc d, int T, size_t fwd_bytes, uint64_t* __restrict__ packed, int64_t B,
                  const float* __restrict__ u_sites, const float* __restrict__ u_acc,
                  int32_t* down_out, int32_t* up_out, float* log_ratio_out, uint8_t* accept_out) {
  extern __shared__ __align__(16) float smem[];
  NET::prepare(d, smem);
  Carver carve(smem, fwd_bytes);
  uint64_t* prop = carve.take<uint64_t>((size_t)T * d.NW);
  uint64_t* cur = carve.take<uint64_t>((size_t)T * d.NW);
  float* z_new = carve.take<float>(T);
  float* z_cur = carve.take<float>(T);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t b0 = (int64_t)blockIdx.x * T; b0 < B; b0 += (int64_t)gridDim.x * T) {
    for (int e = threadIdx.x; e < T * d.NW; e += kThreads) {
      const int64_t b = b0 + e / d.NW;
      const uint64_t wv = b < B ? packed[b * d.NW + e % d.NW] : 0ull;
      cur[e] = wv;
      prop[e] = wv;
    }
    __syncthreads();
    NET::forward(d, prop, T, smem, z_new);
    for (int t = threadIdx.x; t < T; t += kThreads) z_cur[t] = z_new[t];
    __syncthreads();
    // argmin / argmax of sigma * u with first-occurrence ties: a warp per walker
    for (int t = warp; t < T; t += kWarps) {
      const int64_t b = b0 + t;
      if (b >= B) continue;
      float vmin = INFINITY, vmax = -INFINITY;
      int imin = d.N, imax = d.N;
      for (int i = lane; i < d.N; i += 32) {
        const float u = u_sites[b * d.N + i];
        const float v = word_bit(cur + t * d.NW, i) ? u : -u;
        if (v < vmin) { vmin = v; imin = i; }
        if (v > vmax) { vmax = v; imax = i; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(CGSVMC_FULL_MASK, vmin, o);
        const int oi = __shfl_xor_sync(CGSVMC_FULL_MASK, imin, o);
        if (ov < vmin || (ov == vmin && oi < imin)) { vmin = ov; imin = oi; }
        const float ow = __shfl_xor_sync(CGSVMC_FULL_MASK, vmax, o);
        const int oj = __shfl_xor_sync(CGSVMC_FULL_MASK, imax, o);
        if (ow > vmax || (ow == vmax && oj < imax)) { vmax = ow; imax = oj; }
      }
      if (lane == 0) {
        prop[t * d.NW + (imin >> 6)] ^= 1ull << (imin & 63);   // +2 at the down site
        prop[t * d.NW + (imax >> 6)] ^= 1ull << (imax & 63);   // -2 at the up site
        if (down_out) down_out[b] = imin;
        if (up_out) up_out[b] = imax;
      }
    }
    __syncthreads();
    NET::forward(d, prop, T, smem, z_new);
    for (int t = threadIdx.x; t < T; t += kThreads) {
      const int64_t b = b0 + t;
      if (b >= B) continue;
      const float dl = z_new[t] - z_cur[t];
      const bool acc = expf(dl) > sqrtf(u_acc[b]);
      if (log_ratio_out) log_ratio_out[b] = dl;
      if (accept_out) accept_out[b] = acc ? 1 : 0;
      if (acc)
        for (int w = 0; w < d.NW; ++w) packed[b * d.NW + w] = prop[t * d.NW + w];
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------
// K3: local energy.  A CTA takes WG walkers, lists (walker, active bond)
// items plus the WG base configurations, evaluates the list tile by tile.
// ---------------------------------------------------------------------------
constexpr int kElocWalkers = 8;

template <class NET>
__global__ void __launch_bounds__(kThreads, NET::kMinCtas)
net_eloc_kernel(NetDesc d, int T, size_t fwd_bytes, const int2* __restrict__ bonds_ij,
                const float* __restrict__ bonds_jx, const float* __restrict__ bonds_jz,
                int n_bonds, const uint64_t* __restrict__ packed, int64_t B,
                float* __restrict__ e_loc, float* __restrict__ log_amp_out,
                float* __restrict__ diag_out, float* __restrict__ off_out) {
  extern __shared__ __align__(16) float smem[];
  NET::prepare(d, smem);
  const int max_items = kElocWalkers * (n_bonds + 1);
  Carver carve(smem, fwd_bytes);
  uint64_t* cfg = carve.take<uint64_t>((size_t)T * d.NW);               // tile configs [T][NW]
  uint64_t* base = carve.take<uint64_t>((size_t)kElocWalkers * d.NW);   // walkers of the group
  float* z_tile = carve.take<float>(T);
  float* z_item = carve.take<float>(max_items);
  int* item_code = carve.take<int>(max_items);                          // walker << 16 | bond + 1
  float* diag_s = carve.take<float>(kElocWalkers);
  __shared__ int n_items_s;
  __shared__ int count_s[kElocWalkers], start_s[kElocWalkers];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t g0 = (int64_t)blockIdx.x * kElocWalkers; g0 < B; g0 += (int64_t)gridDim.x * kElocWalkers) {
    const int wg = (int)min((int64_t)kElocWalkers, B - g0);
    // items 0..wg-1 are the base configurations
    for (int e = threadIdx.x; e < wg * d.NW; e += kThreads) base[e] = packed[g0 * d.NW + e];
    if (threadIdx.x < wg) item_code[threadIdx.x] = threadIdx.x << 16;
    __syncthreads();
    // enumerate antiparallel bonds: a warp per walker; two passes (count, then
    // ballot compaction at deterministic offsets) so the evaluation order and
    // therefore the float32 sums do not depend on warp scheduling
    for (int wl = warp; wl < wg; wl += kWarps) {
      float diag = 0.f;
      int count = 0;
      for (int k0 = 0; k0 < n_bonds; k0 += 32) {
        const int k = k0 + lane;
        bool anti = false;
        if (k < n_bonds) {
          const int2 bd = bonds_ij[k];
          anti = word_bit(base + wl * d.NW, bd.x) != word_bit(base + wl * d.NW, bd.y);
          diag += (anti ? -0.25f : 0.25f) * bonds_jz[k];               // operators.py:165,169
        }
        count += __popc(__ballot_sync(CGSVMC_FULL_MASK, anti));
      }
      diag = warp_sum(diag);
      if (lane == 0) { diag_s[wl] = diag; count_s[wl] = count; }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int run = wg;
      for (int wl = 0; wl < wg; ++wl) { start_s[wl] = run; run += count_s[wl]; }
      n_items_s = run;
    }
    __syncthreads();
    for (int wl = warp; wl < wg; wl += kWarps) {
      int slot = start_s[wl];
      for (int k0 = 0; k0 < n_bonds; k0 += 32) {
        const int k = k0 + lane;
        bool anti = false;
        if (k < n_bonds) {
          const int2 bd = bonds_ij[k];
          anti = word_bit(base + wl * d.NW, bd.x) != word_bit(base + wl * d.NW, bd.y);
        }
        const uint32_t vote = __ballot_sync(CGSVMC_FULL_MASK, anti);
        if (anti) item_code[slot + __popc(vote & ((1u << lane) - 1u))] = (wl << 16) | (k + 1);
        slot += __popc(vote);
      }
    }
    __syncthreads();
    const int n_items = n_items_s;
    for (int i0 = 0; i0 < n_items; i0 += T) {
      for (int e = threadIdx.x; e < T * d.NW; e += kThreads) {
        const int t = e / d.NW, w = e - t * d.NW;
        const int it = min(i0 + t, n_items - 1);           // padding repeats the last item
        const int code = item_code[it];
        const int wl = code >> 16, kb = (code & 0xffff) - 1;
        uint64_t word = base[wl * d.NW + w];
        if (kb >= 0) {
          const int2 bd = bonds_ij[kb];
          if ((bd.x >> 6) == w) word ^= 1ull << (bd.x & 63);           // operators.py:158-164
          if ((bd.y >> 6) == w) word ^= 1ull << (bd.y & 63);
        }
        cfg[e] = word;
      }
      __syncthreads();
      NET::forward(d, cfg, T, smem, z_tile);
      for (int t = threadIdx.x; t < T; t += kThreads)
        if (i0 + t < n_items) z_item[i0 + t] = z_tile[t];
      __syncthreads();
    }
    // E_loc = diag + sum_active jx/2 exp(z' - z)   (operators.py:168-169, 259)
    for (int wl = warp; wl < wg; wl += kWarps) {
      const float z0 = z_item[wl];
      float off = 0.f;
      for (int it = start_s[wl] + lane; it < start_s[wl] + count_s[wl]; it += 32)
        off = fmaf(0.5f * bonds_jx[(item_code[it] & 0xffff) - 1], fast_exp(z_item[it] - z0), off);
      off = warp_sum(off);
      if (lane == 0) {
        const int64_t b = g0 + wl;
        e_loc[b] = diag_s[wl] + off;
        if (log_amp_out) log_amp_out[b] = z0;
        if (diag_out) diag_out[b] = diag_s[wl];
        if (off_out) off_out[b] = off;
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------
// K4: weighted sums of O_b = d z_b / d params (training.py:545-548, 169-175).
// Forward (activations kept in shared memory) and backward per tile; every
// thread owns kGradE entries of the flat parameter vector of its chunk
// (blockIdx.y) and accumulates them over the CTA's walkers; partial sums go to
// scratch and are reduced in a fixed order (deterministic).
// ---------------------------------------------------------------------------
constexpr int kGradE = 16;

__global__ void transpose_batched_kernel(const float* __restrict__ src, int batch, int rows,
                                         int cols, float* __restrict__ dst) {
  const int64_t total = (int64_t)batch * rows * cols;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t bi = e / ((int64_t)rows * cols);
    const int r = (int)((e / cols) % rows), c = (int)(e % cols);
    dst[(bi * cols + c) * rows + r] = src[e];
  }
}

template <int TW, int K>
__global__ void __launch_bounds__(kThreads)
mlp_grad_kernel(NetDesc d, const uint64_t* __restrict__ packed, const float* __restrict__ weights,
                int64_t B, int64_t walkers_per_cta, int64_t P, float* __restrict__ partials) {
  using M = Mlp<TW>;
  constexpr int T = M::T, TS = M::TS;
  extern __shared__ __align__(16) float smem[];
  const bool rbm = d.kind == CGSVMC_ANSATZ_RBM;
  const int N = d.N, L = d.L, D = d.D;
  // float offsets of the row blocks inside the dynamic shared memory
  const int offH0 = 0;
  const int offHl = offH0 + N * TS;            // H[1..L]
  const int offDL = offHl + L * D * TS;        // delta of hidden layer 0..L-1
  const int offDH = offDL + L * D * TS;        // rbm: tanh(theta)
  const int offOnes = offDH + (rbm ? D * TS : 0);
  const int offW = offOnes + TS;               // K weight rows
  const int offEnd = offW + K * TS;
  uint64_t* cfg = reinterpret_cast<uint64_t*>(smem + (offEnd + 3) / 4 * 4);
  // one layer's weight matrix staged in shared memory (see Mlp::stage): the
  // layers of the forward AND backward pass would otherwise be streamed
  // through L1 by every warp (ncu: long-scoreboard 7.6 per issue, issue-active 15 %)
  float* wbuf = d.stage_weights
      ? reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(cfg + (size_t)T * d.NW) + 15) / 16 * 16)
      : nullptr;
  auto Hrow = [&](int l, int i) { return l == 0 ? offH0 + i * TS : offHl + ((l - 1) * D + i) * TS; };

  // decode this thread's entries into (A row, B row) shared-memory offsets
  int ea[kGradE], eb[kGradE];
  const int64_t e0 = (int64_t)blockIdx.y * (kThreads * kGradE) + threadIdx.x;
#pragma unroll
  for (int m = 0; m < kGradE; ++m) {
    const int64_t e = e0 + (int64_t)m * kThreads;
    int a = offOnes, b = offOnes;
    if (e < P) {
      bool done = false;
      if (rbm) {
        if (e < N) { a = Hrow(0, (int)e); done = true; }
        else if (e == N) { done = true; }
      }
      int din = N;
      for (int l = 0; l < L && !done; ++l) {
        if (e >= d.w_off[l] && e < d.w_off[l] + (int64_t)din * D) {
          const int q = (int)(e - d.w_off[l]);
          a = Hrow(l, q / D); b = offDL + (l * D + q % D) * TS; done = true;
        } else if (e >= d.b_off[l] && e < d.b_off[l] + D) {
          b = offDL + (l * D + (int)(e - d.b_off[l])) * TS; done = true;
        }
        din = D;
      }
      if (!done) {
        const int dinL = L == 0 ? N : D;
        if (rbm) {
          if (e >= d.w_off[L] && e < d.w_off[L] + (int64_t)dinL * D) {
            const int q = (int)(e - d.w_off[L]);
            a = Hrow(L, q / D); b = offDH + (q % D) * TS;
          } else if (e >= d.b_off[L]) {
            b = offDH + (int)(e - d.b_off[L]) * TS;
          }
        } else if (e >= d.w_off[L] && e < d.w_off[L] + dinL) {
          a = Hrow(L, (int)(e - d.w_off[L]));
        }
      }
    }
    ea[m] = a; eb[m] = b;
  }
  float acc[K][kGradE];
#pragma unroll
  for (int k = 0; k < K; ++k)
#pragma unroll
    for (int m = 0; m < kGradE; ++m) acc[k][m] = 0.f;

  for (int t = threadIdx.x; t < TS; t += kThreads) smem[offOnes + t] = 1.f;
  const int64_t b_begin = (int64_t)blockIdx.x * walkers_per_cta;
  const int64_t b_end = min(B, b_begin + walkers_per_cta);
  for (int64_t b0 = b_begin; b0 < b_end; b0 += T) {
    for (int e = threadIdx.x; e < T * d.NW; e += kThreads) {
      const int64_t b = b0 + e / d.NW;
      cfg[e] = b < b_end ? packed[b * d.NW + e % d.NW] : 0ull;
    }
    for (int e = threadIdx.x; e < K * T; e += kThreads) {
      const int k = e / T, t = e - k * T;
      smem[offW + k * TS + t] = b0 + t < b_end ? weights[(int64_t)k * B + b0 + t] : 0.f;
    }
    __syncthreads();
    M::load_spins(d, cfg, smem + offH0);
    __syncthreads();
    // forward, keeping every layer's output
    int din = N;
    for (int l = 0; l < L; ++l) {
      if (wbuf != nullptr)
        M::template layer<0, true>(d, M::stage(d.w[l], din, D, wbuf), d.b[l], din, D, d.act,
                                   smem + Hrow(l, 0), smem + Hrow(l + 1, 0), nullptr);
      else
        M::template layer<0>(d, d.w[l], d.b[l], din, D, d.act, smem + Hrow(l, 0), smem + Hrow(l + 1, 0), nullptr);
      __syncthreads();
      din = D;
    }
    if (rbm) {
      if (wbuf != nullptr)
        M::template layer<3, true>(d, M::stage(d.w[L], din, D, wbuf), d.b[L], din, D, 0, smem + Hrow(L, 0),
                                   smem + offDH, nullptr);
      else
        M::template layer<3>(d, d.w[L], d.b[L], din, D, 0, smem + Hrow(L, 0), smem + offDH, nullptr);
      __syncthreads();
    }
    // backward: delta of the top hidden layer, then down to hidden layer 0
    if (L >= 1) {
      float* top = smem + offDL + (L - 1) * D * TS;
      if (rbm) {
        if (wbuf != nullptr)
          M::template layer<4, true>(d, M::stage(d.wt[L], D, D, wbuf), nullptr, D, D, d.act, smem + offDH, top,
                                     nullptr, smem + Hrow(L, 0));
        else
          M::template layer<4>(d, d.wt[L], nullptr, D, D, d.act, smem + offDH, top, nullptr, smem + Hrow(L, 0));
      } else {
        for (int e = threadIdx.x; e < D * T; e += kThreads) {
          const int i = e / T, t = e - i * T;
          top[i * TS + t] = __ldg(d.w[L] + i) * activate_grad(d.act, smem[Hrow(L, i) + t]);
        }
      }
      __syncthreads();
      for (int l = L - 1; l >= 1; --l) {
        if (wbuf != nullptr)
          M::template layer<4, true>(d, M::stage(d.wt[l], D, D, wbuf), nullptr, D, D, d.act,
                                     smem + offDL + l * D * TS, smem + offDL + (l - 1) * D * TS, nullptr,
                                     smem + Hrow(l, 0));
        else
          M::template layer<4>(d, d.wt[l], nullptr, D, D, d.act, smem + offDL + l * D * TS,
                               smem + offDL + (l - 1) * D * TS, nullptr, smem + Hrow(l, 0));
        __syncthreads();
      }
    }
    // accumulate: acc[k][entry] += sum_t w_k[t] * A[t] * B[t]
#pragma unroll
    for (int m = 0; m < kGradE; ++m) {
      const float* A = smem + ea[m];
      const float* Bv = smem + eb[m];
#pragma unroll
      for (int t = 0; t < T; t += 4) {
        const float4 a4 = *reinterpret_cast<const float4*>(A + t);
        const float4 b4 = *reinterpret_cast<const float4*>(Bv + t);
        const float p0 = a4.x * b4.x, p1 = a4.y * b4.y, p2 = a4.z * b4.z, p3 = a4.w * b4.w;
#pragma unroll
        for (int k = 0; k < K; ++k) {
          const float4 w4 = *reinterpret_cast<const float4*>(smem + offW + k * TS + t);
          acc[k][m] = fmaf(w4.x, p0, acc[k][m]);
          acc[k][m] = fmaf(w4.y, p1, acc[k][m]);
          acc[k][m] = fmaf(w4.z, p2, acc[k][m]);
          acc[k][m] = fmaf(w4.w, p3, acc[k][m]);
        }
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int m = 0; m < kGradE; ++m) {
    const int64_t e = e0 + (int64_t)m * kThreads;
    if (e < P) {
#pragma unroll
      for (int k = 0; k < K; ++k) partials[((int64_t)blockIdx.x * K + k) * P + e] = acc[k][m];
    }
  }
}

template <int K>
__global__ void __launch_bounds__(kThreads, 1)  // one CTA per SM (shared memory): no register cap, no spills
conv_grad_kernel(NetDesc d, int T, const uint64_t* __restrict__ packed,
                 const float* __restrict__ weights, int64_t B, int64_t walkers_per_cta, int64_t P,
                 float* __restrict__ partials) {
  extern __shared__ __align__(16) float smem[];
  const int N = d.N, L = d.L, C = d.C;
  const size_t szH0 = (size_t)T * N, szC = (size_t)T * C * N;
  float* H0 = smem;                        // [T][1][N]
  float* Hl = H0 + szH0;                   // H[1..L-1]: [T][C][N] each
  float* DL = Hl + (size_t)(L - 1) * szC;  // delta of layer 0..L-1: [T][C][N] each
  float* wk = DL + (size_t)L * szC;        // [K][T]
  int* xi = reinterpret_cast<int*>(wk + K * T);
  int* yi = xi + d.X * d.kx;
  int* xib = yi + d.Y * d.ky;
  int* yib = xib + d.X * d.kx;
  uint64_t* cfg = reinterpret_cast<uint64_t*>(
      (reinterpret_cast<uintptr_t>(yib + d.Y * d.ky) + 15) / 16 * 16);
  auto H = [&](int l) { return l == 0 ? H0 : Hl + (size_t)(l - 1) * szC; };
  Conv::build_tables(d, xi, yi, false);
  Conv::build_tables(d, xib, yib, true);

  // every CTA owns a slice [K][P] of `partials` (L2 resident) and adds the
  // contribution of each of its walker tiles to it; the forward / backward pass
  // of a tile is done once for all parameter entries
  float* part = partials + (int64_t)blockIdx.x * K * P;
  bool first_tile = true;

  const int64_t b_begin = (int64_t)blockIdx.x * walkers_per_cta;
  const int64_t b_end = min(B, b_begin + walkers_per_cta);
  for (int64_t b0 = b_begin; b0 < b_end; b0 += T) {
    for (int e = threadIdx.x; e < T * d.NW; e += kThreads) {
      const int64_t b = b0 + e / d.NW;
      cfg[e] = b < b_end ? packed[b * d.NW + e % d.NW] : 0ull;
    }
    for (int e = threadIdx.x; e < K * T; e += kThreads) {
      const int k = e / T, t = e - k * T;
      wk[e] = b0 + t < b_end ? weights[(int64_t)k * B + b0 + t] : 0.f;
    }
    __syncthreads();
    Conv::load_spins(d, cfg, T, H0);
    __syncthreads();
    float* top = DL + (size_t)(L - 1) * szC;
    if (d.resnet) {
      // H(l) = input of conv l: H(1) = x_0 = conv0(sigma); block b (convs 2b+1,
      // 2b+2): H(2b+2) = h_b = selu(conv(x_b)), H(2b+3) = x_{b+1} = x_b + conv(h_b)
      // (the last x is not needed: z is its sum, so its delta is 1).
      Conv::layer<0>(d, d.w[0], d.b[0], 1, C, d.act, false, T, H(0), H(1), nullptr, xi, yi, nullptr);
      __syncthreads();
      for (int l = 1; l < L; l += 2) {
        Conv::layer<0>(d, d.w[l], d.b[l], C, C, d.act, true, T, H(l), H(l + 1), nullptr, xi, yi, nullptr);
        __syncthreads();
        if (l + 2 < L) {
          Conv::layer<0, true>(d, d.w[l + 1], d.b[l + 1], C, C, d.act, false, T, H(l + 1), H(l + 2), nullptr,
                               xi, yi, nullptr, H(l));
          __syncthreads();
        }
      }
      // DL[l] = delta at the output of conv l: DL[2b+2] = delta x_{b+1},
      // DL[2b+1] = conv_{2b+2}^T(DL[2b+2]) * selu'(h_b), DL[2b] = DL[2b+2] + conv_{2b+1}^T(DL[2b+1])
      for (size_t e = threadIdx.x; e < szC; e += kThreads) top[e] = 1.f;
      __syncthreads();
      for (int l = L - 1; l >= 2; l -= 2) {
        Conv::layer<2>(d, d.wt[l], nullptr, C, C, d.act, false, T, DL + (size_t)l * szC,
                       DL + (size_t)(l - 1) * szC, nullptr, xib, yib, H(l));
        __syncthreads();
        Conv::layer<2, true, false>(d, d.wt[l - 1], nullptr, C, C, d.act, false, T, DL + (size_t)(l - 1) * szC,
                                    DL + (size_t)(l - 2) * szC, nullptr, xib, yib, nullptr, DL + (size_t)l * szC);
        __syncthreads();
      }
    } else {
    // forward up to the input of the last layer (its output is not needed)
    int cin = 1;
    for (int l = 0; l + 1 < L; ++l) {
      Conv::layer<0>(d, d.w[l], d.b[l], cin, C, d.act, true, T, H(l), H(l + 1), nullptr, xi, yi, nullptr);
      __syncthreads();
      cin = C;
    }
    // backward: z = sum of the last layer's outputs => its delta is 1
    for (size_t e = threadIdx.x; e < szC; e += kThreads) top[e] = 1.f;
    __syncthreads();
    for (int l = L - 1; l >= 1; --l) {
      Conv::layer<2>(d, d.wt[l], nullptr, C, C, d.act, false, T, DL + (size_t)l * szC,
                     DL + (size_t)(l - 1) * szC, nullptr, xib, yib, H(l));
      __syncthreads();
    }
    }
    // accumulate the weight gradients: dW_l[tap][ci][co] = sum_t w_k[t] sum_pos
    // h_l[t][ci][pos + tap] delta_l[t][co][pos].  Register tiles of 4 input x 4
    // output channels per (layer, tap): 8 shared-memory loads feed 16 FMAs
    // (one entry per thread needed 2 loads + 2 table look-ups per FMA and left
    // the kernel LSU-bound: ncu r01z, FMA pipe 11.6 %).
    if ((C & 3) == 0) {
      const int taps = d.kx * d.ky, cb = C / 4;
      const int items0 = taps * cb;                 // layer 0: one input channel
      const int items1 = taps * cb * cb;            // layers >= 1
      const int n_items = items0 + (L - 1) * items1;
#pragma unroll 1
      for (int it = threadIdx.x; it < n_items; it += kThreads) {
        int l = 0, r = it;
        if (it >= items0) { l = 1 + (it - items0) / items1; r = (it - items0) % items1; }
        const int cl = l == 0 ? 1 : C;
        const int cob = r % cb;
        const int cib = l == 0 ? 0 : (r / cb) % cb;
        const int tap = l == 0 ? r / cb : r / (cb * cb);
        const int dx = tap / d.ky, dy = tap - dx * d.ky;
        const float* hl = H(l);
        const float* dl = DL + (size_t)l * szC;
        float acc[K][4][4];
#pragma unroll
        for (int k = 0; k < K; ++k)
#pragma unroll
          for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[k][a][b] = 0.f;
        for (int t = 0; t < T; ++t) {
          const float* hp = hl + ((size_t)t * cl + (l == 0 ? 0 : 4 * cib)) * N;
          const float* dp = dl + ((size_t)t * C + 4 * cob) * N;
          float g[4][4];
#pragma unroll
          for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) g[a][b] = 0.f;
          for (int x = 0; x < d.X; ++x) {
            const int sx = xi[x * d.kx + dx];
            for (int y = 0; y < d.Y; ++y) {
              const int hidx = sx + yi[y * d.ky + dy], pos = x * d.Y + y;
              float hv[4], dv[4];
#pragma unroll
              for (int a = 0; a < 4; ++a) hv[a] = (l == 0 && a > 0) ? 0.f : hp[(size_t)a * N + hidx];
#pragma unroll
              for (int b = 0; b < 4; ++b) dv[b] = dp[(size_t)b * N + pos];
#pragma unroll
              for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) g[a][b] = fmaf(hv[a], dv[b], g[a][b]);
            }
          }
#pragma unroll
          for (int k = 0; k < K; ++k) {
            const float wkt = wk[k * T + t];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
              for (int b = 0; b < 4; ++b) acc[k][a][b] = fmaf(wkt, g[a][b], acc[k][a][b]);
          }
        }
        // flat layout of a conv weight tensor: [tap][ci][co]
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          if (l == 0 && a > 0) continue;
          const int ci = l == 0 ? 0 : 4 * cib + a;
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            const int64_t e = d.w_off[l] + ((int64_t)tap * cl + ci) * C + 4 * cob + b;
#pragma unroll
            for (int k = 0; k < K; ++k) {
              float* dst = part + (int64_t)k * P + e;
              *dst = first_tile ? acc[k][a][b] : *dst + acc[k][a][b];
            }
          }
        }
      }
    }
    // bias gradients (and, for channel counts that are not a multiple of 4, the
    // weight gradients entry by entry)
#pragma unroll 1
    for (int64_t e = threadIdx.x; e < P; e += kThreads) {
      float acc[K];
#pragma unroll
      for (int k = 0; k < K; ++k) acc[k] = 0.f;
      int l = 0;
      while (l + 1 < L && e >= d.w_off[l + 1]) ++l;
      const int cl = l == 0 ? 1 : C;
      const bool is_bias = e >= d.b_off[l];
      if (!is_bias && (C & 3) == 0) continue;       // done above
      const float* dl = DL + (size_t)l * szC;
      if (is_bias) {
        const int co = (int)(e - d.b_off[l]);
        for (int t = 0; t < T; ++t) {
          const float* dp = dl + ((size_t)t * C + co) * N;
          float g = 0.f;
          for (int pos = 0; pos < N; ++pos) g += dp[pos];
#pragma unroll
          for (int k = 0; k < K; ++k) acc[k] = fmaf(wk[k * T + t], g, acc[k]);
        }
      } else {
        const int q = (int)(e - d.w_off[l]);
        const int co = q % C, ci = (q / C) % cl, tap = q / (C * cl);
        const int dx = tap / d.ky, dy = tap - dx * d.ky;
        const float* hl = H(l);
        for (int t = 0; t < T; ++t) {
          const float* hp = hl + ((size_t)t * cl + ci) * N;
          const float* dp = dl + ((size_t)t * C + co) * N;
          float g = 0.f;
          for (int x = 0; x < d.X; ++x) {
            const int sx = xi[x * d.kx + dx];
            for (int y = 0; y < d.Y; ++y) g = fmaf(hp[sx + yi[y * d.ky + dy]], dp[x * d.Y + y], g);
          }
#pragma unroll
          for (int k = 0; k < K; ++k) acc[k] = fmaf(wk[k * T + t], g, acc[k]);
        }
      }
#pragma unroll
      for (int k = 0; k < K; ++k) {
        float* dst = part + (int64_t)k * P + e;
        *dst = first_tile ? acc[k] : *dst + acc[k];
      }
    }
    first_tile = false;
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
int build_desc(const cgsvmc_ansatz* a, NetDesc* d) {
  const cgsvmc_ansatz_desc& s = a->desc;
  memset(d, 0, sizeof(*d));
  d->kind = s.kind; d->N = s.n_sites; d->L = s.num_layers; d->act = s.nonlinearity;
  d->NW = n_words(s.n_sites);
  if (s.num_layers > kMaxLayers) { set_error("more than 16 layers are not supported"); return CGSVMC_ERR_UNSUPPORTED; }
  const float* p = a->params;
  int idx = 0;
  if (s.kind == CGSVMC_ANSATZ_RBM) {
    d->rbm_a = p + a->offsets[0];
    d->rbm_a0 = p + a->offsets[1];
    idx = 2;
  }
  if (s.kind == CGSVMC_ANSATZ_FULLY_CONNECTED || s.kind == CGSVMC_ANSATZ_RBM) {
    d->D = s.layer_size;
    // forward kernels stage one layer's weights in shared memory when they fit 64 KB
    d->stage_weights = (size_t)std::max(s.layer_size, s.n_sites) * s.layer_size * sizeof(float) <= 64 * 1024 &&
                       (s.num_layers > 0 || s.kind == CGSVMC_ANSATZ_RBM);
    {   // all matrices resident when they fit 96 KB (C1: 59 KB)
      const int n_mat = s.num_layers + (s.kind == CGSVMC_ANSATZ_RBM ? 1 : 0);
      const size_t all = n_mat == 0 ? 0 : ((size_t)s.n_sites * s.layer_size +
                                           (size_t)(n_mat - 1) * s.layer_size * s.layer_size) * sizeof(float);
      d->resident_weights = n_mat > 0 && all <= 96 * 1024;
    }
    for (int l = 0; l <= s.num_layers; ++l) {
      d->w[l] = p + a->offsets[idx]; d->w_off[l] = a->offsets[idx]; ++idx;
      d->b[l] = p + a->offsets[idx]; d->b_off[l] = a->offsets[idx]; ++idx;
    }
  } else {
    d->C = s.num_filters;
    if (s.kind == CGSVMC_ANSATZ_RESNET_1D || s.kind == CGSVMC_ANSATZ_RESNET_2D) {
      d->resnet = 1;
      d->L = 1 + 2 * s.num_layers;        // initial conv + two convs per block
      d->act = CGSVMC_ACT_SELU;
    }
    if (s.kind == CGSVMC_ANSATZ_CONV_1D || s.kind == CGSVMC_ANSATZ_RESNET_1D) {
      d->X = s.n_sites; d->Y = 1; d->kx = s.kernel_size; d->ky = 1;
      d->pad_x = s.kernel_size % 2 ? (s.kernel_size - 1) / 2 : s.kernel_size / 2;   // layers.py:64-73
      d->pad_y = 0;
    } else {
      d->X = s.size_x; d->Y = s.size_y; d->kx = d->ky = s.kernel_size;
      d->pad_x = d->pad_y = s.kernel_size % 2 ? (s.kernel_size - 1) / 2 : s.kernel_size / 2 - 1;   // layers.py:132-141
    }
    for (int l = 0; l < d->L; ++l) {
      d->w[l] = p + a->offsets[2 * l]; d->w_off[l] = a->offsets[2 * l];
      d->b[l] = p + a->offsets[2 * l + 1]; d->b_off[l] = a->offsets[2 * l + 1];
    }
  }
  return CGSVMC_OK;
}

bool is_conv(const NetDesc& d) {
  return d.kind == CGSVMC_ANSATZ_CONV_1D || d.kind == CGSVMC_ANSATZ_CONV_2D ||
         d.kind == CGSVMC_ANSATZ_RESNET_1D || d.kind == CGSVMC_ANSATZ_RESNET_2D;
}

template <typename F>
int set_smem(F kernel, size_t bytes, const cgsvmc_ansatz* a) {
  if ((int64_t)bytes > a->max_smem_optin) {
    set_error("network needs more shared memory per tile than the device offers");
    return CGSVMC_ERR_UNSUPPORTED;
  }
  if (bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(smem)");
  }
  return CGSVMC_OK;
}

// Tile size for the conv family: as many configurations as fit ~100 KB of
// activations, rounded so T * N is close to a multiple of the CTA size.
int conv_tile(const NetDesc& d, int64_t work_items, int num_sms, size_t budget = 96 * 1024) {
  const size_t per_cfg = 2 * (size_t)d.C * d.N * sizeof(float);
  int t = (int)std::max<size_t>(1, budget / per_cfg);
  t = std::min(t, 64);
  // do not starve the grid: at least ~2 tiles per SM when the work allows
  while (t > 1 && work_items / t < 2 * (int64_t)num_sms) t = (t + 1) / 2;
  return t;
}

struct Plan { int T; bool small_tile; size_t fwd_bytes; };

Plan make_plan(const cgsvmc_ansatz* a, const NetDesc& d, int64_t work_items) {
  Plan p;
  if (is_conv(d)) {
    p.T = conv_tile(d, work_items, a->num_sms);
    p.small_tile = false;
    p.fwd_bytes = (ConvNet::fwd_smem(d, p.T) + 15) / 16 * 16;
  } else {
    p.small_tile = work_items / 32 < 2 * (int64_t)a->num_sms;
    p.T = p.small_tile ? 8 : 32;
    p.fwd_bytes = ((p.small_tile ? MlpNet<1>::fwd_smem(d, 0) : MlpNet<4>::fwd_smem(d, 0)) + 15) / 16 * 16;
  }
  return p;
}

int grid_for(const cgsvmc_ansatz* a, int64_t tiles) {
  return (int)std::max<int64_t>(1, std::min<int64_t>(tiles, (int64_t)a->num_sms * 4));
}

#define NET_LAUNCH(KERNEL, SMEM, GRID, ...)                                          \
  do {                                                                               \
    if (is_conv(d)) {                                                                \
      auto kern = KERNEL<ConvNet>;                                                   \
      if (int rc = set_smem(kern, SMEM, a)) return rc;                               \
      kern<<<GRID, kThreads, SMEM, st>>>(__VA_ARGS__);                               \
    } else if (plan.small_tile) {                                                    \
      auto kern = KERNEL<MlpNet<1>>;                                                 \
      if (int rc = set_smem(kern, SMEM, a)) return rc;                               \
      kern<<<GRID, kThreads, SMEM, st>>>(__VA_ARGS__);                               \
    } else {                                                                         \
      auto kern = KERNEL<MlpNet<4>>;                                                 \
      if (int rc = set_smem(kern, SMEM, a)) return rc;                               \
      kern<<<GRID, kThreads, SMEM, st>>>(__VA_ARGS__);                               \
    }                                                                                \
  } while (0)

}  // namespace

int net_log_amp(const cgsvmc_ansatz* a, const uint64_t* packed, int64_t B, float* out,
                cudaStream_t st) {
  NetDesc d;
  if (int rc = build_desc(a, &d)) return rc;
  const Plan plan = make_plan(a, d, B);
  const size_t smem = plan.fwd_bytes + carve_bytes((size_t)plan.T * d.NW, 8) + carve_bytes(plan.T, 4);
  const int grid = grid_for(a, (B + plan.T - 1) / plan.T);
  NET_LAUNCH(net_log_amp_kernel, smem, grid, d, plan.T, plan.fwd_bytes, packed, B, out);
  return cuda_fail(cudaGetLastError(), "net_log_amp launch");
}

int net_mc_steps(const cgsvmc_ansatz* a, uint64_t* packed, int64_t B, int n_steps, uint64_t seed,
                 uint64_t walker0, uint64_t step0, unsigned long long* accept_count,
                 float* log_amp_out, cudaStream_t st) {
  NetDesc d;
  if (int rc = build_desc(a, &d)) return rc;
  const Plan plan = make_plan(a, d, B);
  const size_t smem = plan.fwd_bytes + 2 * carve_bytes((size_t)plan.T * d.NW, 8) + 3 * carve_bytes(plan.T, 4);
  const int grid = grid_for(a, (B + plan.T - 1) / plan.T);
  NET_LAUNCH(net_mc_kernel, smem, grid, d, plan.T, plan.fwd_bytes, packed, B, n_steps, seed,
             walker0, step0, a->step_counter_dev, accept_count, log_amp_out);
  return cuda_fail(cudaGetLastError(), "net_mc_steps launch");
}

int net_mc_replay(const cgsvmc_ansatz* a, uint64_t* packed, int64_t B, const float* u_sites,
                  const float* u_acc, int32_t* down, int32_t* up, float* log_ratio,
                  uint8_t* accept, cudaStream_t st) {
  NetDesc d;
  if (int rc = build_desc(a, &d)) return rc;
  const Plan plan = make_plan(a, d, B);
  const size_t smem = plan.fwd_bytes + 2 * carve_bytes((size_t)plan.T * d.NW, 8) + 2 * carve_bytes(plan.T, 4);
  const int grid = grid_for(a, (B + plan.T - 1) / plan.T);
  NET_LAUNCH(net_replay_kernel, smem, grid, d, plan.T, plan.fwd_bytes, packed, B, u_sites, u_acc,
             down, up, log_ratio, accept);
  return cuda_fail(cudaGetLastError(), "net_mc_replay launch");
}

int net_local_energy(const cgsvmc_ansatz* a, const cgsvmc_ham* h, const uint64_t* packed,
                     int64_t B, float* e_loc, float* log_amp_out, float* diag_out, float* off_out,
                     cudaStream_t st) {
  NetDesc d;
  if (int rc = build_desc(a, &d)) return rc;
  if (h->n_bonds >= 65535) { set_error("local_energy: more than 65534 bonds"); return CGSVMC_ERR_UNSUPPORTED; }
  const Plan plan = make_plan(a, d, B * (1 + h->n_bonds / 2));
  const size_t max_items = (size_t)kElocWalkers * (h->n_bonds + 1);
  const size_t smem = plan.fwd_bytes + carve_bytes((size_t)plan.T * d.NW, 8) +
                      carve_bytes((size_t)kElocWalkers * d.NW, 8) + carve_bytes(plan.T, 4) +
                      2 * carve_bytes(max_items, 4) + carve_bytes(kElocWalkers, 4);
  const int grid = grid_for(a, (B + kElocWalkers - 1) / kElocWalkers);
  NET_LAUNCH(net_eloc_kernel, smem, grid, d, plan.T, plan.fwd_bytes, h->ij, h->jx, h->jz,
             h->n_bonds, packed, B, e_loc, log_amp_out, diag_out, off_out);
  return cuda_fail(cudaGetLastError(), "net_local_energy launch");
}

int net_grad(cgsvmc_ansatz* a, const uint64_t* packed, const float* weights, int64_t B, int K,
             float* out, cudaStream_t st) {
  NetDesc d;
  if (int rc = build_desc(a, &d)) return rc;
  if (d.act == CGSVMC_ACT_COS && d.L > (is_conv(d) ? 1 : 0)) {
    set_error("weighted_grad_sum: the cos nonlinearity has no gradient kernel");
    return CGSVMC_ERR_UNSUPPORTED;
  }
  const int64_t P = a->n_params;
  const int chunks = (int)((P + kThreads * kGradE - 1) / (kThreads * kGradE));
  const bool conv = is_conv(d);
  // shared memory plan
  int T = 0;
  size_t smem = 0;
  bool tw4 = true;
  if (conv) {
    const size_t per_cfg = ((size_t)d.N + (size_t)(2 * d.L - 1) * d.C * d.N) * 4;
    const size_t fixed = (size_t)2 * (d.X * d.kx + d.Y * d.ky) * 4 + 64;
    T = (int)std::min<size_t>(4, (size_t)(a->max_smem_optin - fixed - 256) / (per_cfg + d.NW * 8 + 8));
    if (T < 1) { set_error("weighted_grad_sum: network does not fit in shared memory"); return CGSVMC_ERR_UNSUPPORTED; }
    smem = (size_t)T * per_cfg + 2 * (size_t)T * 4 + fixed + (size_t)T * d.NW * 8 + 32;
  } else {
    const bool rbm = d.kind == CGSVMC_ANSATZ_RBM;
    const size_t stage_bytes = (size_t)std::max(d.D, d.N) * d.D * 4 + 16;
    auto need = [&](int ts, int t) {
      return ((size_t)d.N + 2 * (size_t)d.L * d.D + (rbm ? d.D : 0) + 1 + 2) * ts * 4 + 16 +
             (size_t)t * d.NW * 8 + (d.stage_weights ? stage_bytes : 0);
    };
    if (d.stage_weights && (int64_t)need(36, 32) > a->max_smem_optin - 1024) d.stage_weights = 0;
    if ((int64_t)need(36, 32) > a->max_smem_optin - 1024) tw4 = false;
    T = tw4 ? 32 : 8;
    smem = need(tw4 ? 36 : 12, T);
    if ((int64_t)smem > a->max_smem_optin) { set_error("weighted_grad_sum: network does not fit in shared memory"); return CGSVMC_ERR_UNSUPPORTED; }
  }
  // conv: one CTA per SM handles all parameter entries of its walkers
  int64_t groups = conv ? (int64_t)a->num_sms : std::max<int64_t>(1, (2 * (int64_t)a->num_sms) / chunks);
  groups = std::min<int64_t>(groups, (B + T - 1) / T);
  int64_t per_cta = (B + groups - 1) / groups;
  per_cta = (per_cta + T - 1) / T * T;
  groups = (B + per_cta - 1) / per_cta;
  // scratch: transposed weights (P floats is an upper bound) + partials
  const size_t wt_floats = (size_t)P;
  if (int rc = ensure_scratch(a, (wt_floats + (size_t)groups * 2 * P) * sizeof(float))) return rc;
  float* wt = a->scratch;
  float* partials = a->scratch + wt_floats;
  {
    size_t off = 0;
    auto transpose = [&](const float* src, int batch, int rows, int cols) {
      float* dst = wt + off;
      const int64_t total = (int64_t)batch * rows * cols;
      const int blocks = (int)std::min<int64_t>((total + 255) / 256, 1024);
      transpose_batched_kernel<<<blocks, 256, 0, st>>>(src, batch, rows, cols, dst);
      off += (size_t)total;
      return (const float*)dst;
    };
    if (conv) {
      for (int l = 1; l < d.L; ++l) d.wt[l] = transpose(d.w[l], d.kx * d.ky, d.C, d.C);
    } else {
      for (int l = 1; l < d.L; ++l) d.wt[l] = transpose(d.w[l], 1, d.D, d.D);
      if (d.kind == CGSVMC_ANSATZ_RBM && d.L >= 1) d.wt[d.L] = transpose(d.w[d.L], 1, d.D, d.D);
    }
    if (int rc = cuda_fail(cudaGetLastError(), "transpose launch")) return rc;
  }
  dim3 grid((unsigned)groups, conv ? 1u : (unsigned)chunks);
  for (int k0 = 0; k0 < K; k0 += 2) {
    const int kk = std::min(2, K - k0);
    const float* w = weights + (int64_t)k0 * B;
#define GRAD_LAUNCH(KERN, ...)                                                       \
    do {                                                                             \
      auto kern = KERN;                                                              \
      if (int rc = set_smem(kern, smem, a)) return rc;                               \
      kern<<<grid, kThreads, smem, st>>>(__VA_ARGS__);                               \
    } while (0)
    if (conv) {
      if (kk == 1) GRAD_LAUNCH(conv_grad_kernel<1>, d, T, packed, w, B, per_cta, P, partials);
      else GRAD_LAUNCH(conv_grad_kernel<2>, d, T, packed, w, B, per_cta, P, partials);
    } else if (tw4) {
      if (kk == 1) GRAD_LAUNCH((mlp_grad_kernel<4, 1>), d, packed, w, B, per_cta, P, partials);
      else GRAD_LAUNCH((mlp_grad_kernel<4, 2>), d, packed, w, B, per_cta, P, partials);
    } else {
      if (kk == 1) GRAD_LAUNCH((mlp_grad_kernel<1, 1>), d, packed, w, B, per_cta, P, partials);
      else GRAD_LAUNCH((mlp_grad_kernel<1, 2>), d, packed, w, B, per_cta, P, partials);
    }
#undef GRAD_LAUNCH
    if (int rc = cuda_fail(cudaGetLastError(), "net_grad launch")) return rc;
    if (int rc = launch_reduce_partials(partials, (int)groups, (int64_t)kk * P,
                                        out + (int64_t)k0 * P, st)) return rc;
  }
  return CGSVMC_OK;
}

}  // namespace cgsvmc
