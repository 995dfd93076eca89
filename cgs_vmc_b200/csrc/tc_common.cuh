// tcgen05 / TMEM / TMA building blocks shared by the tensor-core kernels
// (conv_tc.cu: periodic convolutions; fc_tc.cu: fully connected layers):
// mbarriers, TMA bulk copies, UMMA shared-memory descriptors and issue,
// TMEM loads, and the three-way fp16 split that gives float32-grade products.
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"
#include "internal.h"

namespace cgsvmc {
namespace tc {

constexpr float kSplitScale = 2048.f;   // S = 2^11: keeps the lower split terms out of the fp16 subnormals

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ float tc_activate(int act, float x) {
  switch (act) {
    case CGSVMC_ACT_RELU: return fmaxf(x, 0.f);
    case CGSVMC_ACT_TANH: return tanh_accurate(x);
    case CGSVMC_ACT_SIGMOID: return 1.f / (1.f + expf(-x));
    case CGSVMC_ACT_IDENTITY: return x;
    case CGSVMC_ACT_COS: return cosf(x);
    case CGSVMC_ACT_EXP: return expf(x);
    default: return tanf(x);
  }
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
        : "=r"(ok)
        : "r"(a), "r"(parity)
        : "memory");
  }
}

// TMA bulk copy global -> shared, completion on `bar` (one arrival + bytes).
__device__ __forceinline__ void bulk_load_async(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  const uint32_t bar_a = smem_u32(bar);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes) : "memory");
  uint32_t done = 0;
  while (done < bytes) {
    const uint32_t chunk = min(bytes - done, 32768u);
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(reinterpret_cast<char*>(dst) + done)),
        "l"(reinterpret_cast<const char*>(src) + done), "r"(chunk), "r"(bar_a)
        : "memory");
    done += chunk;
  }
}

// Shared-memory matrix descriptors (cute::UMMA::SmemDescriptor), K-major, no
// swizzle: bits [0,14) start address, [16,30) leading byte offset (between the
// two 16-byte K chunks), [32,46) stride byte offset (between 8-row groups), all
// in 16-byte units; bits [46,48) version = 1.  Built inline in tensor_layer().
__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
      "}\n"
      :
      : "r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}

// One lane of the (converged) warp; lets the compiler move the MMA operands to
// uniform registers without a per-lane waterfall loop.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}


// v = h1 + h2 / S + h3 / S^2 with fp16 parts (33 mantissa bits: exact for float32)
__device__ __forceinline__ void split3(float v, __half& h1, __half& h2, __half& h3) {
  h1 = __float2half_rn(v);
  const float r1 = (v - __half2float(h1)) * kSplitScale;
  h2 = __float2half_rn(r1);
  const float r2 = (r1 - __half2float(h2)) * kSplitScale;
  h3 = __float2half_rn(r2);
}

// The same for two values at once (packed conversions; identical roundings):
// h_k = (part k of a, part k of b).
__device__ __forceinline__ void split3_pair(float a, float b, __half2& h1, __half2& h2, __half2& h3) {
  h1 = __floats2half2_rn(a, b);
  const float2 f1 = __half22float2(h1);
  float ra = (a - f1.x) * kSplitScale, rb = (b - f1.y) * kSplitScale;
  h2 = __floats2half2_rn(ra, rb);
  const float2 f2 = __half22float2(h2);
  ra = (ra - f2.x) * kSplitScale;
  rb = (rb - f2.y) * kSplitScale;
  h3 = __floats2half2_rn(ra, rb);
}

// 8-column TMEM load (32 lanes x 32 bit x 8 columns) WITHOUT the wait: issue
// several, then tmem_ld_wait() once.
__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

}  // namespace tc
}  // namespace cgsvmc
