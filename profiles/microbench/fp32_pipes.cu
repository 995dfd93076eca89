// Micro-benchmark (development aid): issue rate of scalar FFMA / FMUL, packed
// fma.rn.f32x2 / mul.rn.f32x2, FFMA interleaved with integer ALU work, and
// LDS.128 on one B200 SM sub-partition, to decide how the RBM ratio loop should
// be written.   nvcc -arch=sm_100a -O3 -o fp32_pipes fp32_pipes.cu && ./fp32_pipes
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096

template <int MODE>
__global__ void __launch_bounds__(512) kern(float* out, float a, float b, long long* cycles) {
  __shared__ float4 sm[512];
  sm[threadIdx.x] = make_float4(a, b, a, b);
  __syncthreads();
  float x[16];
  unsigned int u[8];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = a + i;
#pragma unroll
  for (int i = 0; i < 8; ++i) u[i] = threadIdx.x + i;
  const long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
    if (MODE == 0) {          // 16 independent scalar FFMA
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], a, b);
    } else if (MODE == 1) {   // 8 independent fma.rn.f32x2 (same 16 FMAs)
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        asm volatile("{ .reg .b64 r, s, t; mov.b64 r, {%0, %1}; mov.b64 s, {%2, %2}; mov.b64 t, {%3, %3};\n"
                     "  fma.rn.f32x2 r, r, s, t; mov.b64 {%0, %1}, r; }"
                     : "+f"(x[i]), "+f"(x[i + 1]) : "f"(a), "f"(b));
      }
    } else if (MODE == 2) {   // 16 scalar FMUL
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = x[i] * a;
    } else if (MODE == 3) {   // 16 FFMA + 8 integer ALU ops (LOP3/IADD3)
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], a, b);
#pragma unroll
      for (int i = 0; i < 8; ++i) u[i] = (u[i] ^ (u[i] >> 3)) + 0x9e3779b9u;
    } else if (MODE == 4) {   // 4 LDS.128 (conflict-free, 512 B per warp each)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 v = sm[(threadIdx.x + 32 * i + it) & 511];
        x[4 * i] += v.x; x[4 * i + 1] += v.y; x[4 * i + 2] += v.z; x[4 * i + 3] += v.w;
      }
    } else if (MODE == 5) {   // 8 independent mul.rn.f32x2
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        asm volatile("{ .reg .b64 r, s; mov.b64 r, {%0, %1}; mov.b64 s, {%2, %2};\n"
                     "  mul.rn.f32x2 r, r, s; mov.b64 {%0, %1}, r; }"
                     : "+f"(x[i]), "+f"(x[i + 1]) : "f"(a));
      }
    } else if (MODE == 6) {   // 16 FFMA with distinct register operands x[i] = x[i] * x[(i+1)&15] + x[(i+2)&15]
#pragma unroll
      for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], x[(i + 5) & 15], x[(i + 9) & 15]);
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i];
#pragma unroll
  for (int i = 0; i < 8; ++i) s += (float)u[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int ops_per_iter, int threads) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
  kern<MODE><<<148, threads>>>(out, 1.0001f, 0.0001f, cyc);
  kern<MODE><<<148, threads>>>(out, 1.0001f, 0.0001f, cyc);
  long long h[148];
  cudaMemcpy(h, cyc, 148 * 8, cudaMemcpyDeviceToHost);
  const double warps_per_smsp = threads / 32 / 4.0;
  const double per_warp_instr = (double)ITERS * ops_per_iter;
  printf("%-44s threads %4d  cycles %9lld  warp-instr/clk/SMSP %.3f\n", name, threads, h[0],
         per_warp_instr * warps_per_smsp / (double)h[0]);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int threads : {128, 256, 512}) {
    run<0>("16 scalar FFMA (2 reg + 1 reg)", 16, threads);
    run<6>("16 scalar FFMA (3 distinct regs)", 16, threads);
    run<1>("8 fma.rn.f32x2", 8, threads);
    run<2>("16 scalar FMUL", 16, threads);
    run<5>("8 mul.rn.f32x2", 8, threads);
    run<3>("16 FFMA + 8x(LOP3,SHF,IADD) mixed", 16 + 24, threads);
    run<4>("4 LDS.128 + 16 FADD", 4 + 16, threads);
  }
  return 0;
}
