// Micro-benchmark: MUFU.EX2 throughput on f32 vs packed bf16x2 / f16x2 inputs, and a Cody-Waite polynomial exp2 on the
// FMA pipe. Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_bench mufu_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int kIters = 4096;
constexpr int kIlp = 8;

__device__ __forceinline__ float ex2f(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t ex2bf2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ uint32_t ex2h2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ float ex2poly(float x) {
  x = fmaxf(x, -126.0f);
  const float t = x + 12582912.0f;
  const float fl = t - 12582912.0f;
  const float f = x - fl;
  float p = fmaf(f, 0.0555041f, 0.2402265f);
  p = fmaf(p, f, 0.6931472f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

template <int MODE>
__global__ void bench(float* out, float seed) {
  float a[kIlp];
  uint32_t u[kIlp];
#pragma unroll
  for (int i = 0; i < kIlp; ++i) { a[i] = seed * (threadIdx.x + i) * 1e-3f - 1.0f; u[i] = __float_as_uint(a[i]); }
  for (int it = 0; it < kIters; ++it) {
#pragma unroll
    for (int i = 0; i < kIlp; ++i) {
      if (MODE == 0) a[i] = ex2f(a[i]) - 1.5f;
      if (MODE == 1) u[i] = ex2bf2(u[i]) ^ 0x80008000u;
      if (MODE == 2) u[i] = ex2h2(u[i]) ^ 0x80008000u;
      if (MODE == 3) a[i] = ex2poly(a[i]) - 1.5f;
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kIlp; ++i) s += a[i] + __uint_as_float(u[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, int per_op) {
  float* out;
  cudaMalloc(&out, 148 * 8 * 1024 * sizeof(float));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  bench<MODE><<<148 * 8, 256>>>(out, 1.0f);
  cudaEventRecord(e0);
  bench<MODE><<<148 * 8, 256>>>(out, 1.0f);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double ops = 148.0 * 8 * 256 * kIters * kIlp * per_op;
  printf("%-28s %8.3f ms  %8.2f Gexp/s  (%.2f exp/clk/SM at 1.9 GHz)\n", name, ms, ops / ms / 1e6, ops / ms / 1e6 / 148 / 1.9);
  cudaFree(out);
}

int main() {
  run<0>("ex2.approx.ftz.f32", 1);
  run<1>("ex2.approx.ftz.bf16x2", 2);
  run<2>("ex2.approx.f16x2", 2);
  run<3>("poly exp2 (FMA pipe)", 1);
  // accuracy of the polynomial
  return 0;
}
