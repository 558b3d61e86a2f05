// Micro-benchmark: issue cost (cycles per warp instruction per scheduler) of the instructions in the softmax pass:
// FFMA, FFMA2, FADD2, FMNMX, FMNMX3, F2FP.BF16.PACK_AB, IMAD (shift-add), MUFU.EX2. 8 warps per scheduler, 8 chains.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o pipe_bench pipe_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t pack2(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack2(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }

template <int MODE>
__global__ void __launch_bounds__(1024, 1) bench(float* out, float seed, int iters) {
  constexpr int N = 8;
  float a[N], b[N];
  uint64_t v[N];
  uint32_t w[N];
#pragma unroll
  for (int i = 0; i < N; ++i) { a[i] = seed * (threadIdx.x + i) * 1e-3f; b[i] = a[i] + 1.f; v[i] = pack2(a[i], b[i]); w[i] = threadIdx.x + i; }
  const uint64_t c2 = pack2(seed * 0.999f, seed * 0.999f), d2 = pack2(seed * 1e-3f, seed * 1e-3f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
      if (MODE == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(seed), "f"(b[i]));
      if (MODE == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v[i]) : "l"(c2), "l"(d2));
      if (MODE == 2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(v[i]) : "l"(d2));
      if (MODE == 3) asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b[i]));
      if (MODE == 4) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b[i]), "f"(seed));
      if (MODE == 5) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w[i]) : "f"(a[i]), "f"(__uint_as_float(w[i])));
      if (MODE == 6) asm volatile("mad.lo.s32 %0, %0, 8388608, %1;" : "+r"(w[i]) : "r"(w[(i + 1) % N]));
      if (MODE == 7) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (MODE == 8) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(v[i]) : "l"(c2));
      if (MODE == 9) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(w[i]) : "r"(w[(i + 1) % N]), "r"(w[(i + 2) % N]));
      if (MODE == 10) asm volatile("prmt.b32 %0, %0, %1, 0x7632;" : "+r"(w[i]) : "r"(w[(i + 1) % N]));
      if (MODE == 11) asm volatile("max.bf16x2 %0, %0, %1;" : "+r"(w[i]) : "r"(w[(i + 1) % N]));
      if (MODE == 12) asm volatile("or.b32 %0, %0, %1;" : "+r"(w[i]) : "r"(w[(i + 1) % N]));
      if (MODE == 13) asm volatile("max.u32 %0, %0, %1;" : "+r"(w[i]) : "r"(w[(i + 1) % N]));
      if (MODE == 14) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b[i]));
      if (MODE == 15) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i])); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(b[i]) : "f"(seed), "f"(b[i])); asm volatile("fma.rn.f32 %0, %0, %1, %0;" : "+f"(b[i]) : "f"(seed)); asm volatile("fma.rn.f32 %0, %0, %1, %0;" : "+f"(b[i]) : "f"(seed)); }
      if (MODE == 16) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i])); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(b[i]) : "f"(seed), "f"(b[i])); asm volatile("fma.rn.f32 %0, %0, %1, %0;" : "+f"(b[i]) : "f"(seed)); asm volatile("fma.rn.f32 %0, %0, %1, %0;" : "+f"(b[i]) : "f"(seed)); asm volatile("fma.rn.f32 %0, %0, %1, %0;" : "+f"(b[i]) : "f"(seed)); asm volatile("fma.rn.f32 %0, %0, %1, %0;" : "+f"(b[i]) : "f"(seed)); asm volatile("fma.rn.f32 %0, %0, %1, %0;" : "+f"(b[i]) : "f"(seed)); }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < N; ++i) { float x, y; unpack2(v[i], x, y); s += a[i] + x + y + __uint_as_float(w[i]); }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name) {
  float* out;
  cudaMalloc(&out, 148 * 1024 * sizeof(float));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 4096;
  bench<MODE><<<148, 1024>>>(out, 1.0f, iters);
  cudaEventRecord(e0);
  bench<MODE><<<148, 1024>>>(out, 1.0f, iters);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double winst_per_sched = 8.0 * iters * 8;  // warps/scheduler * iters * chains
  printf("%-28s %7.3f ms  %.2f cycles per warp instruction per scheduler (at 1.9 GHz)\n", name, ms, ms * 1e-3 * 1.9e9 / winst_per_sched);
  cudaFree(out);
}

int main() {
  run<0>("FFMA");
  run<1>("FFMA2 (fma.rn.f32x2)");
  run<2>("FADD2 (add.rn.f32x2)");
  run<8>("FMUL2 (mul.rn.f32x2)");
  run<3>("FMNMX");
  run<4>("FMNMX3");
  run<5>("F2FP.BF16.PACK_AB");
  run<6>("IMAD (x*2^23 + y)");
  run<9>("LOP3");
  run<7>("MUFU.EX2");
  run<10>("PRMT");
  run<11>("HMNMX2.BF16 (max.bf16x2)");
  run<12>("LOP (2-input or)");
  run<13>("IMNMX.U32");
  run<14>("FADD");
  run<15>("1 MUFU + 3 FFMA (per group)");
  run<16>("1 MUFU + 6 FFMA (per group)");
  return 0;
}
