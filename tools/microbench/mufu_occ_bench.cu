// Micro-benchmark: MUFU.EX2 throughput per SM as a function of the number of warps per scheduler (1, 2, 3, 4, 8) and of
// the instruction mix (pure MUFU; 1 MUFU : k FMA-pipe instructions, interleaved in program order). One CTA per SM.
// Question it answers: can the two softmax warps per scheduler of the attention kernel saturate the 16 exp/clk/SM unit?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_occ_bench mufu_occ_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2f(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int FMA_PER_MUFU, int ILP>
__global__ void __launch_bounds__(1024, 1) bench(float* out, float seed, int iters) {
  float a[ILP], b[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { a[i] = seed * (threadIdx.x + i) * 1e-3f - 1.0f; b[i] = a[i] * 0.5f; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      a[i] = ex2f(a[i]) - 1.5f;   // MUFU + FADD
#pragma unroll
      for (int k = 0; k < FMA_PER_MUFU; ++k) b[i] = fmaf(b[i], 0.999f, 0.001f);
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += a[i] + b[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int F, int ILP>
void run(int warps_per_sched) {
  float* out;
  cudaMalloc(&out, 148 * 1024 * sizeof(float));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int threads = warps_per_sched * 4 * 32, iters = 4096;
  bench<F, ILP><<<148, threads>>>(out, 1.0f, iters);
  cudaEventRecord(e0);
  bench<F, ILP><<<148, threads>>>(out, 1.0f, iters);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double ops = 148.0 * threads * iters * ILP;
  printf("warps/sched=%d  fma:mufu=%d  ilp=%d : %7.3f ms  %6.2f exp/clk/SM (at 1.9 GHz)\n", warps_per_sched, F, ILP, ms,
         ops / ms / 1e6 / 148 / 1.9);
  cudaFree(out);
}

int main() {
  for (int w : {1, 2, 3, 4, 8}) run<0, 8>(w);
  for (int w : {1, 2, 3, 4, 8}) run<3, 8>(w);
  for (int w : {1, 2, 3, 4}) run<3, 16>(w);
  for (int w : {1, 2, 3, 4}) run<1, 16>(w);
  return 0;
}
