// Micro-benchmark: (A) tcgen05.ld throughput per SM for the access pattern of the attention softmax (32x32b.x32, one
// 128-column score row per thread), with 4 / 8 / 16 warps; (B) cost of one tcgen05.mma (M = 128, K = 16, bf16) as a
// function of N and of where A lives (shared memory vs TMEM), issued back to back by one thread; (C) both at once.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../video_latent_diffusion_panoptic_segmentation_b200/csrc
//        -o tmem_bench tmem_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda.h>

#include "common.cuh"

using namespace ldm;

constexpr int kLdIters = 2000;

// mode 0: ld only; 1: MMA only; 2: both.  n_mma: N of the MMA; ts: A operand from TMEM; per_commit: MMAs per commit
__global__ void __launch_bounds__(640, 1)
bench(int mode, int ld_warps, int n_mma, int ts, int mma_iters, int cols_per_ld, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  // zero the operand tiles (A: 128 x 64 bf16 = 16 KB, B: 256 x 64 bf16 = 32 KB)
  for (int i = threadIdx.x; i < (16 + 32) * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  fence_proxy_async_smem();
  if (warp == 0) {
    tmem_alloc(&tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  long long t0 = clock64();
  if (warp == 0) {
    if (mode != 0) {
      const uint32_t idesc = umma_idesc_bf16(128, n_mma);
      const uint32_t a = smem_u32(smem), b = smem_u32(smem + 16 * 1024);
      if (elect_one()) {
        for (int it = 0; it < mma_iters; ++it) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint64_t da = umma_desc_k_sw128(a + kk * 32);
            const uint64_t db = umma_desc_k_sw128(b + kk * 32);
            // cols_per_ld doubles as a switch here: 1 = alternate between two accumulators (are back-to-back MMAs into the
            // same D serialised?), 2 = one accumulator per k-step (four)
            const uint32_t d_off = cols_per_ld == 1 ? (kk & 1) * n_mma : (cols_per_ld == 2 ? kk * n_mma : 0);
            if (ts == 2) {
              // stage this k-step's A slab (128 rows x 32 B) in TMEM, then read it from there
              const uint32_t stage_col = tmem_base + 128 + ((it & 1) * 4 + kk) * 8;
              asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(stage_col), "l"(da) : "memory");
              umma_bf16_ts(tmem_base + 256 + d_off, stage_col, db, idesc, 1);
            } else if (ts == 3) {
              // the four copies of a k-block first, then its four MMAs
              if (kk == 0) {
#pragma unroll
                for (int k2 = 0; k2 < 4; ++k2)
                  asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(tmem_base + 128 + ((it & 1) * 4 + k2) * 8),
                               "l"(umma_desc_k_sw128(a + k2 * 32))
                               : "memory");
              }
              umma_bf16_ts(tmem_base + 256 + d_off, tmem_base + 128 + ((it & 1) * 4 + kk) * 8, db, idesc, 1);
            } else if (ts)
              umma_bf16_ts(tmem_base + 256 + d_off, tmem_base + 128 + kk * 8, db, idesc, 1);
            else
              umma_bf16(tmem_base + 256 + d_off, da, db, idesc, 1);
          }
        }
        umma_commit(&bar);
      }
      __syncwarp();
      mbar_wait(&bar, 0);
      tc_fence_after();
    }
  } else if (warp >= 4 && warp < 4 + ld_warps) {
    if (mode != 1) {
      const uint32_t t = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
      uint32_t acc = 0;
      for (int it = 0; it < kLdIters; ++it) {
        uint32_t v[128];
        if (cols_per_ld == 128) {
#pragma unroll
          for (int c = 0; c < 128; c += 32) tmem_ld32(t + c, v + c);
        } else {
#pragma unroll
          for (int c = 0; c < 128; c += 32) tmem_ld32(t + (c & 32), v + c);
        }
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 128; ++c) acc ^= v[c];
      }
      out[256 + blockIdx.x * 1024 + threadIdx.x] = acc;
    }
  }
  tc_fence_before();
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// load-only kernel: every warp reads `cols` columns of its lane quadrant per iteration
template <int COLS>
__global__ void __launch_bounds__(512, 1) ld_kernel(uint32_t* out, int iters) {
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    tmem_alloc(&tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t t = tmem_slot + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 128;
  uint32_t acc = 0;
  for (int it = 0; it < iters; ++it) {
    uint32_t v[COLS];
#pragma unroll
    for (int c = 0; c < COLS; c += 32) tmem_ld32(t + c, v + c);
    tmem_ld_wait();
#pragma unroll
    for (int c = 0; c < COLS; ++c) acc ^= v[c];
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_slot, 512);
  }
}

template <int COLS>
static void run_ld(int warps) {
  uint32_t* out;
  cudaMalloc(&out, 148 * 512 * 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 20000;
  ld_kernel<COLS><<<148, warps * 32>>>(out, iters);
  cudaEventRecord(e0);
  ld_kernel<COLS><<<148, warps * 32>>>(out, iters);
  cudaEventRecord(e1);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("CUDA error: %s\n", cudaGetErrorString(e));
    exit(1);
  }
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double bytes = (double)warps * 32 * COLS * 4 * iters;  // per SM
  printf("ld cols=%3d warps=%2d : %.3f ms, %.1f GB/s per SM (%.1f B/clk/SM at 1.9 GHz)\n", COLS, warps, ms,
         bytes / ms / 1e6, bytes / ms / 1e6 / 1.9);
  cudaFree(out);
}

static double run(int mode, int ld_warps, int n_mma, int ts, int mma_iters, int cols_per_ld = 128) {
  long long* out;
  cudaMalloc(&out, (256 + 148 * 1024) * sizeof(long long));
  const int smem = 64 * 1024;
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int threads = 128 + 32 * ld_warps;
  bench<<<148, threads, smem>>>(mode, ld_warps, n_mma, ts, mma_iters, cols_per_ld, out);
  bench<<<148, threads, smem>>>(mode, ld_warps, n_mma, ts, mma_iters, cols_per_ld, out);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("CUDA error: %s\n", cudaGetErrorString(e));
    exit(1);
  }
  long long h[148];
  cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  double s = 0;
  for (int i = 0; i < 148; ++i) s += (double)h[i];
  cudaFree(out);
  return s / 148;
}

int main() {
  printf("== (A0) tcgen05.ld 32x32b.x32 alone, CUDA-event timed\n");
  for (int w : {4, 8, 16}) run_ld<128>(w);
  for (int w : {4, 8, 16}) run_ld<64>(w);
  for (int w : {4, 8, 16}) run_ld<32>(w);
  printf("== (B) tcgen05.mma M=128 K=16 bf16, back to back from one thread\n");
  const int iters = 4000;
  for (int ts = 0; ts < 2; ++ts)
    for (int n : {16, 32, 48, 64, 96, 128, 192, 256}) {
      const double cyc = run(1, 0, n, ts, iters);
      printf("A in %s, N=%3d : %.1f cycles per MMA (math floor %.1f)\n", ts ? "TMEM" : "smem", n, cyc / (4.0 * iters),
             128.0 * n * 16 / 4096);
    }
  printf("== (B3) A staged into TMEM by tcgen05.cp (128x256b) in front of every TS MMA / in front of every four\n");
  for (int mode : {2, 3})
    for (int n : {64, 128, 192, 256}) {
      const double cyc = run(1, 0, n, mode, iters);
      printf("cp+TS mode %d, N=%3d : %.1f cycles per MMA (SS: %.0f, TS: %.0f)\n", mode, n, cyc / (4.0 * iters), 43.0 + n / 2, 10.0 + n / 2);
    }
  printf("== (B2) same, alternating between two / four accumulators\n");
  for (int alt : {1, 2})
    for (int n : {32, 64, 128}) {
      const double cyc = run(1, 0, n, 0, iters, alt);
      printf("A in smem, N=%3d, %d accumulators : %.1f cycles per MMA (math floor %.1f)\n", n, alt * 2, cyc / (4.0 * iters),
             128.0 * n * 16 / 4096);
    }
  return 0;
}
