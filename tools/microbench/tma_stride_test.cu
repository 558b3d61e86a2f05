// Does a TMA tensor map with elementStrides = 2 in the pixel dimensions (a) load every other pixel -- the A operand of a
// stride-2 convolution as an implicit GEMM -- and (b) STORE to every other pixel -- the sub-pixel classes of a
// nearest-upsample + conv written into the dense output? Prints what the hardware does (box extents are in traversal
// space: boxDim = stride * pixels).   nvcc -gencode arch=compute_100a,code=sm_100a -o tma_stride_test tma_stride_test.cu -lcuda
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr int C = 64, W = 16, H = 8, BW = 8, BH = 4;

__global__ void k(const __grid_constant__ CUtensorMap tmIn, const __grid_constant__ CUtensorMap tmOut,
                  __nv_bfloat16* dump, int* status, int x0, int y0, int sx, int sy) {
  __shared__ __align__(1024) __nv_bfloat16 tile[BH * BW * C];
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < BH * BW * C; i += blockDim.x) tile[i] = __float2bfloat16(-1.f);
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(BH * BW * C * 2) : "memory");
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(s32(tile)), "l"(reinterpret_cast<uint64_t>(&tmIn)), "r"(s32(&bar)), "r"(0), "r"(x0), "r"(y0), "r"(0) : "memory");
  }
  int ok = 0;
  const long long t0 = clock64();
  while (!ok && clock64() - t0 < 400000000LL) {
    uint32_t done;
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done) : "r"(s32(&bar)), "r"(0) : "memory");
    ok = done;
  }
  if (threadIdx.x == 0) status[0] = ok;
  __syncthreads();
  for (int i = threadIdx.x; i < BH * BW * C; i += blockDim.x) dump[i] = tile[i];
  if (!ok) return;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(&tmOut)), "r"(s32(tile)), "r"(0), "r"(sx), "r"(sy), "r"(0) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

typedef CUresult (*enc_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                           const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                           CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  setvbuf(stdout, nullptr, _IONBF, 0);
  enc_fn enc = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q));
  __nv_bfloat16 *in, *out, *dump;
  int* status;
  CK(cudaMalloc(&in, H * W * C * 2)); CK(cudaMalloc(&out, H * W * C * 2)); CK(cudaMalloc(&dump, BH * BW * C * 2));
  CK(cudaMalloc(&status, 4));
  __nv_bfloat16 h[H * W * C];
  for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) for (int c = 0; c < C; ++c) h[(y * W + x) * C + c] = __float2bfloat16((float)(y * W + x + 1));
  CK(cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice));
  for (int variant = 0; variant < 2; ++variant) {
    // variant 0: boxDim in traversal space (2*BW, 2*BH); variant 1: boxDim = elements delivered (BW, BH)
    CK(cudaMemset(out, 0, H * W * C * 2));
    CUtensorMap tmIn, tmOut;
    cuuint64_t gd[4] = {C, W, H, 1}, gs[3] = {C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t bx[4] = {C, (cuuint32_t)(variant == 0 ? 2 * BW : BW), (cuuint32_t)(variant == 0 ? 2 * BH : BH), 1}, es[4] = {1, 2, 2, 1};
    CUresult r1 = enc(&tmIn, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, in, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult r2 = enc(&tmOut, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, out, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("variant %d: encode rc %d %d\n", variant, (int)r1, (int)r2);
    if (r1 || r2) continue;
    k<<<1, 128>>>(tmIn, tmOut, dump, status, -1, -1, 1, 1);
    cudaError_t e = cudaDeviceSynchronize();
    printf("  kernel: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    int st; __nv_bfloat16 d[BH * BW * C], o[H * W * C];
    CK(cudaMemcpy(&st, status, 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(d, dump, sizeof(d), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(o, out, sizeof(o), cudaMemcpyDeviceToHost));
    printf("  load completed: %d\n  tile (channel 0 of each pixel; expected in(2i-1, 2j-1) = (2i-1)*16 + 2j-1 + 1, 0 outside):\n", st);
    int bad = 0;
    for (int i = 0; i < BH; ++i) {
      printf("   ");
      for (int j = 0; j < BW; ++j) {
        // SWIZZLE_128B: 16-byte chunk c16 of row r sits at chunk c16 ^ (r & 7); channel 0 is chunk 0
        const int r = i * BW + j;
        const float v = __bfloat162float(d[r * C + ((0 ^ (r & 7)) * 8)]);
        const int yy = 2 * i - 1, xx = 2 * j - 1;
        const float want = (yy < 0 || xx < 0 || yy >= H || xx >= W) ? 0.f : (float)(yy * W + xx + 1);
        bad += v != want;
        printf(" %5.0f", v);
      }
      printf("\n");
    }
    printf("  load mismatches: %d\n", bad);
    if (!st) continue;
    printf("  out (channel 0; expected tile(i, j) at (2i+1, 2j+1), 0 elsewhere):\n");
    int sbad = 0;
    for (int y = 0; y < H; ++y) {
      printf("   ");
      for (int x = 0; x < W; ++x) {
        const float v = __bfloat162float(o[(y * W + x) * C]);
        float want = 0.f;
        if ((y & 1) && (x & 1)) {
          const int i = y / 2, j = x / 2, yy = 2 * i - 1, xx = 2 * j - 1;
          want = (yy < 0 || xx < 0) ? 0.f : (float)(yy * W + xx + 1);
        }
        sbad += v != want;
        printf(" %4.0f", v);
      }
      printf("\n");
    }
    printf("  store mismatches: %d\n", sbad);
  }
  return 0;
}
