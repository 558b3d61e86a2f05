// One translation unit per epilogue mode instantiates the kernel (single CTA and CTA pair) and exports its launcher.
#pragma once
#include "gemm_tc_kernel.cuh"

namespace ldm_gemm {

template <int kEpi>
cudaError_t launch_epi(bool pair, int grid, int smem_bytes, cudaStream_t stream, const CUtensorMap& tmA1,
                       const CUtensorMap& tmA2, const CUtensorMap& tmB, const CUtensorMap& tmO, const CUtensorMap& tmR,
                       const CUtensorMap& tmE, const GemmParams& p) {
  // (per launch, not cached: the attribute belongs to the current device's copy of the function and a process may
  // drive several devices; the call is a few hundred nanoseconds and is not a stream operation)
  cudaError_t e = pair ? cudaFuncSetAttribute(gemm_tc_kernel<true, kEpi>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget)
                       : cudaFuncSetAttribute(gemm_tc_kernel<false, kEpi>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
  if (e != cudaSuccess) return e;
  if (pair)
    return ldm_host::launch_pdl(gemm_tc_kernel<true, kEpi>, dim3(grid), dim3(kThreads), (size_t)smem_bytes, stream, 2, tmA1,
                                tmA2, tmB, tmO, tmR, tmE, p);
  return ldm_host::launch_pdl(gemm_tc_kernel<false, kEpi>, dim3(grid), dim3(kThreads), (size_t)smem_bytes, stream, 1, tmA1,
                              tmA2, tmB, tmO, tmR, tmE, p);
}

}  // namespace ldm_gemm

#define LDM_GEMM_DEFINE_LAUNCHER(NAME, EPI)                                                                          \
  namespace ldm_gemm {                                                                                               \
  cudaError_t NAME(bool pair, int grid, int smem_bytes, cudaStream_t stream, const CUtensorMap& tmA1,                \
                   const CUtensorMap& tmA2, const CUtensorMap& tmB, const CUtensorMap& tmO, const CUtensorMap& tmR,  \
                   const CUtensorMap& tmE, const GemmParams& p) {                                                    \
    return launch_epi<EPI>(pair, grid, smem_bytes, stream, tmA1, tmA2, tmB, tmO, tmR, tmE, p);                       \
  }                                                                                                                  \
  }
