// gemm_tc_kernel<*, kEpiSplit>: the staged epilogue behind a split-K main loop, see gemm_tc_kernel.cuh
#include "gemm_tc_inst.cuh"

LDM_GEMM_DEFINE_LAUNCHER(launch_gemm_split, kEpiSplit)
