// gemm_tc_kernel<*, kEpiQkvLn>: see gemm_tc_kernel.cuh
#include "gemm_tc_inst.cuh"

LDM_GEMM_DEFINE_LAUNCHER(launch_gemm_qkv_ln, kEpiQkvLn)
