// gemm_tc_kernel<*, kEpiGegluLn>: see gemm_tc_kernel.cuh
#include "gemm_tc_inst.cuh"

LDM_GEMM_DEFINE_LAUNCHER(launch_gemm_geglu_ln, kEpiGegluLn)
