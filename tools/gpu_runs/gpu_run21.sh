#!/bin/bash
export LDM_GEMM_PAIR=0
for dbg in none notma nomma; do echo "== $dbg"; LDM_GEMM_DEBUG=$dbg timeout 100 python tools/profile_kernels.py --iters 20 --only gemm1x1_res_L0,gemm_ff2_L0,gemm_qkv_L0,gemm1x1_res_L1 2>&1 | cut -c1-100; done
echo "== staged off"; LDM_GEMM_STAGED=0 timeout 100 python tools/profile_kernels.py --iters 20 --only gemm1x1_res_L0,gemm_ff2_L0,gemm1x1_res_L1 2>&1 | cut -c1-100
