#!/bin/bash
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for poly in 0 2 3; do echo "== attn poly=$poly"; LDM_ATTN_POLY=$poly timeout 100 python tools/profile_kernels.py --iters 10 --only attn_L0 2>&1 | cut -c1-100; done
echo "== attn others"; timeout 100 python tools/profile_kernels.py --iters 10 --only attn_L1,attn_L2 2>&1 | cut -c1-100
for pair in 0 1; do echo "== gemm pair=$pair"; LDM_GEMM_PAIR=$pair timeout 200 python tools/profile_kernels.py --iters 10 --only gemm,conv 2>&1 | cut -c1-100; done
echo "== pair correctness"; LDM_GEMM_PAIR=1 timeout 300 python tools/gpu_diag.py gemm conv 2>&1 | tail -3
echo "== sweep single"; LDM_GEMM_PAIR=0 timeout 300 python tools/profile_kernels.py --sweep --iters 10 --json gpurun_out/sweep_bn2.json > gpurun_out/sweep_bn2.log 2>&1
echo "== sweep pair"; LDM_GEMM_PAIR=1 timeout 300 python tools/profile_kernels.py --sweep --iters 10 --json gpurun_out/sweep_bn2_pair.json > gpurun_out/sweep_bn2_pair.log 2>&1
