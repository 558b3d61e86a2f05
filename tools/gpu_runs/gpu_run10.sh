#!/bin/bash
echo "== checks single"; LDM_GEMM_PAIR=0 timeout 300 python tools/gpu_diag.py gemm conv 2>&1 | grep -v '"ok": true' | tail -8
echo "== checks pair"; LDM_GEMM_PAIR=1 timeout 300 python tools/gpu_diag.py gemm conv 2>&1 | grep -v '"ok": true' | tail -8
for st in 0 1; do for pair in 0 1; do echo "== staged=$st pair=$pair"; LDM_GEMM_STAGED=$st LDM_GEMM_PAIR=$pair timeout 200 python tools/profile_kernels.py --iters 10 --only gemm1x1,gemm_geglu,gemm_ff2,conv3x3_L0,conv3x3_L2 2>&1 | cut -c1-100; done; done
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
