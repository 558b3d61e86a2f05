#!/bin/bash
for rm in 0 1; do LDM_GEMM_RESMMA=$rm LDM_GEMM_PAIR=0 python tools/gemm_dram_probe.py 2>&1 | tail -1; done
echo "== checks single"; LDM_GEMM_PAIR=0 timeout 300 python tools/gpu_diag.py gemm conv 2>&1 | grep -v '"ok": true' | tail -4 | cut -c1-400
echo "== checks pair"; LDM_GEMM_PAIR=1 timeout 300 python tools/gpu_diag.py gemm conv 2>&1 | grep -v '"ok": true' | tail -4 | cut -c1-400
for rm in 0 1; do echo "== resmma=$rm"; LDM_GEMM_RESMMA=$rm timeout 200 python tools/profile_kernels.py --iters 20 --only gemm1x1,gemm_ff2 2>&1 | cut -c1-100; done
