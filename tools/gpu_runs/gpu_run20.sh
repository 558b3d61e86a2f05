#!/bin/bash
echo "== checks single"; LDM_GEMM_PAIR=0 timeout 300 python tools/gpu_diag.py gemm conv 2>&1 | grep -v '"ok": true' | tail -4 | cut -c1-300
echo "== checks pair"; LDM_GEMM_PAIR=1 timeout 300 python tools/gpu_diag.py gemm conv 2>&1 | grep -v '"ok": true' | tail -4 | cut -c1-300
for pair in 0 1 -1; do echo "== pair=$pair"; LDM_GEMM_PAIR=$pair timeout 200 python tools/profile_kernels.py --iters 20 --only gemm,conv3x3_L0,conv3x3_L2 2>&1 | cut -c1-100; done
export LDM_GEMM_PAIR=0
PK="python tools/profile_kernels.py --iters 1 --only gemm1x1_res_L0,gemm_ff2_L0,gemm_qkv_L0"
$PK > gpurun_out/pk_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'gemm_tc' -o gpurun_out/prof_r01e $PK > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
