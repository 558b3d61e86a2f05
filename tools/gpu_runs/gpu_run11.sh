#!/bin/bash
echo "== checks single"; LDM_GEMM_PAIR=0 timeout 300 python tools/gpu_diag.py gemm conv 2>&1 | grep -v '"ok": true' | tail -8 | cut -c1-400
echo "== checks pair"; LDM_GEMM_PAIR=1 timeout 300 python tools/gpu_diag.py gemm conv 2>&1 | grep -v '"ok": true' | tail -8 | cut -c1-400
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== sweep single"; LDM_GEMM_PAIR=0 timeout 300 python tools/profile_kernels.py --sweep --iters 10 --json gpurun_out/sweep_bn3.json > gpurun_out/sweep_bn3.log 2>&1
echo "== sweep pair"; LDM_GEMM_PAIR=1 timeout 300 python tools/profile_kernels.py --sweep --iters 10 --json gpurun_out/sweep_bn3_pair.json > gpurun_out/sweep_bn3_pair.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench4.log 2>&1; tail -c 1500 gpurun_out/bench4.log
PK="python tools/profile_kernels.py --iters 1 --only attn_L0,gemm1x1_res_L0,layernorm_L0"
$PK > gpurun_out/pk_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'flash_attn|gemm_tc|layernorm' -o gpurun_out/prof_r01d $PK > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
