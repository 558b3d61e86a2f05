#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python tools/profile_kernels.py --iters 10 --json gpurun_out/kernels_events3.json > gpurun_out/kernels_events3.log 2>&1; echo "kernels rc=$?"
cat gpurun_out/kernels_events3.log
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench3.log 2>&1; tail -c 1200 gpurun_out/bench3.log
PK="python tools/profile_kernels.py --iters 1 --only attn_L0,attn_L1,groupnorm_L0,gemm_geglu_L0,gemm1x1_res_L0"
$PK > gpurun_out/pk_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'flash_attn|gemm_tc|gn_' -o gpurun_out/prof_r01c $PK > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
