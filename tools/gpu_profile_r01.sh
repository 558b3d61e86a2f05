#!/bin/bash
# One gpurun call: GPU tests, kernel micro-timings, ncu launch list of the bench command, ncu --set full of the top kernels.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python tools/profile_kernels.py --iters 10 --json gpurun_out/kernels_events.json > gpurun_out/kernels_events.log 2>&1; echo "kernels rc=$?"
BENCH="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
$BENCH > gpurun_out/bench_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_bench.csv $BENCH > gpurun_out/ncu_bench.log 2>&1
echo "ncu launch list rc=$?"
PK="python tools/profile_kernels.py --iters 1 --only attn_L0,gemm1x1_res_L0,gemm_geglu_L0,conv3x3_L0,conv3x3_L3,groupnorm_L0"
$PK > gpurun_out/pk_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'flash_attn|gemm_tc|groupnorm' -o gpurun_out/prof_r01_top $PK > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
ls -la gpurun_out
