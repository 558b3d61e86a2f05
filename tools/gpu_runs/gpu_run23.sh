#!/bin/bash
echo "== checks single"; LDM_GEMM_PAIR=0 timeout 300 python tools/gpu_diag.py gemm conv 2>&1 | grep -v '"ok": true' | tail -4 | cut -c1-300
echo "== checks pair"; LDM_GEMM_PAIR=1 timeout 300 python tools/gpu_diag.py gemm conv 2>&1 | grep -v '"ok": true' | tail -4 | cut -c1-300
echo "== timing"; timeout 200 python tools/profile_kernels.py --iters 20 --only gemm,conv 2>&1 | cut -c1-100
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench8.log 2>&1; tail -c 700 gpurun_out/bench8.log
