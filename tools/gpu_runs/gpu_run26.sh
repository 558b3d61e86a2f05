#!/bin/bash
for rm in 32 100000; do echo "== resmma<=$rm"; LDM_GEMM_RESMMA=$rm timeout 200 python tools/profile_kernels.py --iters 20 --only gemm_ff2,conv3x3_L0,conv3x3_L1,conv3x3_L2,conv3x3_L3 2>&1 | cut -c1-100; done
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench9.log 2>&1; tail -c 700 gpurun_out/bench9.log
