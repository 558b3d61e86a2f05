#!/bin/bash
echo "== checks"; timeout 300 python tools/gpu_diag.py attn layernorm 2>&1 | cut -c1-330 | tail -12
for split in 1 2; do for poly in 0 1 2; do echo "== attn split=$split poly=$poly"; LDM_ATTN_SPLIT=$split LDM_ATTN_POLY=$poly timeout 100 python tools/profile_kernels.py --iters 10 --only attn_L0 2>&1 | cut -c1-100; done; done
echo "== others"; timeout 100 python tools/profile_kernels.py --iters 10 --only attn_L1,attn_L2,layernorm,gemm,conv 2>&1 | cut -c1-100
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench5.log 2>&1; tail -c 900 gpurun_out/bench5.log
