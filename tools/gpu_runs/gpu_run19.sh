#!/bin/bash
timeout 200 python tools/profile_kernels.py --iters 20 --only gemm,attn_L0 2>&1 | cut -c1-100
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench7.log 2>&1; tail -c 700 gpurun_out/bench7.log
