#!/bin/bash
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
LDM_GEMM_PAIR=1 timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/unet_profile3.json > gpurun_out/bench10.log 2>&1; tail -c 800 gpurun_out/bench10.log
