#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k 'ln_fold or gemm' 2>&1 | tail -25
