#!/bin/bash
timeout 200 python tools/gpu_diag.py gemm_qkv 2>&1 | cut -c1-170 | tail -6
timeout 100 python tools/profile_kernels.py --iters 20 --only gemm_qkv 2>&1 | cut -c1-120 | tail -4
