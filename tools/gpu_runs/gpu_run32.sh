#!/bin/bash
timeout 300 python tools/gpu_diag.py gemm_qkv 2>&1 | cut -c1-200 | tail -3
timeout 100 python tools/profile_kernels.py --iters 20 --only gemm_qkv 2>&1 | cut -c1-100
