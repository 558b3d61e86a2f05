#!/bin/bash
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "gemm_qkv or cross_attn" 2>&1 | tail -2
timeout 200 python tools/determinism_diag.py 8 48 156 3 2>&1 | tail -2
timeout 100 python tools/profile_kernels.py --iters 20 --only gemm_qkv_L0 2>&1 | cut -c1-100 | tail -1
