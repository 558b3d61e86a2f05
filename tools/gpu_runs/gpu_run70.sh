#!/bin/bash
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "gemm_qkv or cross_attn or attn_" 2>&1 | tail -8
timeout 200 python tools/profile_kernels.py --iters 20 --only gemm_qkv 2>&1 | cut -c1-120 | tail -4
LDM_GEMM_QKV_TMA=0 timeout 200 python tools/profile_kernels.py --iters 20 --only gemm_qkv 2>&1 | cut -c1-120 | tail -4
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
timeout 300 python tools/determinism_diag.py 8 48 156 4 2>&1 | tail -3
