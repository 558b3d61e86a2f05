#!/bin/bash
timeout 600 python -m pytest tests/test_pipeline_gpu.py -q -x -k "cross_attention" 2>&1 | tail -5
timeout 200 python tools/profile_kernels.py --iters 20 --only gemm_qkv,gemm1x1_res,gemm_geglu_L0 2>&1 | cut -c1-150 | tail -10
