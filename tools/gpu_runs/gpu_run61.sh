#!/bin/bash
# cross-attention (kv_seq in the flash kernels, q-only / k|v head-split epilogue, UNet attn2 plan) + full suite
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "cross_attn or gemm_qkv or attn_" 2>&1 | tail -12
timeout 900 python -m pytest tests/test_pipeline_gpu.py -q -x -k "cross_attention" 2>&1 | tail -25
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
