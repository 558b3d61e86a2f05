#!/bin/bash
for poly in 9 1; do echo "== attn poly=$poly"; LDM_ATTN_POLY=$poly timeout 100 python tools/profile_kernels.py --iters 10 --only attn_L0 2>&1 | cut -c1-100; done
echo "== gemm with sleeps"; timeout 200 python tools/profile_kernels.py --iters 10 --only gemm,conv 2>&1 | cut -c1-100
