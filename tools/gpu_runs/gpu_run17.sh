#!/bin/bash
echo "== checks DUAL=1"; LDM_ATTN_DUAL=1 timeout 300 python tools/gpu_diag.py attn_40 2>&1 | cut -c1-250 | tail -3
for dual in 0 1; do for poly in 9 0 1 2; do echo "== attn dual=$dual poly=$poly"; LDM_ATTN_DUAL=$dual LDM_ATTN_POLY=$poly timeout 100 python tools/profile_kernels.py --iters 10 --only attn_L0 2>&1 | cut -c1-100; done; done
