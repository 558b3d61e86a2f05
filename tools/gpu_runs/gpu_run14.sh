#!/bin/bash
echo "== checks PT=1"; LDM_ATTN_PT=1 timeout 300 python tools/gpu_diag.py attn_40 2>&1 | cut -c1-330 | tail -4
for pt in 0 1; do for poly in 0 1 2; do echo "== attn pt=$pt poly=$poly"; LDM_ATTN_PT=$pt LDM_ATTN_POLY=$poly timeout 100 python tools/profile_kernels.py --iters 10 --only attn_L0 2>&1 | cut -c1-100; done; done
