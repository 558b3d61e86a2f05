#!/bin/bash
timeout 300 python tools/gpu_diag.py attn 2>&1 | cut -c1-200 | tail -6
for poly in 0 2 3 4; do echo "== poly $poly"; LDM_ATTN_POLY=$poly timeout 100 python tools/profile_kernels.py --iters 20 --only attn_L0 2>&1 | cut -c1-120 | tail -1; done
for lv in L1 L2; do timeout 100 python tools/profile_kernels.py --iters 20 --only attn_$lv 2>&1 | cut -c1-120 | tail -1; done
