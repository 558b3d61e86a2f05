#!/bin/bash
timeout 60 python tools/gpu_diag.py attn_40 2>&1 | cut -c1-200 | tail -3
for poly in 2 3; do echo "== poly $poly"; LDM_ATTN_POLY=$poly timeout 60 python tools/profile_kernels.py --iters 20 --only attn_L0 2>&1 | cut -c1-120 | tail -1; done
