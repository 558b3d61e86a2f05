#!/bin/bash
timeout 300 python tools/gpu_diag.py attn 2>&1 | cut -c1-200 | tail -6
for poly in 2 1 3; do echo "== poly $poly"; LDM_ATTN_POLY=$poly timeout 100 python tools/profile_kernels.py --iters 20 --only attn_L0 2>&1 | cut -c1-120 | tail -1; done
LDM_ATTN40=0 timeout 100 python tools/profile_kernels.py --iters 20 --only attn_L0 2>&1 | cut -c1-120 | tail -1
timeout 60 tools/microbench/attn_trace_bin 20 | tail -24
