#!/bin/bash
timeout 300 python tools/gpu_diag.py attn_80 2>&1 | cut -c1-250 | tail -2
for v in 0 1; do echo "== d80 variant $v"; LDM_ATTN_D80=$v timeout 100 python tools/profile_kernels.py --iters 20 --only attn_L1 2>&1 | cut -c1-100; done
