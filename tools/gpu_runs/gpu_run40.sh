#!/bin/bash
timeout 300 python tools/gpu_diag.py attn 2>&1 | cut -c1-250 | tail -7
for lv in L0 L1 L2; do timeout 100 python tools/profile_kernels.py --iters 20 --only attn_$lv 2>&1 | cut -c1-120 | tail -2; done
