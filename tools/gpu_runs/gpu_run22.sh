#!/bin/bash
for pf in 0 1; do echo "== prefetch=$pf"; LDM_GEMM_PREFETCH=$pf timeout 100 python tools/profile_kernels.py --iters 20 --only gemm 2>&1 | cut -c1-100; done
echo "== checks"; timeout 300 python tools/gpu_diag.py gemm conv 2>&1 | grep -v '"ok": true' | tail -4 | cut -c1-300
