#!/bin/bash
for dbg in nowait nowait,altacc; do echo "== $dbg"; LDM_GEMM_PAIR=0 LDM_GEMM_DEBUG=$dbg timeout 200 python tools/profile_kernels.py --sweep --iters 5 2>&1 | grep -E '"N": 320, "K": 2880|"N": 640, "K": 5760' | grep -E '"block_n": (64|96|128),' | cut -c1-120; done
