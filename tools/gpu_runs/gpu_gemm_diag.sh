#!/bin/bash
for dbg in none notma notma,nofence nofence nowait nowait,keepcommit; do
  echo "== debug=$dbg"
  LDM_GEMM_PAIR=0 LDM_GEMM_DEBUG=$dbg timeout 120 python tools/profile_kernels.py --iters 10 --only conv3x3_L0,conv3x3_L1,conv3x3_L2,gemm_ff2_L1 2>&1 | cut -c1-90
done
