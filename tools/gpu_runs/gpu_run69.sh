#!/bin/bash
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
timeout 200 python tools/attn_determinism.py 8 8 40 7488 12 2>&1 | grep -c identical
timeout 200 python tools/attn_determinism.py 8 8 80 1872 12 2>&1 | grep -c identical
PK="python tools/profile_kernels.py --iters 1 --only attn_L0,attn_L1,conv3x3_L0,conv3x3_L2,groupnorm_L0,layernorm_L0"
timeout 120 $PK 2>&1 | cut -c1-100
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'flash_attn|gemm_tc|gn_|layernorm' -o gpurun_out/r01h_top_kernels_b $PK > gpurun_out/ncu_full_b.log 2>&1
echo "ncu full rc=$?"; grep -c "passes" gpurun_out/ncu_full_b.log; tail -2 gpurun_out/ncu_full_b.log | cut -c1-150
