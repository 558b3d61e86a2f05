#!/bin/bash
PK="python tools/profile_kernels.py --iters 1 --only attn_L0"
$PK > gpurun_out/pk_plain.log 2>&1 || exit 1
LDM_ATTN_POLY=2 ncu --set full --clock-control none --import-source on -k regex:'flash_attn' -c 1 -f -o gpurun_out/attn_p2 $PK > gpurun_out/ncu_attn.log 2>&1
echo "ncu rc=$?"
LDM_ATTN_POLY=0 ncu --set full --clock-control none --import-source on -k regex:'flash_attn' -c 1 -f -o gpurun_out/attn_p0 $PK > gpurun_out/ncu_attn0.log 2>&1
echo "ncu rc=$?"
