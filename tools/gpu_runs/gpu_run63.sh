#!/bin/bash
# source-level profile of the QKV head-split GEMM at the 48x156 level
PK="python tools/profile_kernels.py --iters 1 --only gemm_qkv_L0"
$PK > gpurun_out/pk_qkv_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -c 2 -o gpurun_out/qkv_L0 $PK > gpurun_out/ncu_qkv.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/ncu_qkv.log
