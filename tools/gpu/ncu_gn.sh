#!/bin/bash
# 1 GPU: ncu --set full of the GroupNorm launch of one shape (HW,C) at batch B
mkdir -p gpurun_out
SHAPE=${SHAPE:-120,1280}; B=${B:-8}
timeout 300 python tools/bench_norms.py --B $B --only $SHAPE 2>&1 | tail -3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gn_ --launch-skip 1 --launch-count 1 -o gpurun_out/gn_${SHAPE/,/_}_B$B -f python tools/bench_norms.py --B $B --only $SHAPE > gpurun_out/ncu_gn.log 2>&1
tail -3 gpurun_out/ncu_gn.log
