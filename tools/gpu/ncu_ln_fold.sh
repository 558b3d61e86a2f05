#!/bin/bash
# 1 GPU: graph times of the folded / unfolded GEMMs around a LayerNorm, then ncu --set full of the folded and the plain
# GEGLU launch at the 48x156 level (kernel launches 2 and 3 of the --ncu run)
mkdir -p gpurun_out
timeout 300 python tools/bench_ln_fold.py --json gpurun_out/ln_fold_B8.json 2>&1 | tail -5
K=${K:-geglu}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc --launch-skip 1 --launch-count 2 -o gpurun_out/ln_fold_$K -f python tools/bench_ln_fold.py --ncu $K > gpurun_out/ncu_ln_fold.log 2>&1
tail -3 gpurun_out/ncu_ln_fold.log
