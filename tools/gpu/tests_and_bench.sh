#!/bin/bash
# 1 GPU: the whole GPU test suite, then the default bench line (configs[1]).  gpurun --timeout 1500 -- bash tools/gpu/tests_and_bench.sh
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps ${STEPS:-3} --warmup 3 > gpurun_out/bench_batch8.json 2> gpurun_out/bench_batch8.err
echo "bench rc=$?"; tail -c 3000 gpurun_out/bench_batch8.json; tail -5 gpurun_out/bench_batch8.err
