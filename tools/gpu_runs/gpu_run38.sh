#!/bin/bash
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 3 --warmup 3 > gpurun_out/bench12.log 2>&1; tail -c 2500 gpurun_out/bench12.log
