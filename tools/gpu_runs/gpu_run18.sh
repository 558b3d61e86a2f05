#!/bin/bash
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/unet_profile2.json > gpurun_out/bench6.log 2>&1; tail -c 1300 gpurun_out/bench6.log
