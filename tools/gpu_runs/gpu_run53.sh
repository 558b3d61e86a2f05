#!/bin/bash
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 100 python tools/profile_kernels.py --iters 20 --only groupnorm 2>&1 | cut -c1-120 | tail -5
python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench15.json; python -c "
import json; d=json.load(open('gpurun_out/bench15.json')); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['breakdown_ms_per_unet_forward'])"
