#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench13.json; python -c "
import json; d=json.load(open('gpurun_out/bench13.json')); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['phases_ms_per_step'], d['breakdown_ms_per_unet_forward'], d['roofline']['frac'])"
