#!/bin/bash
timeout 120 python tools/gpu_diag.py groupnorm 2>&1 | cut -c1-220 | tail -4
for f in 0 1; do echo "== fused=$f"; LDM_GN_FUSED=$f timeout 100 python tools/profile_kernels.py --iters 20 --only groupnorm 2>&1 | cut -c1-120 | tail -6; done
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench14.json; python -c "
import json; d=json.load(open('gpurun_out/bench14.json')); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['breakdown_ms_per_unet_forward'])"
