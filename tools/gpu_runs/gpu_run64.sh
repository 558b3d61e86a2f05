#!/bin/bash
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench16.json 2> gpurun_out/bench16.err; echo "bench rc=$?"; tail -3 gpurun_out/bench16.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench16.json'))
print(d['value'], d['e2e']['value'], d['ms_per_step'], d['phases_ms_per_step'], d['dvpq'], d['pq'], d['clocks'])
print(d['breakdown_ms_per_unet_forward'], d['roofline']['frac'], d.get('hbm_kernels',{}).get('frac'))
PY
