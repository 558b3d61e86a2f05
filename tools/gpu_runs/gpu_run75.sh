#!/bin/bash
timeout 900 python -m pytest tests -m gpu -q -x -k "joint_hist or vpq or cityscapes or compute_pq or dvpq or clip" 2>&1 | tail -4
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench19.json 2> gpurun_out/bench19.err; echo "bench rc=$?"; tail -3 gpurun_out/bench19.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench19.json'))
print(d['value'], d['e2e']['value'], d['ms_per_step'], d['phases_ms_per_step'], d['clocks'])
PY
