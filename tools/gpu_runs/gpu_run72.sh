#!/bin/bash
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench18.json 2> gpurun_out/bench18.err; echo "bench rc=$?"; tail -3 gpurun_out/bench18.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench18.json'))
print(d['value'], d['e2e']['value'], d['hbm_tail_kernels'], d['rgb_vae_encode_ms_per_batch'])
PY
