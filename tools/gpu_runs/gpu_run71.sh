#!/bin/bash
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "gemm_qkv or cross_attn" 2>&1 | tail -3
timeout 200 python tools/profile_kernels.py --iters 20 --only gemm_qkv 2>&1 | cut -c1-120 | tail -4
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench17.json 2> gpurun_out/bench17.err; echo "bench rc=$?"; tail -3 gpurun_out/bench17.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench17.json'))
print(d['value'], d['e2e']['value'], d['ms_per_step'], d['phases_ms_per_step'], d['clocks'])
print(d['breakdown_ms_per_unet_forward'], d['roofline']['frac'], d['roofline']['whole_job_frac'], d.get('hbm_kernels',{}).get('frac'))
PY
