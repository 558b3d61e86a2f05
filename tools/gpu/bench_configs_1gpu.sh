#!/bin/bash
# 1 GPU: the default bench line (configs[1]) with its per-launch profile, then configs[2] (N = 1 leg of the strong split) and configs[4]
mkdir -p gpurun_out
timeout 700 python bench.py --steps ${STEPS:-3} --warmup 3 --profile-out gpurun_out/forward_profile.json > gpurun_out/bench_batch8.json 2> gpurun_out/bench_batch8.err
echo "batch8 rc=$?"
timeout 500 python bench.py --steps 3 --warmup 3 --config clip8_strong,k2 --no-cpu-baseline --no-gpu-baseline > gpurun_out/bench_clip8_k2.json 2> gpurun_out/bench_clip8_k2.err
echo "clip8/k2 rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/bench_batch8.json", "gpurun_out/bench_clip8_k2.json"):
    for l in open(f).read().strip().splitlines():
        d = json.loads(l)
        print(d["config"]["name"], {k: d.get(k) for k in ("value", "e2e", "ms_per_step", "pq", "dvpq", "phases_ms_per_batch", "roofline", "attention_tflops", "hbm_kernels", "breakdown_ms_per_unet_forward", "torch_gpu_baseline", "cpu_baseline", "clocks")})
PY
tail -3 gpurun_out/*.err
