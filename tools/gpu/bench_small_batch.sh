#!/bin/bash
# 1 GPU: the per-GPU share of configs[2] at N = 8 / 4 / 2 (1 / 2 / 4 frames per batch) with a per-launch profile each
mkdir -p gpurun_out
for n in ${FRAMES:-1 2 4}; do
  timeout 400 python bench.py --steps 3 --warmup 3 --config clip8_strong --clip-frames $n --no-cpu-baseline --no-gpu-baseline --profile-out gpurun_out/forward_profile_b$n.json > gpurun_out/bench_clip_b$n.json 2> gpurun_out/bench_clip_b$n.err
  echo "frames=$n rc=$?"
done
python - <<'PY'
import json, glob, collections
for f in sorted(glob.glob("gpurun_out/bench_clip_b*.json")):
    for l in open(f).read().strip().splitlines():
        d = json.loads(l)
        print(f, {k: d.get(k) for k in ("value", "ms_per_step", "phases_ms_per_batch", "breakdown_ms_per_unet_forward", "roofline")})
PY
