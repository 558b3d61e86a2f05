#!/bin/bash
# 1 GPU: LayerNorm fold off / on, one box, with the per-launch profile of each
mkdir -p gpurun_out
for rep in $(seq 1 ${REPS:-2}); do
for f in nofold fold; do
  flag=""; [ $f = fold ] && flag="--ln-fold"
  timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-baseline $flag --profile-out gpurun_out/prof_$f.json > gpurun_out/ab_$f.json 2> gpurun_out/ab.err
  python - $f <<'PY'
import json, sys, collections
d = json.loads(open(f"gpurun_out/ab_{sys.argv[1]}.json").read().strip().splitlines()[-1])
print(sys.argv[1], "fps %.3f sampler_ms %.1f clocks %s" % (d["value"], d["phases_ms_per_batch"]["sampler_unet_ddim"], d["clocks"]["sm_mhz"]), {k: round(v, 2) for k, v in d["breakdown_ms_per_unet_forward"].items()}, d["pq"]["pq"])
p = json.load(open(f"gpurun_out/prof_{sys.argv[1]}.json"))
c = collections.defaultdict(list)
for r in p:
    if r["op"] == "gemm" and r["shape"][2] in (320, 640, 1280) and r["shape"][3] == 1: c[tuple(r["shape"])].append(r["ms"] * 1e3)
print("   ", {k: (len(v), round(sum(v) / len(v), 1)) for k, v in sorted(c.items())})
PY
done
done
