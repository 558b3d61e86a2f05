#!/bin/bash
# 1 GPU: one frame per batch (the per-GPU work of configs[2] at 8 GPUs): bench line + per-launch profile of one forward
mkdir -p gpurun_out
B=${B:-1}
timeout 400 python bench.py --steps 3 --warmup 3 --config clip8_strong --clip-frames $B --no-cpu-baseline --no-gpu-baseline --profile-out gpurun_out/prof_b$B.json > gpurun_out/bench_b$B.json 2> gpurun_out/bench_b$B.err
python - $B <<'PY'
import json, sys, collections
b = sys.argv[1]
d = json.loads(open(f"gpurun_out/bench_b{b}.json").read().strip().splitlines()[-1])
print("frames", b, "fps %.3f ms_per_step %.1f sampler_ms %.1f" % (d["value"], d["ms_per_step"], d["phases_ms_per_batch"]["sampler_unet_ddim"]), d["clocks"]["sm_mhz"], {k: round(v, 3) for k, v in d["breakdown_ms_per_unet_forward"].items()})
p = json.load(open(f"gpurun_out/prof_b{b}.json"))
print("launches", len(p), "sum ms %.3f" % sum(r["ms"] for r in p))
c = collections.defaultdict(list)
for r in p:
    key = (r["op"],) + tuple(r.get("shape", ()))
    c[key].append(r["ms"] * 1e3)
for k, v in sorted(c.items(), key=lambda kv: -sum(kv[1]))[:40]:
    print("%7.1f us total  n=%3d avg %6.1f  %s" % (sum(v), len(v), sum(v) / len(v), k))
PY
tail -2 gpurun_out/bench_b$B.err
