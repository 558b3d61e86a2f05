#!/bin/bash
# 1 GPU: UNet parity tests against the oracle, then the default bench line with the per-launch profile of one forward
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_pipeline_gpu.py tests/test_determinism_gpu.py -q -m gpu -x -k "unet or determin or sampler_vs_oracle" 2>&1 | tail -4
timeout 500 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-baseline --profile-out gpurun_out/prof_q.json > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err
python - <<'PY'
import json, collections
d = json.loads(open("gpurun_out/bench_q.json").read().strip().splitlines()[-1])
print("fps %.3f e2e %.3f sampler_ms %.1f clocks %s" % (d["value"], d["e2e"]["value"], d["phases_ms_per_batch"]["sampler_unet_ddim"], d["clocks"]["sm_mhz"]), {k: round(v, 2) for k, v in d["breakdown_ms_per_unet_forward"].items()}, "pq", d["pq"]["pq"], d["ids_digest"]["first8"], "roofline", round(d["roofline"]["frac"], 3))
p = json.load(open("gpurun_out/prof_q.json"))
for r in p:
    if r["op"] == "gemm" and (r["shape"][3] == 4 or (r["shape"][3] == 9 and r["shape"][0] in (960, 3744, 14976) and r["shape"][2] in (2880, 5760, 11520) and r["shape"][1] != 320)):
        pass
c = collections.defaultdict(list)
for r in p:
    if r["op"] == "gemm" and r["shape"][3] in (4,): c[tuple(r["shape"])].append(r["ms"] * 1e3)
print("taps=4 launches:", {k: (len(v), round(sum(v) / len(v), 1)) for k, v in sorted(c.items())})
print("other:", [(r["op"], round(r["ms"] * 1e3, 1)) for r in p if r["op"] in ("upsample_nearest", "im2col3x3_s2")])
PY
tail -2 gpurun_out/bench_q.err
