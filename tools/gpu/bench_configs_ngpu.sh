#!/bin/bash
# N GPUs (gpurun --gpus N -- bash tools/gpu/bench_configs_ngpu.sh N): configs[2] strong split and configs[4]; at N = 8 also configs[3]
N=${1:-8}
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N "$@"; }
timeout 600 bash -c "$(declare -f run); N=$N; run --steps 3 --warmup 3 --config clip8_strong,k2 --no-cpu-baseline --no-gpu-baseline" > gpurun_out/bench_clip8_k2_n$N.json 2> gpurun_out/bench_clip8_k2_n$N.err
echo "clip8/k2 N=$N rc=$?"
if [ "$N" = "8" ] || [ -n "$SEQ" ]; then
  timeout 900 bash -c "$(declare -f run); N=$N; run --steps 1 --warmup 1 --config seq1101 --no-cpu-baseline --no-gpu-baseline" > gpurun_out/bench_seq1101_n$N.json 2> gpurun_out/bench_seq1101_n$N.err
  echo "seq1101 N=$N rc=$?"
fi
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_*_n*.json")):
    for l in open(f).read().strip().splitlines():
        d = json.loads(l)
        print(f, d["config"]["name"], d["n_gpus"], {k: d.get(k) for k in ("value", "e2e", "ms_per_step", "pq", "dvpq", "ids_digest", "clocks")})
PY
for f in gpurun_out/bench_*_n$N.err; do echo "== $f"; tail -n 4 $f; done
