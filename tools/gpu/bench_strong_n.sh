#!/bin/bash
# N GPUs (gpurun --gpus N -- bash tools/gpu/bench_strong_n.sh N): configs[2] strong split at N, and the default weak line (its
# ids_digest.first8 must equal the 1-GPU line's)
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 3 --warmup 3 --config clip8_strong,batch8 --no-cpu-baseline --no-gpu-baseline > gpurun_out/bench_clip8_batch8_n$N.json 2> gpurun_out/bench_clip8_batch8_n$N.err
echo "rc=$?"
python - $N <<'PY'
import json, sys
for l in open(f"gpurun_out/bench_clip8_batch8_n{sys.argv[1]}.json").read().strip().splitlines():
    d = json.loads(l)
    print(d["config"]["name"], d["n_gpus"], {k: d.get(k) for k in ("value", "e2e", "ms_per_step", "pq", "dvpq", "clocks")}, d["ids_digest"]["first8"], d["ids_digest"]["all"])
PY
tail -n 3 gpurun_out/bench_clip8_batch8_n$N.err
