#!/bin/bash
# 8 GPUs (gpurun --gpus 8 -- bash tools/gpu/bench_seq1101_8gpu.sh): configs[3], the 1 101-frame sequence end to end
N=${1:-8}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 1 --warmup 1 --config seq1101 --no-cpu-baseline --no-gpu-baseline > gpurun_out/bench_seq1101_n$N.json 2> gpurun_out/bench_seq1101_n$N.err
echo "rc=$?"
python - $N <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/bench_seq1101_n{sys.argv[1]}.json").read().strip().splitlines()[-1])
print(d["config"]["name"], d["n_gpus"], {k: d.get(k) for k in ("value", "e2e", "ms_per_step", "pq", "dvpq", "clocks", "warmup", "steps")}, d["ids_digest"]["first8"])
PY
tail -n 3 gpurun_out/bench_seq1101_n$N.err
