#!/bin/bash
# 8 GPUs (gpurun --gpus 8 -- bash tools/gpu/bench_8gpu_final.sh): configs[2] strong split (one frame per GPU), configs[4]
# (768x2496, two frames per GPU) and the default weak line (ids_digest.first8 must equal the 1-GPU line's)
N=${1:-8}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 3 --warmup 3 --config clip8_strong,k2,batch8 --no-cpu-baseline --no-gpu-baseline > gpurun_out/bench_final_n$N.json 2> gpurun_out/bench_final_n$N.err
echo "rc=$?"
python - $N <<'PY'
import json, sys
for l in open(f"gpurun_out/bench_final_n{sys.argv[1]}.json").read().strip().splitlines():
    d = json.loads(l)
    print(d["config"]["name"], d["n_gpus"], {k: d.get(k) for k in ("value", "e2e", "ms_per_step", "pq", "dvpq", "clocks")}, d["ids_digest"]["first8"], d["ids_digest"]["all"])
PY
tail -n 3 gpurun_out/bench_final_n$N.err
