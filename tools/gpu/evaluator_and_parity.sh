#!/bin/bash
# 1 GPU: evaluator / DVPQ / entry-point tests, the end-to-end parity tests, then the default bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_pipeline_gpu.py -q -x -m gpu -k "evaluator or joint_hist or compute_pq or dvpq or main_ldm or golden" 2>&1 | tail -15 | tee gpurun_out/pytest_eval.log
timeout 1200 python -m pytest tests/test_e2e_parity_gpu.py -q -m gpu 2>&1 | grep -E "^\{|passed|failed|Error" | cut -c1-3000 | tee gpurun_out/pytest_e2e.log
timeout 600 python bench.py --steps ${STEPS:-3} --warmup 3 > gpurun_out/bench_batch8.json 2> gpurun_out/bench_batch8.err
echo "bench rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/bench_batch8.json').read().strip().splitlines()[-1])
for k in ('value','e2e','pq','dvpq','phases_ms_per_batch','roofline','attention_tflops','hbm_kernels','torch_gpu_baseline','cpu_baseline','ids_digest','clocks'): print(k, d.get(k))
"; tail -5 gpurun_out/bench_batch8.err
