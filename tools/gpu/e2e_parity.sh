#!/bin/bash
# 1 GPU: the end-to-end PQ / DVPQ parity tests only (reports land in gpurun_out/e2e_parity_*.json)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_e2e_parity_gpu.py -q -m gpu 2>&1 | tail -30 | tee gpurun_out/pytest_e2e.log
