#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_determinism_gpu.py -q -m gpu 2>&1 | tail -8
bash tools/gpu/sanitizer.sh
