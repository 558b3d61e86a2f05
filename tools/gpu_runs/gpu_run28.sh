#!/bin/bash
for i in 1 2 3; do timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q 2>&1 | grep -E "AssertionError|passed|failed" | cut -c1-700 | head -4; done
