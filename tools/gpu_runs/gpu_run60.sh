#!/bin/bash
# persistent / prefetching LayerNorm: parity, then its time at the three levels next to GroupNorm
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "layernorm" 2>&1 | tail -5
timeout 200 python tools/profile_kernels.py --iters 20 --only layernorm,groupnorm_L0 2>&1 | cut -c1-150 | tail -8
