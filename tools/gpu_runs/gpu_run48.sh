#!/bin/bash
timeout 120 python tools/gpu_diag.py groupnorm 2>&1 | cut -c1-120 | tail -4
timeout 100 python tools/profile_kernels.py --iters 20 --only groupnorm,layernorm 2>&1 | cut -c1-120 | tail -9
