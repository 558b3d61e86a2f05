#!/bin/bash
timeout 300 python tools/gpu_diag.py groupnorm 2>&1 | cut -c1-200 | tail -5
timeout 100 python tools/profile_kernels.py --iters 20 --only groupnorm 2>&1 | cut -c1-100
