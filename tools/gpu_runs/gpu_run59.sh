#!/bin/bash
# new RGB VAE encoder path + self-conditioned sampler: parity tests, then the encoder's time at the bench shape
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "vae_helpers or im2col or conv_small_cin" 2>&1 | tail -15
timeout 900 python -m pytest tests/test_pipeline_gpu.py -q -x -k "vae_image or encode_inputs or self_condition or seg_encoder" 2>&1 | tail -25
timeout 300 python tools/time_vae_image.py 2>&1 | tail -12
