#!/bin/bash
# (the LDM_* switches exist only in a diagnostic build: LDM_BUILD_DIAG=1 python -m video_latent_diffusion_panoptic_segmentation_b200.build, then rebuild the product library before committing numbers)
# 1 GPU: split-K checks, then the tile-count-bound GEMM shapes with and without split-K at B = 8 and B = 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k 'gemm' 2>&1 | tail -15
for B in 8 1; do
  for sk in 0 1; do
    echo "== LDM_GEMM_SPLITK=$sk B=$B"
    LDM_GEMM_SPLITK=$sk timeout 300 python tools/bench_gemm_shapes.py --B $B --json gpurun_out/gemm_shapes_sk${sk}_B$B.json 2>&1 | tail -19 | cut -c1-220
  done
done
