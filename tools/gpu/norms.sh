#!/bin/bash
# (the LDM_* switches exist only in a diagnostic build: LDM_BUILD_DIAG=1 python -m video_latent_diffusion_panoptic_segmentation_b200.build, then rebuild the product library before committing numbers)
# 1 GPU: kernel checks of the norm kernels, then their device times at B = 8 and B = 1 with the cluster path off / on
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k 'groupnorm or layernorm' 2>&1 | tail -12
for mv in ${MAXVEC:-0 18 24 1000}; do
  for B in 8 1; do
    echo "== LDM_GN_CLUSTER_MAXVEC=$mv B=$B"
    LDM_GN_CLUSTER_MAXVEC=$mv timeout 300 python tools/bench_norms.py --B $B --json gpurun_out/norms_mv${mv}_B$B.json 2>&1 | grep -v "^$" | cut -c1-200
  done
done
