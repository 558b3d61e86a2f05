#!/bin/bash
# (the LDM_* switches exist only in a diagnostic build: LDM_BUILD_DIAG=1 python -m video_latent_diffusion_panoptic_segmentation_b200.build, then rebuild the product library before committing numbers)
# 1 GPU: A/B of the default bench line on ONE box (box-to-box clock spread is ~2 %): baseline switches off / on, twice
mkdir -p gpurun_out
for rep in 1 2; do
  for cfg in "0 0" "1 18"; do
    set -- $cfg
    LDM_GEMM_SPLITK=$1 LDM_GN_CLUSTER_MAXVEC=$2 timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/ab_$1_$2_$rep.json 2> gpurun_out/ab.err
    python - "$1" "$2" "$rep" <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/ab_{sys.argv[1]}_{sys.argv[2]}_{sys.argv[3]}.json").read().strip().splitlines()[-1])
print("splitk", sys.argv[1], "gn_cluster_maxvec", sys.argv[2], "rep", sys.argv[3], "fps %.3f e2e %.3f sampler_ms %.1f clocks %s" % (d["value"], d["e2e"]["value"], d["phases_ms_per_batch"]["sampler_unet_ddim"], d["clocks"]), {k: round(v, 2) for k, v in d["breakdown_ms_per_unet_forward"].items()})
PY
  done
done
