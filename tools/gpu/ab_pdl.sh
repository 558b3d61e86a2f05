#!/bin/bash
# (the LDM_* switches exist only in a diagnostic build: LDM_BUILD_DIAG=1 python -m video_latent_diffusion_panoptic_segmentation_b200.build, then rebuild the product library before committing numbers)
# 1 GPU: programmatic dependent launch off / on at 1, 2, 4 and 8 frames per batch (one box)
mkdir -p gpurun_out
for n in 1 2 4 8; do
  for pdl in 0 1; do
    LDM_PDL=$pdl timeout 400 python bench.py --steps 3 --warmup 3 --config clip8_strong --clip-frames $n --no-cpu-baseline --no-gpu-baseline > gpurun_out/pdl_${pdl}_b$n.json 2> gpurun_out/pdl.err
    python - "$pdl" "$n" <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/pdl_{sys.argv[1]}_b{sys.argv[2]}.json").read().strip().splitlines()[-1])
print("pdl", sys.argv[1], "frames", sys.argv[2], "fps %.3f ms_per_step %.1f sampler_ms %.1f" % (d["value"], d["ms_per_step"], d["phases_ms_per_batch"]["sampler_unet_ddim"]), d["ids_digest"]["all"], d["clocks"]["sm_mhz"])
PY
  done
done
