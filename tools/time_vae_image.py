"""Device time of the RGB VAE encoder at the bench shape: B frames of 384x1248, CUDA events."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import vae_image_oracle as VO  # noqa: E402  (random-init weights only)
from video_latent_diffusion_panoptic_segmentation_b200.ldmseg.models import GeneralVAEImage  # noqa: E402
from video_latent_diffusion_panoptic_segmentation_b200 import _lib as L  # noqa: E402

B = int(os.environ.get("VAE_B", "8"))
H, W = 384, 1248
m = GeneralVAEImage.from_pretrained(state_dict=VO.build_vae_image(seed=0).state_dict(), device="cuda")
x = torch.rand((B, 3, H, W), device="cuda")
for _ in range(2):
    m.encode_moments(x, scale=2.0, shift=-1.0)
torch.cuda.synchronize()
st = m._plans[(B, H, W)]
n0 = L.launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
iters = 5
e0.record()
for _ in range(iters):
    m.encode_moments(x, scale=2.0, shift=-1.0)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
launches = (L.launch_count() - n0) // iters
# per-op times
evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(st.plan) + 1)]
evs[0].record()
for i, (fn, a, k) in enumerate(st.plan):
    fn(*a, **k)
    evs[i + 1].record()
torch.cuda.synchronize()
by = {}
for i, (fn, a, k) in enumerate(st.plan):
    name = getattr(fn, "__name__", "op")
    if name == "gemm":
        name = "gemm_conv3x3" if k.get("taps", 1) == 9 else "gemm_1x1"
    by[name] = by.get(name, 0.0) + evs[i].elapsed_time(evs[i + 1])
print(json.dumps({"what": "GeneralVAEImage.encode_moments", "frames": B, "size": [H, W], "ms": round(ms, 3),
                  "frames_per_s": round(B / ms * 1e3, 2), "launches": launches,
                  "arena_GB": round(st.arena_bytes / 2 ** 30, 2),
                  "by_op_ms": {k: round(v, 3) for k, v in sorted(by.items(), key=lambda kv: -kv[1])}}))
