"""Device time of one UNet forward at the configs[4] frame size (768x2496 -> latent 96x312, 29 952 tokens at the first
level), per op class, CUDA events. Usage: python tools/time_unet_k2.py [B]"""
import copy
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from video_latent_diffusion_panoptic_segmentation_b200 import _lib as L  # noqa: E402
from video_latent_diffusion_panoptic_segmentation_b200.tools import main_ldm  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
h, w = 96, 312
dev = torch.device("cuda", 0)
L.lib()
p = copy.deepcopy(main_ldm.BASE)
vae, unet, sched = main_ldm.build_models(p, dev, seed=0)
st = unet._get_plan(B, h, w, 8)
st.sample.copy_(torch.randn((B, 8, h, w), device=dev))
st.timestep.fill_(499)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(2):
    st.graph.replay()
torch.cuda.synchronize()
e0.record()
for _ in range(3):
    st.graph.replay()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
prof = unet.profile_plan(st, iters=2)
by = {}
for r in prof:
    d = by.setdefault(r["op"], {"ms": 0.0, "flops": 0})
    d["ms"] += r["ms"]; d["flops"] += r["flops"]
flop_frame = 11294.0e9  # BASELINE.md section 3, one DDIM step of one 768x2496 frame
print(json.dumps({"what": "UNet forward, 768x2496 frames (latent 96x312)", "frames": B, "ms_per_forward": round(ms, 2),
                  "algorithmic_tflops": round(B * flop_frame / ms / 1e9, 1),
                  "ddim50_frames_per_s_per_gpu": round(B / (50 * ms / 1e3), 3),
                  "by_op_ms": {k: round(v["ms"], 2) for k, v in sorted(by.items(), key=lambda kv: -kv[1]["ms"])},
                  "attention_tflops": round(by["flash_attn"]["flops"] / by["flash_attn"]["ms"] / 1e9, 1),
                  "gemm_tflops": round(by["gemm"]["flops"] / by["gemm"]["ms"] / 1e9, 1),
                  "arena_GB": round(st.arena_bytes / 2 ** 30, 2)}))
