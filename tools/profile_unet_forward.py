"""One eager UNet forward at the configs[1] shape (B = 8 frames of 384x1248), for ncu per-launch metrics.

    python tools/profile_unet_forward.py [--forwards N]
"""
import argparse
import copy
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from video_latent_diffusion_panoptic_segmentation_b200 import _lib as L  # noqa: E402
from video_latent_diffusion_panoptic_segmentation_b200.tools import main_ldm  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--forwards", type=int, default=1)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    L.lib()
    p = copy.deepcopy(main_ldm.BASE)
    vae, unet, sched = main_ldm.build_models(p, dev, seed=0)
    unet.use_cuda_graph = False
    B, h, w = 8, 48, 156
    x = torch.randn((B, 8, h, w), device=dev)
    t = torch.tensor(999, device=dev)
    n0 = L.launch_count()
    for _ in range(args.forwards):  # the first call builds the plan and runs it once eagerly
        out = unet(x, t, encoder_hidden_states=None).sample
    torch.cuda.synchronize()
    print(f"launches={L.launch_count() - n0} out_mean={out.float().mean().item():.5f}")


if __name__ == "__main__":
    main()
