"""Device time (CUDA graph of 20 launches, no host gaps) of the contraction kernel at the shapes where the tile count,
not the tensor pipe, bounds it: the 6x20 / 12x39 levels at B = 8 and everything at B = 1.

    python tools/bench_gemm_shapes.py [--B 8] [--json out.json]      (LDM_GEMM_SPLITK=0: every tile on one work item)
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from video_latent_diffusion_panoptic_segmentation_b200 import ops  # noqa: E402
from tools.bench_norms import graph_time  # noqa: E402

DEV, bf16, f32 = "cuda", torch.bfloat16, torch.float32
# (H, W, c1, c2, N, taps, launches per forward)
SHAPES = [(6, 20, 1280, 0, 1280, 9, 11), (6, 20, 1280, 1280, 1280, 9, 3), (12, 39, 1280, 0, 1280, 9, 7),
          (12, 39, 1280, 1280, 1280, 9, 2), (12, 39, 1280, 640, 1280, 9, 1), (24, 78, 640, 0, 640, 9, 6),
          (24, 78, 1280, 640, 640, 9, 1), (24, 78, 640, 640, 640, 9, 1), (48, 156, 320, 0, 320, 9, 7),
          (6, 20, 1280, 0, 1280, 1, 3), (6, 20, 5120, 0, 1280, 1, 1), (12, 39, 5120, 0, 1280, 1, 5),
          (12, 39, 1280, 0, 1280, 1, 15), (24, 78, 2560, 0, 640, 1, 5), (24, 78, 640, 0, 640, 1, 15),
          (48, 156, 1280, 0, 320, 1, 5), (48, 156, 320, 0, 320, 1, 15)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=8)
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    B, rows, tot = args.B, [], 0.0
    for H, W, c1, c2, N, taps, n in SHAPES:
        x1 = torch.randn((B, H, W, c1), device=DEV).to(bf16)
        x2 = torch.randn((B, H, W, c2), device=DEV).to(bf16) if c2 else None
        w = (torch.randn((N, taps * (c1 + c2)), device=DEV) * 0.02).to(bf16)
        bias, res = torch.randn(N, device=DEV), torch.randn((B * H * W, N), device=DEV).to(bf16)
        out = torch.empty((B, H, W, N), device=DEV, dtype=bf16)
        us = graph_time(lambda: ops.gemm(x1, w, out, a2=x2, taps=taps, bias=bias, residual=res))
        bn, pair, split = ops.gemm_last_config()
        fl = 2.0 * B * H * W * N * taps * (c1 + c2)
        rows.append({"M": B * H * W, "N": N, "K": taps * (c1 + c2), "taps": taps, "us": round(us, 2),
                     "tflops": round(fl / us / 1e6, 1), "block_n": bn, "pair": pair, "split_k": split, "launches": n})
        tot += us * n
        print(rows[-1], flush=True)
    print(f"B={B}: {tot / 1e3:.3f} ms per forward over these shapes (graph time)")
    if args.json:
        json.dump({"B": B, "rows": rows, "ms": tot / 1e3}, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
