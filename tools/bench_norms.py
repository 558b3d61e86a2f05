"""Device time of every GroupNorm / LayerNorm shape of one UNet forward (B frames of 384x1248), each launched 20x inside
a CUDA graph (no host gaps), against torch for correctness.

    python tools/bench_norms.py [--B 8] [--json out.json]
"""
import argparse
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from video_latent_diffusion_panoptic_segmentation_b200 import ops  # noqa: E402

DEV, bf16, f32 = "cuda", torch.bfloat16, torch.float32
# (HW, c1, c2, launches per forward)
GN_SHAPES = [(7488, 320, 0, 13), (7488, 320, 320, 2), (7488, 640, 320, 1), (1872, 320, 0, 1), (1872, 640, 0, 11),
             (1872, 640, 320, 1), (1872, 640, 640, 1), (1872, 1280, 640, 1), (468, 640, 0, 1), (468, 1280, 0, 11),
             (468, 1280, 640, 1), (468, 1280, 1280, 2), (120, 1280, 0, 12), (120, 1280, 1280, 3)]
LN_SHAPES = [(7488, 320, 10), (1872, 640, 10), (468, 1280, 10), (120, 1280, 2)]


def graph_time(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters):
            fn()
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / (5 * iters)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=8)
    ap.add_argument("--json", default=None)
    ap.add_argument("--only", default=None, help="HW,C: only the GroupNorm shape(s) with this HW and c1+c2")
    args = ap.parse_args()
    B = args.B
    rows, tot = [], 0.0
    only = tuple(int(v) for v in args.only.split(",")) if args.only else None
    for HW, c1, c2, n in GN_SHAPES:
        C = c1 + c2
        if only and (HW, C) != only:
            continue
        x1 = (torch.randn((B, HW, c1), device=DEV) * 1.5 + 0.3).to(bf16)
        x2 = (torch.randn((B, HW, c2), device=DEV) * 0.7 - 0.2).to(bf16) if c2 else None
        g, b = torch.randn(C, device=DEV), torch.randn(C, device=DEV)
        out = torch.empty((B, HW, C), device=DEV, dtype=bf16)
        stats = ops.gn_scratch(B, 32, DEV)
        fn = lambda: ops.groupnorm(x1, g, b, out, stats, x2=x2, groups=32, eps=1e-5, silu=True)
        us = graph_time(fn)
        x = torch.cat([x1, x2], -1) if c2 else x1
        ref = F.silu(F.group_norm(x.float().permute(0, 2, 1), 32, g, b, 1e-5)).permute(0, 2, 1)
        err = (out.float() - ref).abs().max().item()
        nbytes = 2 * B * HW * C * 2
        rows.append({"op": "groupnorm", "HW": HW, "c1": c1, "c2": c2, "us": round(us, 2), "gbs": round(nbytes / us / 1e3),
                     "launches": n, "max_abs_err": err})
        tot += us * n
        print(rows[-1], flush=True)
    gn_tot = tot
    for HW, C, n in ([] if only else LN_SHAPES):
        x = torch.randn((B * HW, C), device=DEV).to(bf16)
        g, b = torch.randn(C, device=DEV), torch.randn(C, device=DEV)
        out = torch.empty_like(x)
        us = graph_time(lambda: ops.layernorm(x, g, b, out, 1e-5))
        err = (out.float() - F.layer_norm(x.float(), (C,), g, b, 1e-5)).abs().max().item()
        rows.append({"op": "layernorm", "HW": HW, "C": C, "us": round(us, 2), "gbs": round(4 * B * HW * C / us / 1e3),
                     "launches": n, "max_abs_err": err})
        tot += us * n
        print(rows[-1], flush=True)
    print(f"B={B}: GroupNorm {gn_tot / 1e3:.3f} ms, LayerNorm {(tot - gn_tot) / 1e3:.3f} ms per forward (graph time)")
    if args.json:
        json.dump({"B": B, "rows": rows, "gn_ms": gn_tot / 1e3, "ln_ms": (tot - gn_tot) / 1e3}, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
