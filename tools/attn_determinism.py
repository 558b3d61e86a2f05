"""Repeat one flash-attention launch on identical inputs; report bitwise differences between runs (count, magnitude,
where). Usage: python tools/attn_determinism.py [B heads d seq reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_latent_diffusion_panoptic_segmentation_b200 import ops  # noqa: E402

B, heads, d, seq, reps = (int(x) for x in (sys.argv[1:6] + ["8", "8", "40", "7488", "12"][len(sys.argv) - 1:]))
bf16 = torch.bfloat16
g = torch.Generator().manual_seed(0)
qkv = ops.alloc_qkv(B, heads, seq, d, "cuda")
qkv["q"][:, :, :d] = torch.randn((B * heads, seq, d), generator=g).to(bf16).cuda()
qkv["k"][:, :, :d] = torch.randn((B * heads, seq, d), generator=g).to(bf16).cuda()
qkv["vt"][:, :d, :seq] = torch.randn((B * heads, d, seq), generator=g).to(bf16).cuda()
out = torch.empty((B * seq, heads * d), dtype=bf16, device="cuda")


def run():
    out.fill_(float("nan"))
    ops.flash_attn(qkv["q"], qkv["k"], qkv["vt"], out, B=B, heads=heads, seq=seq, head_dim=d, dpad=qkv["dpad"],
                   seq_pad=qkv["seq_pad"], scale=d ** -0.5)
    torch.cuda.synchronize()
    return out.clone()


ref = run()
print("nan in ref:", int(torch.isnan(ref.float()).sum()))
for r in range(reps):
    cur = run()
    diff = (cur.float() - ref.float()).abs()
    ne = cur.view(torch.int16) != ref.view(torch.int16)
    n = int(ne.sum())
    if n:
        idx = ne.nonzero()
        rows = idx[:, 0].unique()
        cols = idx[:, 1].unique()
        img = (rows // seq).unique().tolist()
        tok = (rows % seq)
        print(f"rep {r}: {n} elements differ, max abs {float(diff.max()):.4g} (ref max {float(ref.float().abs().max()):.3g}), "
              f"{rows.numel()} rows, images {img}, tokens {int(tok.min())}..{int(tok.max())}, "
              f"heads {sorted(set((cols // d).tolist()))}, first rows {rows[:8].tolist()}")
    else:
        print(f"rep {r}: identical")
