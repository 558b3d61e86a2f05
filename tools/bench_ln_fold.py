"""Device time (CUDA graph, no host gaps) of the GEMMs around a BasicTransformerBlock LayerNorm, with the LayerNorm as
its own pass and folded into them (ldm_gemm_desc.ln_stats), per UNet level at batch B.

    python tools/bench_ln_fold.py [--B 8] [--json out.json] [--ncu geglu|qkv]   (--ncu: launch the folded kernel once)
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from video_latent_diffusion_panoptic_segmentation_b200 import _lib as L  # noqa: E402
from video_latent_diffusion_panoptic_segmentation_b200 import ops  # noqa: E402
from tools.bench_norms import graph_time  # noqa: E402

DEV, bf16, f32 = "cuda", torch.bfloat16, torch.float32
LEVELS = [(48 * 156, 320, 8, 40), (24 * 78, 640, 8, 80), (12 * 39, 1280, 8, 160)]  # (tokens, C, heads, d)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=8)
    ap.add_argument("--json", default=None)
    ap.add_argument("--ncu", default=None)
    args = ap.parse_args()
    B, rows = args.B, []
    for seq, C, heads, d in LEVELS:
        M = B * seq
        a = torch.randn((M, C), device=DEV).to(bf16)
        res = torch.randn((M, C), device=DEV).to(bf16)
        wp, bp = (torch.randn((C, C), device=DEV) * 0.05).to(bf16), torch.randn(C, device=DEV)
        x, xn = torch.empty((M, C), device=DEV, dtype=bf16), torch.empty((M, C), device=DEV, dtype=bf16)
        stats = torch.empty(((C + 31) // 32, M, 2), device=DEV, dtype=f32)
        gamma, beta = torch.rand(C, device=DEV) + 0.5, torch.randn(C, device=DEV) * 0.1
        inner = 4 * C
        w1, b1 = torch.randn((8 * C, C), device=DEV) * 0.05, torch.randn(8 * C, device=DEV)
        wq = torch.randn((3 * C, C), device=DEV) * 0.05
        w1f, b1f, cs1 = ops.fold_layernorm(w1, b1, gamma, beta)
        wqf, bqf, csq = ops.fold_layernorm(wq, None, gamma, beta)
        w1b, wqb = w1.to(bf16), wq.to(bf16)
        g = torch.empty((M, inner), device=DEV, dtype=bf16)
        qkv = ops.alloc_qkv(B, heads, seq, d, DEV)
        qd = dict(q=qkv["q"], k=qkv["k"], vt=qkv["vt"], heads=heads, head_dim=d, dpad=qkv["dpad"], seq=seq,
                  seq_pad=qkv["seq_pad"])
        ops.gemm(a, wp, x, bias=bp, residual=res, row_stats=stats)
        if args.ncu:
            if seq != LEVELS[0][0]:
                continue
            if args.ncu == "geglu":
                ops.gemm(x, w1f, g, bias=b1f, flags=L.LDM_GEMM_GEGLU, ln_fold=(stats, cs1, 1e-5))
                ops.gemm(x, w1b, g, bias=b1, flags=L.LDM_GEMM_GEGLU)
            else:
                ops.gemm(x, wqf, None, bias=bqf, flags=L.LDM_GEMM_QKV_SPLIT, qkv=qd, ln_fold=(stats, csq, 1e-5))
                ops.gemm(x, wqb, None, flags=L.LDM_GEMM_QKV_SPLIT, qkv=qd)
            torch.cuda.synchronize()
            continue
        t = {
            "producer": graph_time(lambda: ops.gemm(a, wp, x, bias=bp, residual=res)),
            "producer_stats": graph_time(lambda: ops.gemm(a, wp, x, bias=bp, residual=res, row_stats=stats)),
            "layernorm": graph_time(lambda: ops.layernorm(x, gamma, beta, xn, 1e-5)),
            "geglu": graph_time(lambda: ops.gemm(xn, w1b, g, bias=b1, flags=L.LDM_GEMM_GEGLU)),
            "geglu_fold": graph_time(lambda: ops.gemm(x, w1f, g, bias=b1f, flags=L.LDM_GEMM_GEGLU,
                                                      ln_fold=(stats, cs1, 1e-5))),
            "qkv": graph_time(lambda: ops.gemm(xn, wqb, None, flags=L.LDM_GEMM_QKV_SPLIT, qkv=qd)),
            "qkv_fold": graph_time(lambda: ops.gemm(x, wqf, None, bias=bqf, flags=L.LDM_GEMM_QKV_SPLIT, qkv=qd,
                                                    ln_fold=(stats, csq, 1e-5))),
        }
        row = {"M": M, "C": C, **{k: round(v, 2) for k, v in t.items()}}
        row["per_block_unfolded"] = round(2 * t["producer"] + 2 * t["layernorm"] + t["geglu"] + t["qkv"], 2)
        row["per_block_folded"] = round(2 * t["producer_stats"] + t["geglu_fold"] + t["qkv_fold"], 2)
        rows.append(row)
        print(row, flush=True)
    if args.json:
        json.dump({"B": B, "rows": rows}, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
