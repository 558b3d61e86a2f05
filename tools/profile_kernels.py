"""Launch the dominant kernels of the UNet at their configs[1] shapes (B = 8 frames of 384x1248, latent 48x156), one
after another, for ncu captures and CUDA-event timings.

    python tools/profile_kernels.py [--iters N] [--only substr,substr] [--json out.json]

Each case is launched `iters` times back to back (inputs of the large cases exceed L2 anyway); the event time is
the mean per launch. Under ncu use --iters 1.
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

DEV, bf16, f32 = "cuda", torch.bfloat16, torch.float32
B = 8
LEVELS = {"L0": (48, 156, 320), "L1": (24, 78, 640), "L2": (12, 39, 1280), "L3": (6, 20, 1280)}


def rn(shape, dtype=bf16, scale=1.0):
    return (torch.randn(shape, device=DEV) * scale).to(dtype)


def cases():
    out = {}
    for lv, (h, w, C) in LEVELS.items():
        M = B * h * w
        if lv != "L3":
            a, wt, bias, res, o = rn((M, C)), rn((C, C), scale=0.05), rn((C,), f32), rn((M, C)), rn((M, C))
            out[f"gemm1x1_res_{lv}"] = (lambda a=a, wt=wt, bias=bias, res=res, o=o: ops.gemm(a, wt, o, bias=bias, residual=res),
                                       2 * M * C * C, 2 * (3 * M * C + C * C))
            w1, b1, g = rn((8 * C, C), scale=0.05), rn((8 * C,), f32), rn((M, 4 * C))
            out[f"gemm_geglu_{lv}"] = (lambda a=a, w1=w1, b1=b1, g=g: ops.gemm(a, w1, g, bias=b1, flags=L.LDM_GEMM_GEGLU),
                                      2 * M * 8 * C * C, 2 * (M * C + 8 * C * C + M * 4 * C))
            w2 = rn((C, 4 * C), scale=0.05)
            out[f"gemm_ff2_{lv}"] = (lambda g=g, w2=w2, bias=bias, res=res, o=o: ops.gemm(g, w2, o, bias=bias, residual=res),
                                    2 * M * 4 * C * C, 2 * (M * 4 * C + 4 * C * C + 2 * M * C))
            d = C // 8
            qkv = ops.alloc_qkv(B, 8, h * w, d, DEV)
            wq = rn((3 * C, C), scale=0.05)
            out[f"gemm_qkv_{lv}"] = (lambda a=a, wq=wq, qkv=qkv: ops.gemm(a, wq, None, flags=L.LDM_GEMM_QKV_SPLIT, qkv=qkv),
                                    2 * M * 3 * C * C, 2 * (M * C + 3 * C * C + 3 * M * C))
            ao = rn((M, C))
            out[f"attn_{lv}"] = (lambda qkv=qkv, ao=ao, h=h, w=w, d=d: ops.flash_attn(
                qkv["q"], qkv["k"], qkv["vt"], ao, B=B, heads=8, seq=h * w, head_dim=d, dpad=qkv["dpad"],
                seq_pad=qkv["seq_pad"], scale=d ** -0.5), 4 * B * 8 * (h * w) ** 2 * d, 2 * 4 * M * C)
            import math
            q2 = (qkv["q"].float() * (d ** -0.5 * math.log2(math.e))).to(bf16)  # scores in log2 units, same spread
            out[f"attnfold_{lv}"] = (lambda qkv=qkv, q2=q2, ao=ao, h=h, w=w, d=d: ops.flash_attn(
                q2, qkv["k"], qkv["vt"], ao, B=B, heads=8, seq=h * w, head_dim=d, dpad=qkv["dpad"],
                seq_pad=qkv["seq_pad"], scale=math.log(2.0)), 4 * B * 8 * (h * w) ** 2 * d, 2 * 4 * M * C)
            ln_g, ln_b = rn((C,), f32), rn((C,), f32)
            out[f"layernorm_{lv}"] = (lambda a=a, o=o, ln_g=ln_g, ln_b=ln_b: ops.layernorm(a, ln_g, ln_b, o, 1e-5), 0,
                                     2 * 2 * M * C)
        x = rn((B, h, w, C))
        w3, bias3, o3, res3 = rn((C, 9 * C), scale=0.02), rn((C,), f32), rn((B, h, w, C)), rn((M, C))
        out[f"conv3x3_{lv}"] = (lambda x=x, w3=w3, bias3=bias3, o3=o3, res3=res3: ops.gemm(x, w3, o3, taps=9, bias=bias3, residual=res3),
                               2 * M * C * 9 * C, 2 * (3 * M * C + 9 * C * C))
        gam, bet, stats = rn((C,), f32), rn((C,), f32), ops.gn_scratch(B, 32, DEV)
        out[f"groupnorm_{lv}"] = (lambda x=x, gam=gam, bet=bet, o3=o3, stats=stats: ops.groupnorm(x, gam, bet, o3, stats, groups=32, eps=1e-5, silu=True),
                                 0, 2 * 2 * M * C)
    # concat-input resnet conv1 of the last up block (960 -> 320 at L0) and its GroupNorm
    h, w, _ = LEVELS["L0"]
    xa, xb = rn((B, h, w, 640)), rn((B, h, w, 320))
    t0, wc, bc, oc = rn((B, h, w, 960)), rn((320, 9 * 960), scale=0.02), rn((320,), f32), rn((B, h, w, 320))
    gam, bet, stats = rn((960,), f32), rn((960,), f32), ops.gn_scratch(B, 32, DEV)
    out["groupnorm_cat960_L0"] = (lambda: ops.groupnorm(xa, gam, bet, t0, stats, x2=xb, groups=32, eps=1e-5, silu=True), 0,
                                  2 * 2 * B * h * w * 960)
    out["conv3x3_960to320_L0"] = (lambda: ops.gemm(t0, wc, oc, taps=9, bias=bc), 2 * B * h * w * 320 * 9 * 960,
                                  2 * (B * h * w * 1280 + 9 * 960 * 320))
    # Upsample2D (nearest x2 + conv3x3) as four sub-pixel 2x2 convolutions, against the gather pass + 3x3 convolution
    # it replaces; Downsample2D (conv3x3 stride 2) through the element-strided A map (ldm_gemm_desc.up2 / a_stride)
    for name, (hl, wl, Cu) in {"L2toL1": (12, 39, 1280), "L1toL0": (24, 78, 640)}.items():
        xl = rn((B, hl, wl, Cu))
        w9 = torch.randn((Cu, 3, 3, Cu), device=DEV) * 0.02
        bu = rn((Cu,), f32)
        w4, b4 = ops.fold_upsample_conv3x3(w9, bu)
        oh_ = rn((B, 2 * hl, 2 * wl, Cu))
        upb = rn((B, 2 * hl, 2 * wl, Cu))
        w9p = w9.reshape(Cu, 9 * Cu).to(bf16).contiguous()
        Ml = B * hl * wl
        out[f"upconv_subpixel_{name}"] = (lambda xl=xl, w4=w4, b4=b4, oh_=oh_: ops.gemm(xl, w4, oh_, taps=4, bias=b4, up2=True),
                                          2 * Ml * 4 * Cu * 4 * Cu, 2 * (5 * Ml * Cu + 16 * Cu * Cu))
        out[f"upconv_gather_{name}"] = (lambda xl=xl, upb=upb, w9p=w9p, bu=bu, oh_=oh_: (ops.upsample_nearest(xl, upb),
                                                                                     ops.gemm(upb, w9p, oh_, taps=9, bias=bu)),
                                        2 * 4 * Ml * Cu * 9 * Cu, 2 * (13 * Ml * Cu + 9 * Cu * Cu))
    hd, wd2, Cd = LEVELS["L0"]
    xd = rn((B, hd, wd2, Cd))
    wdn, bdn, od = rn((Cd, 9 * Cd), scale=0.02), rn((Cd,), f32), rn((B, hd // 2, wd2 // 2, Cd))
    cold = rn((B * (hd // 2) * (wd2 // 2), 9 * Cd))
    out["downconv_strided_L0toL1"] = (lambda: ops.gemm(xd, wdn, od, taps=9, bias=bdn, a_stride=2, a_pad=1),
                                      2 * od.numel() * 9 * Cd, 2 * (xd.numel() + od.numel() + wdn.numel()))
    out["downconv_im2col_L0toL1"] = (lambda: (ops.im2col3x3_s2(xd, cold), ops.gemm(cold, wdn, od.view(-1, Cd), bias=bdn)),
                                     2 * od.numel() * 9 * Cd, 2 * (xd.numel() + 2 * cold.numel() + od.numel() + wdn.numel()))
    # HBM-bound tail / scheduler kernels at full size (8 frames of 384x1248)
    H, W, Cc = 192, 624, 128
    lo = torch.randn((B, H, W, Cc), device=DEV)
    ids = torch.empty((B, 2 * H, 2 * W), dtype=torch.int32, device=DEV)
    counts = torch.empty((B, 2, Cc), dtype=torch.int32, device=DEV)
    cleaned = torch.empty_like(ids)
    out["logits_to_ids_up2"] = (lambda: ops.logits_to_ids(lo, ids, counts, up=2, mask_th=0.5, ignore_label=127), 0,
                                lo.numel() * 4 + ids.numel() * 4)
    out["segment_filter"] = (lambda: ops.segment_filter(ids, counts, cleaned, count_th=512, overlap_th=0.5, ignore_label=127),
                             0, 2 * ids.numel() * 4)
    bits = torch.randn((B, 16, 2 * H, 2 * W), device=DEV)
    out["decode_bitmap16"] = (lambda: ops.decode_bitmap(bits, ids), 0, bits.numel() * 4 + ids.numel() * 4)
    gt = torch.randint(0, 19, (B * 2 * H * 2 * W,), dtype=torch.int32, device=DEV)
    pr = torch.randint(0, 40, (B * 2 * H * 2 * W,), dtype=torch.int32, device=DEV)
    hk = torch.empty(1 << 14, dtype=torch.int64, device=DEV)
    hc = torch.empty(1 << 14, dtype=torch.int32, device=DEV)
    ho = torch.empty(1, dtype=torch.int32, device=DEV)
    import ctypes as Cx
    out["joint_hist"] = (lambda: L.check(L.lib().ldm_joint_hist(Cx.c_void_p(gt.data_ptr()), Cx.c_void_p(pr.data_ptr()), gt.numel(),
                                                               Cx.c_void_p(hk.data_ptr()), Cx.c_void_p(hc.data_ptr()), 1 << 14,
                                                               Cx.c_void_p(ho.data_ptr()),
                                                               Cx.c_void_p(torch.cuda.current_stream().cuda_stream)), "joint_hist"),
                         0, 2 * gt.numel() * 4)
    eps, xs = torch.randn((B, 4, 48, 156), device=DEV), torch.randn((B, 4, 48, 156), device=DEV)
    coef = torch.rand((50, 4), device=DEV) + 0.1
    ti = torch.zeros(1, dtype=torch.int32, device=DEV)
    out["ddim_step"] = (lambda: ops.ddim_step(eps, xs, coef, ti, prev_sample=xs), 0, 3 * eps.numel() * 4)
    return out


def sweep_block_n(iters):
    """Time the contraction kernel for every block_n on the UNet's (M, N, K, taps) shapes (plain bf16 epilogue)."""
    shapes = [(48, 156, 320, 320, 1), (48, 156, 320, 320, 9), (48, 156, 640, 320, 9), (48, 156, 320, 1280, 1),
              (24, 78, 640, 640, 1), (24, 78, 640, 640, 9), (24, 78, 1280, 640, 9), (24, 78, 640, 2560, 1),
              (12, 39, 1280, 1280, 1), (12, 39, 1280, 1280, 9), (12, 39, 2560, 1280, 9), (12, 39, 1280, 5120, 1),
              (6, 20, 1280, 1280, 1), (6, 20, 1280, 1280, 9), (6, 20, 2560, 1280, 9)]
    rows = []
    for (h, w, cin, N, taps) in shapes:
        x = rn((B, h, w, cin))
        wt, bias, o = rn((N, taps * cin), scale=0.02), rn((N,), f32), rn((B, h, w, N))
        M = B * h * w
        best = None
        for bn in (0, 64, 96, 128, 160, 192, 224, 256):
            fn = lambda: ops.gemm(x, wt, o, taps=taps, bias=bias, block_n=bn)
            fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                fn()
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / iters
            rows.append({"M": M, "N": N, "K": taps * cin, "taps": taps, "block_n": bn, "us": round(us, 2),
                         "tflops": round(2 * M * N * taps * cin / us / 1e6, 1)})
            print(json.dumps(rows[-1]), flush=True)
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sweep", action="store_true", help="block_n sweep of the contraction kernel")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--only", default="")
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    torch.manual_seed(0)
    L.lib()
    sel = [s for s in args.only.split(",") if s]
    if args.sweep:
        rows = sweep_block_n(args.iters)
        if args.json:
            with open(args.json, "w") as f:
                json.dump(rows, f, indent=1)
        return
    rows = []
    for name, (fn, flops, nbytes) in cases().items():
        if sel and not any(s in name for s in sel):
            continue
        fn()  # warm-up (sets the kernel attributes)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / args.iters
        rows.append({"case": name, "us": round(us, 2), "tflops": round(flops / us / 1e6, 1),
                     "gbs": round(nbytes / us / 1e3, 1)})
        print(json.dumps(rows[-1]), flush=True)
    if args.json:
        with open(args.json, "w") as f:
            json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
