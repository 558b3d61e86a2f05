"""Run one UNet plan eagerly many times on identical inputs and report the first launch whose output differs between
runs (bitwise). Usage: python tools/determinism_diag.py [B h w repeats]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_latent_diffusion_panoptic_segmentation_b200 import ops  # noqa: E402
from video_latent_diffusion_panoptic_segmentation_b200.ldmseg.models import UNet, unet_init  # noqa: E402

B, h, w, reps = (int(x) for x in (sys.argv[1:5] + ["2", "16", "24", "30"][len(sys.argv) - 1:]))
unet = UNet(device="cuda")
unet.load_state_dict(unet_init.random_unet_state_dict(seed=0, in_channels=8))
unet.remove_cross_attention()
unet.use_cuda_graph = False
st = unet._get_plan(B, h, w, 8)
st.sample.copy_(torch.randn((B, 8, h, w), generator=torch.Generator().manual_seed(1)).cuda())
st.timestep.fill_(499)


def outputs(fn, a, k):
    if fn is ops.gemm:
        if a[2] is not None:
            return [a[2]]
        q = k["qkv"]
        return [q[n] for n in ("q", "k", "vt") if q.get(n) is not None]
    if fn in (ops.groupnorm, ops.layernorm, ops.flash_attn):
        return [a[3]]
    if fn is ops.conv3x3_small_cin:
        return [a[3]]
    if fn in (ops.im2col3x3_s2, ops.upsample_nearest):
        return [a[1]]
    if fn is ops.gemv:
        return [a[2]]
    if fn is ops.timestep_sinusoid:
        return [a[3]]
    return []


def run():
    sums = []
    for fn, a, k in st.plan:
        fn(*a, **k)
        sums.append([t.clone() for t in outputs(fn, a, k)])
    torch.cuda.synchronize()
    return sums


ref = run()
bad = {}
for r in range(reps):
    cur = run()
    for i, (x, y) in enumerate(zip(ref, cur)):
        if any(not torch.equal(p.view(torch.uint8) if p.dtype != torch.bfloat16 else p.view(torch.int16),
                               q.view(torch.uint8) if q.dtype != torch.bfloat16 else q.view(torch.int16))
               for p, q in zip(x, y)):
            fn, a, k = st.plan[i]
            desc = fn.__name__
            if fn is ops.gemm:
                desc += f" M={a[0].numel() // a[0].shape[-1]} N={a[1].shape[0]} K={a[1].shape[1]} taps={k.get('taps', 1)} flags={k.get('flags', 0)} res={'residual' in k}"
            elif fn is ops.flash_attn:
                desc += f" seq={k['seq']} d={k['head_dim']}"
            elif fn is ops.groupnorm:
                desc += f" shape={tuple(a[3].shape)}"
            nd = sum(int((p.float() != q.float()).sum()) for p, q in zip(x, y))
            bad.setdefault((i, desc), []).append((r, nd))
            break  # first differing launch of this repeat
print("launches", len(st.plan), "repeats", reps, "first-differing launches:", len(bad))
for (i, desc), v in sorted(bad.items()):
    print(i, desc, v[:6])
