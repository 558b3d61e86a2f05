"""Generate the golden fixtures of tests/golden/ from the REAL reference code (run in the authoring container only;
/root/reference does not exist on the GPU box, so the fixtures are committed).

    python tests/golden/make_golden.py

Everything written here comes from importing and running reference code (oracle/ref_stubs.py only satisfies its
imports) or from the reference's own sample_outputs/ PNGs; the oracle restatements are then checked against these
files by tests/test_oracle_golden.py.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_stubs  # noqa: E402
sys.path.insert(0, os.path.join(ROOT, "tests"))
from synth import synth_panoptic, vpq_case, city_case  # noqa: E402

ref_stubs.install()
sys.argv = [sys.argv[0]]  # eval/eval_dvpq.py parses argv at import

from ldmseg.schedulers import DDIMNoiseScheduler  # noqa: E402
from ldmseg.models.vae import GeneralVAESeg  # noqa: E402

SCHED_KW = dict(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                clip_sample=False, set_alpha_to_one=False, steps_offset=1, prediction_type="epsilon",
                thresholding=False, weight="none", verbose=False)
AE_SMALL = dict(in_channels=16, int_channels=64, out_channels=16, block_out_channels=(32, 64, 128, 256),
                latent_channels=4, num_latents=2, num_upscalers=2, upscale_channels=64, norm_num_groups=32,
                scaling_factor=0.2, parametrization="gaussian", act_fn="none", num_mid_blocks=0)


def main():
    out = {}
    # ---------------------------------------------------------------- scheduler (ddim_scheduler.py)
    s = DDIMNoiseScheduler(**SCHED_KW)
    sched = {"alphas_cumprod_0_19_999": [float(s.alphas_cumprod[i]) for i in (0, 19, 999)]}
    for T in (10, 50):
        s.set_timesteps_inference(T)
        sched[f"timesteps_{T}"] = s.timesteps.tolist()
    g = torch.Generator().manual_seed(123)
    eps, x = torch.randn((2, 4, 6, 10), generator=g), torch.randn((2, 4, 6, 10), generator=g)
    s.set_timesteps_inference(50)
    steps = {}
    for t in (999, 499, 19):
        o = s.step(eps, torch.tensor(t), x)
        steps[f"prev_{t}"] = o.prev_sample.numpy()
        steps[f"x0_{t}"] = o.pred_original_sample.numpy()
    np.savez_compressed(os.path.join(HERE, "ddim_steps.npz"), eps=eps.numpy(), x=x.numpy(),
                        alphas_cumprod=s.alphas_cumprod.numpy(), **steps)
    out["scheduler"] = sched

    # ---------------------------------------------------------------- seg-AE decoder (vae.py), reduced widths
    torch.manual_seed(7)
    vae = GeneralVAESeg(**AE_SMALL).eval()
    with torch.no_grad():
        for p in vae.parameters():  # non-trivial norm affine parameters
            if p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
    z = torch.randn((2, 4, 6, 10), generator=torch.Generator().manual_seed(8))
    with torch.no_grad():
        logits_lo = vae.decode(z, interpolate=False)
        logits_hi = vae.decode(z, interpolate=True)
    sd = {k: v.numpy() for k, v in vae.state_dict().items() if k.startswith("decoder.")}
    np.savez_compressed(os.path.join(HERE, "seg_decoder_small.npz"), z=z.numpy(), logits_lo=logits_lo.numpy(),
                        logits_hi=logits_hi.numpy(), **{"sd." + k: v for k, v in sd.items()})
    out["seg_decoder_small"] = {"cfg": {k: (list(v) if isinstance(v, tuple) else v) for k, v in AE_SMALL.items()},
                                "keys": sorted(sd), "interpolation_factor": vae.interpolation_factor,
                                "downsample_factor": vae.downsample_factor}

    # ---------------------------------------------------------------- bit codec golden vector (sample_outputs/)
    from PIL import Image
    so = os.path.join(ref_stubs.REFERENCE, "sample_outputs")
    sem = np.array(Image.open(os.path.join(so, "semseg.png")))
    bits = np.stack([np.array(Image.open(os.path.join(so, f"bit_channel_{i}.png"))) for i in range(16)])
    from ldmseg.data.cityscapes import Cityscapes

    class _Self:
        ignore_label = 0
    enc, ign = Cityscapes.encode_bitmap(_Self(), torch.from_numpy(sem.astype(np.int64)), n=16, fill_value=0.5)
    dec = Cityscapes.decode_bitmap(_Self(), enc)
    assert np.array_equal((enc.numpy() * 255).astype(np.uint8), bits), "reference encode_bitmap != sample_outputs PNGs"
    np.savez_compressed(os.path.join(HERE, "bitmap_sample_outputs.npz"), semseg=sem, bits_u8=bits,
                        decoded=dec.numpy().astype(np.int32))
    out["bitmap"] = {"n_ids": int(len(np.unique(sem))), "n_id31": int((sem == 31).sum()),
                     "n_ignore": int(ign.sum())}

    # ---------------------------------------------------------------- vpq_eval (eval_dvpq.py, new_eval.py)
    dv = ref_stubs.load_by_path("ref_eval_dvpq", "eval/eval_dvpq.py")
    ne = ref_stubs.load_by_path("ref_new_eval", "eval/new_eval.py")
    vpq = {}
    for case, (seed, H, W) in {"a": (11, 48, 156), "b": (12, 96, 312), "c": (13, 64, 64)}.items():
        pred, gt = vpq_case(seed, H, W)
        # int64 inputs: with int32 maps (what eval_dvpq.py:112-121 builds) numpy >= 2 overflows int32 in
        # `_ign_id * offset + pred_id` (eval_dvpq.py:60) and silently disables the ignored-overlap FP filter;
        # numpy 1.x promoted to int64. The fixtures pin the non-overflowing (intended, numpy-1.x) semantics and
        # record whether the int32 call differs in this container.
        r = dv.vpq_eval([pred.astype(np.int64), gt.astype(np.int64)])
        with np.errstate(over="ignore"):
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                r32 = dv.vpq_eval([pred.astype(np.int32), gt.astype(np.int32)])
        vpq[case] = {"seed": seed, "H": H, "W": W, "out": [x.tolist() for x in r],
                     "int32_call_differs_here": bool(any((a != b).any() for a, b in zip(r, r32)))}
        # max_ins = 64 variant on re-encoded ids
        gt64 = (gt // 2 ** 20) * 64 + (gt % 2 ** 20) % 64
        pr64 = (pred // 2 ** 20) * 64 + (pred % 2 ** 20) % 64
        r2 = ne.vpq_eval([pr64.astype(np.int64), gt64.astype(np.int64)])
        vpq[case]["out64"] = [x.tolist() for x in r2]
    out["vpq"] = vpq

    # ---------------------------------------------------------------- CityscapesPanopticEvaluator
    ce = ref_stubs.load_by_path("ref_city_eval", "ldmseg/evaluations/cityscapes_pap_eval.py")
    ev = ce.CityscapesPanopticEvaluator(thing_ids={11, 12, 13, 14, 15, 16, 17, 18})
    city = {"images": []}
    for seed in (21, 22, 23):
        pred, gt_sem = city_case(seed)
        ev.add_image(pred, gt_sem)
        city["images"].append({"seed": seed, "tp": ev.TP, "fp": ev.FP, "fn": ev.FN, "iou_sum": ev.iou_sum})
    res = ev.evaluate()
    city["result"] = {k: (v if not isinstance(v, dict) else {str(c): m for c, m in v.items()}) for k, v in res.items()}
    out["cityscapes_pq"] = city

    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()
