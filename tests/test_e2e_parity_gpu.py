"""End-to-end parity on non-degenerate predictions: the FULL oracle pipeline (fp32 UNet -> DDIM ->
fp32 seg-AE decoder -> argmax / threshold / merge -> oracle PQ / DVPQ evaluators) against the FULL CUDA pipeline
(TrainerDiffusion.sample -> panoptic_ids -> CityscapesPanopticEvaluator / dvpq_clip_sharded), on the same weights, image
latents and noise seed, at the 48x156 latent size of a 384x1248 frame.

Weights: random init + the "trained-like" recipe (ldmseg/models/unet_init.py) so that segments survive the merge; the
ground truth is a coarse, partly mislabelled copy of the ORACLE's prediction (ldmseg/data/synthetic.py), so TP, FP and
FN all occur. Stated tolerances:
  |PQ_cuda - PQ_oracle|, |DVPQ_cuda - DVPQ_oracle| (k = 1, 2)  <= 0.1 point          (north_star)
  share of pixels whose merged id differs                       <= 1 %               (measured: 0.4 %)
  latents after every DDIM step: relative L2 <= 1.5e-2, max-abs <= 2e-2 * max|ref|  (trajectory; measured 5e-3, flat)
  UNet epsilon on the oracle's own x_t at every step: relative L2 <= 1.5e-2          (teacher-forced; measured 6e-3)
The per-step curves go to gpurun_out/ (copied to profiles/ by hand after a GPU run).
"""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCHED_KW = dict(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                clip_sample=False, set_alpha_to_one=False, steps_offset=1, prediction_type="epsilon", weight="none")
EVAL = dict(mask_th=0.5, count_th=512, overlap_th=0.5)   # base.yaml:124-127
VAE_KW = dict(in_channels=16, int_channels=256, out_channels=128, latent_channels=4, num_upscalers=2,
              upscale_channels=256, norm_num_groups=32, scaling_factor=0.2)


def _rel(a, b):
    return ((a.float() - b.float()).pow(2).sum().sqrt() / (b.float().pow(2).sum().sqrt() + 1e-12)).item()


@pytest.fixture(scope="module")
def world():
    """Oracle and CUDA models with identical trained-like weights; the head is fitted on the ORACLE's features of frame 0."""
    from oracle import ldmseg_oracle as LO
    from oracle import unet_oracle as UO
    from video_latent_diffusion_panoptic_segmentation_b200.ldmseg.data import trained_like_rgb_latents
    from video_latent_diffusion_panoptic_segmentation_b200.ldmseg.models import GeneralVAESeg, UNet, unet_init
    from video_latent_diffusion_panoptic_segmentation_b200.ldmseg.schedulers import DDIMNoiseScheduler
    from video_latent_diffusion_panoptic_segmentation_b200.ldmseg.trainers import TrainerDiffusion
    h, w = 48, 156
    o_unet = UO.build_unet(seed=0, model_kwargs=unet_init.TRAINED_LIKE_MODEL_KWARGS)
    o_unet.load_state_dict(unet_init.trained_like_unet_(dict(o_unet.state_dict())))
    vsd = unet_init.trained_like_seg_decoder_(unet_init.random_seg_decoder_state_dict(seed=1, **VAE_KW))
    o_vae = LO.SegDecoderOracle(**VAE_KW)
    o_vae.load_state_dict(vsd)
    o_unet, o_vae = o_unet.to(DEV), o_vae.to(DEV).eval()
    heads = {}

    def fitted(T):
        """Models whose head was fitted to the oracle's T-step teacher sample of frame 0."""
        if T not in heads:
            rgb0 = trained_like_rgb_latents(1, h, w).to(DEV)
            lat0 = LO.sample(o_unet, LO.DDIMOracle(), rgb0, num_inference_steps=T, seed=42)
            with torch.no_grad():
                feats = o_vae.decoder[:-1](lat0 / o_vae.scaling_factor)
            heads[T] = unet_init.fit_seg_head(feats.float())
        sd = unet_init.set_seg_head_(dict(vsd), *heads[T])
        o_vae.load_state_dict({k: v.to(DEV) for k, v in sd.items()})
        vae = GeneralVAESeg(**VAE_KW, device=DEV)
        vae.load_state_dict(sd)
        return o_vae, vae

    unet = UNet(device=DEV)
    unet.load_state_dict({k: v.cpu() for k, v in o_unet.state_dict().items()})
    unet.remove_cross_attention()

    def trainer(vae):
        return TrainerDiffusion(p={"eval_kwargs": dict(EVAL), "ignore_label": 127}, vae_semseg=vae, unet_model=unet,
                                noise_scheduler=DDIMNoiseScheduler(**SCHED_KW), args={"gpu": 0})
    return dict(o_unet=o_unet, unet=unet, fitted=fitted, trainer=trainer, hw=(h, w))


# 16 frames: with 8 (about 470 segments in 19 classes) ONE borderline segment that the bf16 path keeps and the fp32 path
# drops (count_th = 512 pixels) moves the class-averaged PQ by 0.05 - 0.1 point, i.e. the 0.1-point tolerance of
# north_star was being tested at the resolution of the statistic itself (seen in round 2: T = 50, 8 frames, FP 221 vs 219
# -> |dPQ| 0.113 with unchanged per-step errors and 0.48 % differing pixels).
@pytest.mark.parametrize("T,B", [(10, 16), (50, 16)])
def test_full_pipeline_pq_dvpq_vs_oracle(world, T, B):
    from oracle import eval_oracle as EO
    from oracle import ldmseg_oracle as LO
    from video_latent_diffusion_panoptic_segmentation_b200.eval import clip_dvpq as CD
    from video_latent_diffusion_panoptic_segmentation_b200.ldmseg.data import (split_cat_ins, teacher_ground_truth,
                                                                               trained_like_rgb_latents)
    from video_latent_diffusion_panoptic_segmentation_b200.ldmseg.evaluations import CityscapesPanopticEvaluator
    h, w = world["hw"]
    o_vae, vae = world["fitted"](T)
    tr = world["trainer"](vae)
    rgb = trained_like_rgb_latents(B, h, w).to(DEV)

    # ---- oracle pipeline (fp32, TF32 off) with the (x_t, eps) trace of every step
    trace = []
    ref = LO.pipeline(world["o_unet"], o_vae, rgb, T, seed=42, ignore_label=127, trace=trace, **EVAL)
    # ---- CUDA pipeline
    lat_steps = tr.sample([""] * B, num_inference_steps=T, seed=42, rgb_latents=rgb, return_all_latents=True)
    lat_steps = lat_steps.view(T, B, 4, h, w)
    lat = lat_steps[-1].contiguous()
    ids, cleaned, _ = tr.panoptic_ids(lat)
    assert torch.equal(lat, tr.sample([""] * B, num_inference_steps=T, seed=42, rgb_latents=rgb))

    # ---- per-step error curves
    curve = []
    o_sched = LO.DDIMOracle()
    o_sched.set_timesteps_inference(T)
    for i, t in enumerate(o_sched.timesteps):
        x_ref, eps_ref = trace[i]
        nxt = trace[i + 1][0] if i + 1 < T else ref["latents"]
        eps_cuda = world["unet"](torch.cat([x_ref, rgb], 1), torch.tensor(int(t), device=DEV), None).sample
        curve.append({"step": i, "t": int(t),
                      "latents_rel_l2": _rel(lat_steps[i], nxt),
                      "latents_max_abs_over_max_ref": ((lat_steps[i] - nxt).abs().max() / nxt.abs().max()).item(),
                      "eps_rel_l2_teacher_forced": _rel(eps_cuda, eps_ref),
                      "eps_max_abs_over_max_ref": ((eps_cuda - eps_ref).abs().max() / eps_ref.abs().max()).item()})
    checks = []   # (ok, message): evaluated after the report has been written, so a failing run still leaves its numbers
    for c in curve:
        checks.append((c["latents_rel_l2"] <= 1.5e-2 and c["latents_max_abs_over_max_ref"] <= 2e-2, f"latents {c}"))
        checks.append((c["eps_rel_l2_teacher_forced"] <= 1.5e-2, f"eps {c}"))

    # ---- ids
    cl_cuda = cleaned.cpu().numpy().astype(np.int64)
    cl_ref = ref["cleaned"]
    diff_share = float((cl_cuda != cl_ref).mean())
    kept_ref = [len(np.unique(c[c >= 0])) for c in cl_ref]
    kept_cuda = [len(np.unique(c[c >= 0])) for c in cl_cuda]
    checks.append((min(kept_ref) >= 10 and min(kept_cuda) >= 10, f"degenerate: {kept_ref} {kept_cuda}"))
    checks.append((diff_share <= 0.01, f"share of different merged ids {diff_share}"))

    # ---- PQ: oracle evaluator on the oracle ids vs CUDA evaluator on the CUDA ids, same ground truth
    gt = teacher_ground_truth(torch.from_numpy(cl_ref))
    ev_o = EO.CityscapesPQOracle()
    ev_c = CityscapesPanopticEvaluator(thing_ids={11, 12, 13, 14, 15, 16, 17, 18}, device=DEV)
    for b in range(B):
        ev_o.add_image(cl_ref[b].copy(), gt[b].numpy())
    ev_c.add_images(cleaned, gt.to(DEV))
    pq_o, pq_c = ev_o.evaluate(), ev_c.evaluate()
    checks.append((pq_o["tp"] > 0 and pq_o["fp"] > 0 and pq_o["fn"] > 0, f"PQ statistics degenerate: {pq_o}"))
    checks.append((abs(pq_o["pq"] - pq_c["pq"]) <= 0.1, f"PQ oracle {pq_o['pq']} vs cuda {pq_c['pq']}"))

    # ---- DVPQ over the batch as a clip, windows of 1 and 2 frames
    gc, gi = split_cat_ins(gt.to(torch.int32), ignore=0)
    pc_o, pi_o = split_cat_ins(torch.from_numpy(cl_ref).to(torch.int32))
    pc_c, pi_c = split_cat_ins(cleaned)
    dvpq = {}
    for k in (1, 2):
        if B < k:
            continue
        rows = [EO.dvpq_window([pc_o[i + j].numpy() for j in range(k)], [pi_o[i + j].numpy() for j in range(k)],
                               [gc[i + j].numpy() for j in range(k)], [gi[i + j].numpy() for j in range(k)])
                for i in range(B - k + 1)]
        want = EO.dvpq_aggregate(rows)
        got = CD.dvpq_clip_sharded(pc_c, pi_c, gc.to(DEV), gi.to(DEV), n_frames=B, eval_frames=k)
        checks.append((want["tp"].sum() > 0, f"DVPQ k={k} has no TP"))
        checks.append((abs(want["pq"] - got["pq"]) <= 0.1, f"DVPQ k={k}: oracle {want['pq']} vs cuda {got['pq']}"))
        dvpq[k] = {"oracle": float(want["pq"]), "cuda": float(got["pq"]),
                   "oracle_tp_fn_fp": [int(want[x].sum()) for x in ("tp", "fn", "fp")],
                   "cuda_tp_fn_fp": [int(got[x].sum()) for x in ("tp", "fn", "fp")]}

    report = {"latent": [h, w], "frames": B, "ddim_steps": T, "eval_kwargs": EVAL,
              "segments_kept_oracle": kept_ref, "segments_kept_cuda": kept_cuda,
              "share_of_pixels_with_different_merged_id": diff_share,
              "share_of_pixels_with_different_id_before_merge": float((ids.cpu().numpy() != ref["ids"]).mean()),
              "pq": {"oracle": {k_: pq_o[k_] for k_ in ("pq", "sq", "rq", "tp", "fp", "fn")},
                     "cuda": {k_: pq_c[k_] for k_ in ("pq", "sq", "rq", "tp", "fp", "fn")}},
              "dvpq": dvpq, "per_step": curve}
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, f"e2e_parity_48x156_T{T}_B{B}.json"), "w") as f:
            json.dump(report, f, indent=1)
    print(json.dumps({k_: v for k_, v in report.items() if k_ != "per_step"}))
    failed = [msg for ok, msg in checks if not ok]
    assert not failed, failed
