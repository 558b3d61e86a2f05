"""CPU tests of the host-side logic (no kernels are launched): parameter shapes vs the oracle modules, the scheduler
mirror's host schedule vs the reference golden fixture, modify_encoder, the cross-rank statistics reduction
(world_size 2, gloo), and the loud failure when compute is attempted without a GPU."""
import json
import os

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from oracle import ldmseg_oracle as LO
from oracle import unet_oracle as UO
from video_latent_diffusion_panoptic_segmentation_b200 import _lib as L
from video_latent_diffusion_panoptic_segmentation_b200.ldmseg.models import unet_init
from video_latent_diffusion_panoptic_segmentation_b200.ldmseg.models.unet import UNet
from video_latent_diffusion_panoptic_segmentation_b200.ldmseg.schedulers import DDIMNoiseScheduler

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLD = json.load(open(os.path.join(G, "golden.json")))
SCHED_KW = dict(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                clip_sample=False, set_alpha_to_one=False, steps_offset=1, prediction_type="epsilon", weight="none")


def test_unet_param_shapes_match_oracle_module():
    cfg = dict(block_out_channels=(64, 128, 256, 256))
    net = UO.UNetOracle(**cfg)
    want = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    got = unet_init.unet_param_shapes(**cfg)
    assert got == want
    assert list(got) == list(want)  # same order as well


def test_full_size_unet_parameter_count():
    shapes = unet_init.unet_param_shapes(in_channels=8)
    n = sum(int(np.prod(s)) for s in shapes.values())
    assert abs(n - 815.4e6) < 0.5e6  # ~815.4 M with cross-attention removed and 8-ch conv_in


def test_seg_decoder_param_shapes_match_oracle_module():
    kw = dict(out_channels=128, int_channels=256, num_upscalers=2, upscale_channels=256)
    dec = LO.SegDecoderOracle(**kw)
    assert unet_init.seg_decoder_param_shapes(**kw) == {k: tuple(v.shape) for k, v in dec.state_dict().items()}
    sd = unet_init.random_seg_decoder_state_dict(seed=3, **kw)
    dec.load_state_dict(sd, strict=True)
    assert sum(v.numel() for v in sd.values()) == 830848  # 0.830848 M params


def test_scheduler_mirror_host_schedule_matches_reference():
    s = DDIMNoiseScheduler(**SCHED_KW)
    z = np.load(os.path.join(G, "ddim_steps.npz"))
    assert np.array_equal(s.alphas_cumprod.numpy(), z["alphas_cumprod"])
    for T in (10, 50):
        s.set_timesteps_inference(T)
        assert s.timesteps.tolist() == GOLD["scheduler"][f"timesteps_{T}"]
    s.set_timesteps_inference(50)
    coef = s.step_coefficients()
    o = LO.DDIMOracle()
    o.set_timesteps_inference(50)
    for t in (999, 499, 19, 0):
        assert [float(c) for c in o.coefficients(t)] == coef[t].tolist()
    for T in (50, 10, 7, 1000):  # the vectorised table against the reference's per-timestep 0-dim tensor ops, every row
        s.set_timesteps_inference(T)
        o.set_timesteps_inference(T)
        ref = torch.stack([torch.stack(o.coefficients(t)) for t in range(1000)])
        assert torch.equal(s.step_coefficients(), ref), T
    s.set_timesteps_inference(50)
    # the fused kernel's formula, evaluated with torch on the CPU, reproduces the reference's step bit for bit
    eps, x = torch.from_numpy(z["eps"]), torch.from_numpy(z["x"])
    for t in (999, 499, 19):
        c = coef[t]
        x0 = (x - c[0] * eps) / c[1]
        prev = c[2] * x0 + c[3] * eps
        assert np.array_equal(x0.numpy(), z[f"x0_{t}"]) and np.array_equal(prev.numpy(), z[f"prev_{t}"])
    assert s.init_noise_sigma == 1.0 and len(s) == 1000 and s.weights.shape == (1000,)


def test_scheduler_mirror_rejects_unbuilt_modes_and_cpu_tensors():
    s = DDIMNoiseScheduler(**SCHED_KW)
    s.set_timesteps_inference(10)
    with pytest.raises(L.LdmError):
        s.step(torch.zeros(4), 999, torch.zeros(4))  # CPU tensors: no fallback
    s2 = DDIMNoiseScheduler(**{**SCHED_KW, "prediction_type": "v_prediction"})
    s2.set_timesteps_inference(10)
    with pytest.raises(NotImplementedError):
        s2.step(torch.zeros(4), 999, torch.zeros(4))
    with pytest.raises(NotImplementedError):
        DDIMNoiseScheduler(beta_schedule="nope")


def test_modify_encoder_matches_oracle():
    cfg = dict(block_out_channels=(64, 128, 256, 256))
    torch.manual_seed(0)
    o = UO.UNetOracle(**cfg)
    sd = {k: v.clone() for k, v in o.state_dict().items()}
    m = UNet(device="cpu", **cfg)
    m.load_state_dict(sd)
    m.remove_cross_attention()
    torch.manual_seed(5)
    o.modify_encoder(in_channels=8, init_mode_seg="copy", init_mode_image="zero", cond_channels=4)
    torch.manual_seed(5)
    m.modify_encoder(in_channels=8, init_mode_seg="copy", init_mode_image="zero", cond_channels=4)
    assert m.conv_in.in_channels == 12 and m.conv_in.out_channels == 64
    assert torch.equal(m.state_dict()["conv_in.weight"], o.conv_in.weight.detach())
    assert torch.equal(m.state_dict()["new_conv.bias"], o.conv_in.bias.detach())
    assert bool((m.conv_in.weight[:, 4:8] == 0).all()) and torch.equal(m.conv_in.weight[:, :4], sd["conv_in.weight"])
    assert tuple(m.config.block_out_channels) == (64, 128, 256, 256)


def test_unet_parameter_count_matches_published_sd14():
    """Known answer that pins the restated architecture's shapes: Stable Diffusion 1.x's UNet2DConditionModel has
    859 520 964 parameters (the published "860M UNet"); without attn2 / norm2 (unet.py:83-105) 815 533 444 remain."""
    full = unet_init.unet_param_shapes(cross_attention_dim=768)
    assert sum(int(np.prod(v)) for v in full.values()) == 859520964
    removed = unet_init.unet_param_shapes()
    assert sum(int(np.prod(v)) for v in removed.values()) == 815533444
    ref = UO.UNetOracle(in_channels=4)
    assert sum(p.numel() for p in ref.parameters()) == 815533444


def test_vae_image_param_shapes_match_oracle_module():
    from oracle import vae_image_oracle as VO
    net = VO.VAEImageOracle()
    want = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    got = unet_init.vae_image_param_shapes()
    assert got == want and sum(int(np.prod(v)) for v in got.values()) == 34163664
    with_cross = unet_init.unet_param_shapes(block_out_channels=(64, 128, 256, 256), cross_attention_dim=768)
    ref = UO.UNetOracle(block_out_channels=(64, 128, 256, 256), cross_attention_dim=768)
    assert with_cross == {k: tuple(v.shape) for k, v in ref.state_dict().items()}


def test_unet_forward_without_gpu_fails_loudly():
    cfg = dict(block_out_channels=(64, 128, 256, 256))
    m = UNet(device="cpu", **cfg)
    m.load_state_dict(unet_init.random_unet_state_dict(0, in_channels=8, **cfg))
    with pytest.raises(L.LdmError):
        m(torch.zeros(1, 8, 8, 8), torch.tensor(999), encoder_hidden_states=None)
    with pytest.raises(L.LdmError):  # CPU tensors are refused before anything else, with or without a context
        m(torch.zeros(1, 8, 8, 8), torch.tensor(999), encoder_hidden_states=torch.zeros(1, 77, 768))


def _reduce_worker(rank, world, port, q):
    import torch.distributed as dist
    from video_latent_diffusion_panoptic_segmentation_b200.ldmseg.trainers.trainers_ldm_cond import reduce_evaluator_

    class Ev:
        pass
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    ev = Ev()
    ev.TP, ev.FP, ev.FN, ev.iou_sum = 3 + rank, 10 * (rank + 1), rank, 0.1 + rank * 0.7
    ev.TP_per_class = {11: 1 + rank, 3: 2} if rank == 0 else {11: 1 + rank, 7: 1}
    ev.FP_per_class = {11: 4, 3: 6} if rank == 0 else {11: 5, 7: 15, 2: 1}
    ev.FN_per_class = {11: 0, 3: 0} if rank == 0 else {11: 1, 7: 0}
    ev.iou_sum_per_class = {11: 0.05, 3: 0.05} if rank == 0 else {11: 0.4, 7: 0.4}
    reduce_evaluator_(ev, torch.device("cpu"))
    q.put((rank, ev.TP, ev.FP, ev.FN, ev.iou_sum, ev.TP_per_class, ev.FP_per_class, ev.FN_per_class,
           ev.iou_sum_per_class))
    dist.barrier()
    dist.destroy_process_group()


def test_stats_reduction_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_reduce_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in res:
        _, tp, fp, fn, iou, tpc, fpc, fnc, iouc = r
        assert (tp, fp, fn) == (7, 30, 1)
        assert iou == (0.0 + 0.1) + (0.1 + 1 * 0.7)  # summed in rank order
        assert tpc == {3: 2, 7: 1, 11: 3}
        assert fpc[11] == 9 and fpc[7] == 15 and fpc[3] == 6 and fpc[2] == 1
        assert fnc == {3: 0, 7: 0, 11: 1}
        assert iouc[11] == 0.05 + 0.4
    assert res[0][1:] == res[1][1:]


# ------------------------------------------------------------------------------------------- sharded clip DVPQ (8e)
def _clip(n_frames, H=40, W=56, seed=3):
    """Synthetic clip: per frame (pred_cat, pred_ins, gt_cat, gt_ins) int32 maps with drifting Voronoi tubes."""
    from synth import synth_panoptic
    rng = np.random.default_rng(seed)
    _, cat0, ins0 = synth_panoptic(rng, H, W + 2 * n_frames, n_seeds=14)
    out = []
    for f in range(n_frames):
        gc, gi = cat0[:, 2 * f:2 * f + W].copy(), ins0[:, 2 * f:2 * f + W].copy()
        pc, pi = np.roll(gc, (1, 2), axis=(0, 1)).copy(), np.roll(gi, (1, 2), axis=(0, 1)).copy()
        flip = rng.random((H, W)) < 0.04
        pc[flip], pi[flip] = 5, 7
        pc[pc == 255], pi[pc == 255] = 3, 1
        out.append((pc, pi, gc, gi))
    return [np.stack([fr[j] for fr in out]).astype(np.int32) for j in range(4)]


def _oracle_eval_fn(pc, pi, gc, gi, dp, dg):
    from oracle import eval_oracle as EO
    k = pc.shape[0]
    return EO.dvpq_window([pc[j].numpy() for j in range(k)], [pi[j].numpy() for j in range(k)],
                          [gc[j].numpy() for j in range(k)], [gi[j].numpy() for j in range(k)])


def _clip_worker(rank, world, port, n_frames, k, q):
    import torch.distributed as dist
    from video_latent_diffusion_panoptic_segmentation_b200.eval import clip_dvpq as CD
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    maps = _clip(n_frames)
    lo, hi = CD.shard_range(n_frames, rank, world)
    local = [torch.from_numpy(m[lo:hi]) for m in maps]
    res = CD.dvpq_clip_sharded(*local, n_frames=n_frames, eval_frames=k, eval_fn=_oracle_eval_fn)
    q.put((rank, res["pq"], res["iou"].tolist(), res["tp"].tolist(), res["fn"].tolist(), res["fp"].tolist(),
           res["n_windows"]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_frames,k", [(2, 8, 2), (3, 7, 3), (4, 5, 3)])
def test_clip_dvpq_sharded_gloo_matches_single_process(world, n_frames, k):
    """frames sharded contiguously, the k-1 frame halo from the following shard(s) (a halo may span several
    short or empty shards), windows evaluated by the owner of their first frame, integer counts all-reduced, float rows
    gathered in window order: bit-identical to the reference's single-process aggregation (eval_dvpq.py:153-210)."""
    from oracle import eval_oracle as EO
    from video_latent_diffusion_panoptic_segmentation_b200.ldmseg.evaluations.new_eval import aggregate
    maps = _clip(n_frames)
    rows = [EO.dvpq_window(*[[m[i + j] for j in range(k)] for m in maps]) for i in range(n_frames - k + 1)]
    want = aggregate(rows)
    assert want["tp"].sum() > 0 and want["fp"].sum() + want["fn"].sum() > 0
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + 7 * world + n_frames) % 2000
    procs = [ctx.Process(target=_clip_worker, args=(r, world, port, n_frames, k, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in res:
        assert r[1] == want["pq"] and r[2] == want["iou"].tolist()  # bit-identical float64
        assert r[3] == want["tp"].tolist() and r[4] == want["fn"].tolist() and r[5] == want["fp"].tolist()
        assert r[6] == n_frames - k + 1


def test_shard_range_covers_every_frame_once():
    from video_latent_diffusion_panoptic_segmentation_b200.eval.clip_dvpq import shard_range
    for n in (1, 5, 8, 1101):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                lo, hi = shard_range(n, r, world)
                seen += list(range(lo, hi))
            assert seen == list(range(n))


# ------------------------------------------------------------------------------------------- host logic of the 8f rows
def test_vae_image_state_dict_key_translation_and_guards():
    """GeneralVAEImage.load_state_dict: decoder / post_quant_conv keys are dropped (main_ldm.py:139), the pre-0.15
    diffusers AttentionBlock names map onto to_q / to_k / to_v / to_out.0; compute without a GPU fails loudly."""
    from oracle import vae_image_oracle as VO
    from video_latent_diffusion_panoptic_segmentation_b200.ldmseg.models import GeneralVAEImage
    sd = VO.VAEImageOracle(block_out_channels=(64, 64, 128, 128)).state_dict()
    old = {}
    for k, v in sd.items():
        for new_name, old_name in (("to_q", "query"), ("to_k", "key"), ("to_v", "value"), ("to_out.0", "proj_attn")):
            if f".attentions.0.{new_name}." in k:
                k = k.replace(f".{new_name}.", f".{old_name}.")
        old[k] = v
    old["decoder.conv_in.weight"] = torch.zeros(1)
    old["post_quant_conv.weight"] = torch.zeros(1)
    m = GeneralVAEImage(device="cpu", block_out_channels=(64, 64, 128, 128))
    m.load_state_dict(old)
    assert sorted(m.state_dict()) == sorted(sd)
    m.set_scaling_factor(0.5)
    assert m.scaling_factor == 0.5
    with pytest.raises(L.LdmError):
        m.encode(torch.zeros(1, 3, 64, 64))
    with pytest.raises(NotImplementedError):
        GeneralVAEImage(device="cpu", block_out_channels=(32, 64))


def test_trainer_descriptor_branches_on_cpu():
    """TrainerDiffusion._descriptors (trainers_ldm_cond.py:1100-1122): text encoder -> (text, uncond) contexts of the
    tokenizer's max length; image descriptor model -> one context, [B, C, n] -> [B, n, C]; p_get mirrors rgb_size."""
    from types import SimpleNamespace
    from video_latent_diffusion_panoptic_segmentation_b200.ldmseg.trainers import TrainerDiffusion
    from video_latent_diffusion_panoptic_segmentation_b200.ldmseg.trainers.trainers_ldm_cond import p_get

    class Tok:
        model_max_length = 6

        def __call__(self, texts, padding=None, max_length=None, truncation=True, return_tensors="pt"):
            ids = torch.zeros((len(texts), max_length), dtype=torch.long)
            for i, t in enumerate(texts):
                ids[i, :min(len(t), max_length)] = 1
            return SimpleNamespace(input_ids=ids)

    class Enc(torch.nn.Module):
        def forward(self, ids):
            return (torch.nn.functional.one_hot(ids, 4).float(),)

    class ToyClipVision(torch.nn.Module):
        def forward(self, x):
            assert x.shape[-2:] == (224, 224)  # norm_resize_images: 'clip' in the class name
            return {"last_feat": torch.ones((x.shape[0], 5, 7))}

    tr = TrainerDiffusion(p={}, tokenizer=Tok(), text_encoder=Enc(), args={"gpu": "cpu"})
    ctx, unc = tr._descriptors(["ab", "abcd"], None, 2)
    assert ctx.shape == unc.shape == (2, 6, 4) and unc[..., 0].all() and not torch.equal(ctx, unc)
    tr = TrainerDiffusion(p={}, image_descriptor_model=ToyClipVision(), args={"gpu": "cpu"})
    ctx, unc = tr._descriptors([""], torch.rand(1, 3, 32, 48), 1)
    assert ctx.shape == (1, 7, 5) and unc is None
    assert p_get({"transformation_kwargs": {"size_rgb": 192}}, "rgb_size") == (192, 192)
    assert p_get({"rgb_size": (384, 1248)}, "rgb_size") == (384, 1248) and p_get({}, "rgb_size") is None


def test_scheduler_training_helpers_match_reference_fixture():
    """add_noise / remove_noise (ddim_scheduler.py:155-216) and compute_loss_weights (:97-117) of the mirror against
    vectors of the REAL reference scheduler (tests/golden/make_golden_scheduler_extra.py): bit-exact, on the CPU (these
    helpers are plain tensor math, not kernels). The reference's 'inverse_log_snr' mode raises; no vector for it."""
    z = np.load(os.path.join(G, "scheduler_extra.npz"))
    s = DDIMNoiseScheduler(**SCHED_KW)
    x0, noise, t = torch.from_numpy(z["x0"]), torch.from_numpy(z["noise"]), torch.from_numpy(z["t"])
    assert torch.equal(s.add_noise(x0, noise.clone(), t), torch.from_numpy(z["noisy"]))
    assert torch.equal(s.add_noise(x0, noise.clone(), t, scale=0.5), torch.from_numpy(z["noisy_scaled"]))
    assert torch.equal(s.remove_noise(torch.from_numpy(z["noisy"]), noise, t), torch.from_numpy(z["rec"]))
    assert torch.equal(s.remove_noise(torch.from_numpy(z["noisy_scaled"]), noise, t, scale=0.5),
                       torch.from_numpy(z["rec_scaled"]))
    assert float(s.init_noise_sigma) == float(z["init_noise_sigma"])
    assert float(s.final_alpha_cumprod) == float(z["final_alpha_cumprod"])
    for mode in ("none", "max_clamp_snr", "fixed", "linear"):
        kw = dict(SCHED_KW)
        kw["weight"] = mode
        w = DDIMNoiseScheduler(**kw, max_snr=5.0).weights
        assert np.array_equal(np.asarray(w, dtype=np.float64), z[f"weights_{mode}"]), mode


def test_posterior_matches_reference_fixture():
    """DiagonalGaussianDistribution mirror (vae.py:371-425) against vectors of the REAL reference class for every act_fn
    and clamp_output: mean, logvar (clamped at -30 / 20), std, var, kl, get_range and a seeded sample, bit-exact."""
    from video_latent_diffusion_panoptic_segmentation_b200.ldmseg.models.vae import DiagonalGaussianDistribution
    z = np.load(os.path.join(G, "posterior.npz"))
    params = torch.from_numpy(z["params"])
    for act in ("none", "sigmoid", "tanh", "clip"):
        for clamp in (False, True):
            d = DiagonalGaussianDistribution(params.clone(), clamp_output=clamp, act_fn=act)
            k = f"{act}_{int(clamp)}"
            for name, got in (("mean", d.mean), ("logvar", d.logvar), ("std", d.std), ("var", d.var), ("kl", d.kl())):
                assert np.array_equal(got.numpy(), z[f"{k}_{name}"]), (k, name)
            assert np.array_equal(d.sample(generator=torch.Generator().manual_seed(5)).numpy(), z[k + "_sample"]), k
            r = d.get_range()
            assert [float(r.min), float(r.max)] == z[k + "_range"].tolist() and torch.equal(d.mode(), d.mean)
    with pytest.raises(NotImplementedError):
        DiagonalGaussianDistribution(params, act_fn="relu")


def test_output_dict_types_keep_items_and_attributes_in_sync():
    """T1 (ldmseg/utils/utils.py:26-31 and its subclasses unet.py:20-21, ddim_scheduler.py:21-23, vae.py:22-33):
    OrderedDict subclasses whose __setitem__ mirrors keys to attributes; construction by keyword goes through it."""
    from collections import OrderedDict
    from video_latent_diffusion_panoptic_segmentation_b200.ldmseg.models.unet import UNetOutput
    from video_latent_diffusion_panoptic_segmentation_b200.ldmseg.models.vae import EncoderOutput, RangeDict, VAEOutput
    from video_latent_diffusion_panoptic_segmentation_b200.ldmseg.schedulers.ddim_scheduler import DDIMNoiseSchedulerOutput
    from video_latent_diffusion_panoptic_segmentation_b200.ldmseg.utils import OutputDict
    t = torch.arange(3.0)
    o = UNetOutput(sample=t)
    assert isinstance(o, OutputDict) and isinstance(o, OrderedDict) and o.sample is t and o["sample"] is t
    o["extra"] = 5
    assert o.extra == 5 and list(o.keys()) == ["sample", "extra"]
    s = DDIMNoiseSchedulerOutput(prev_sample=t, pred_original_sample=t + 1)
    assert s.prev_sample is t and torch.equal(s["pred_original_sample"], t + 1)
    assert tuple(s.values())[0] is t  # positional unpacking order = insertion order, as callers rely on
    assert EncoderOutput(latent_dist="d").latent_dist == "d"
    v = VAEOutput(sample=t, posterior=None)
    assert v.sample is t and v["posterior"] is None
    assert RangeDict(min=t.min(), max=t.max()).max == 2.0


def test_fold_upsample_conv3x3_algebra():
    """ops.fold_upsample_conv3x3 (ldm_gemm_desc.up2): the four 2x2 sub-pixel kernels reproduce
    F.interpolate(scale 2, nearest) + Conv2d(3x3, padding 1) exactly (fp32 weights; the packed ones are bf16)."""
    import torch.nn.functional as F
    from video_latent_diffusion_panoptic_segmentation_b200 import ops
    torch.manual_seed(3)
    H, W, C, N = 5, 7, 8, 6
    x, w, b = torch.randn(2, C, H, W), torch.randn(N, C, 3, 3), torch.randn(N)
    ref = F.conv2d(F.interpolate(x, scale_factor=2, mode="nearest"), w, b, padding=1)
    w4, b4 = ops.fold_upsample_conv3x3(w.permute(0, 2, 3, 1).contiguous(), b)
    assert w4.shape == (4 * N, 4 * C) and w4.dtype == torch.bfloat16 and torch.equal(b4, b.repeat(4))
    # the same sums in fp32 (the bf16 rounding of the packed weights is covered by the GPU check)
    wf = w.permute(0, 2, 3, 1)
    rows = {0: ([0], [1, 2]), 1: ([0, 1], [2])}
    xp = F.pad(x, (1, 1, 1, 1))
    out = torch.zeros(2, N, 2 * H, 2 * W)
    for a in (0, 1):
        for bb in (0, 1):
            k = torch.stack([sum(wf[:, ky, kx] for ky in rows[a][i] for kx in rows[bb][j])
                             for i in (0, 1) for j in (0, 1)], dim=1).reshape(N, 2, 2, C).permute(0, 3, 1, 2)
            cls = 2 * a + bb
            assert torch.allclose(w4[cls * N:(cls + 1) * N].float().reshape(N, 2, 2, C).permute(0, 3, 1, 2), k,
                                  atol=2e-2, rtol=1e-2)   # bf16 rounding of the packed class weights
            # tap (i, j) reads input (y + i - 1 + a, x + j - 1 + b): rows a .. a + H of the zero-padded input
            out[:, :, a::2, bb::2] = F.conv2d(xp[:, :, a:a + H + 1, bb:bb + W + 1], k, b)
    assert torch.allclose(out, ref, atol=1e-4, rtol=1e-4)


def test_fold_layernorm_algebra():
    """ops.fold_layernorm (ldm_gemm_desc.ln_stats): rstd (x W'^T - mean g) + b' == LayerNorm(x) W^T + b."""
    import torch.nn.functional as F
    from video_latent_diffusion_panoptic_segmentation_b200 import ops
    torch.manual_seed(4)
    M, C, N = 12, 64, 24
    x = torch.randn(M, C) * 2 + 0.7
    w, b = torch.randn(N, C) * 0.1, torch.randn(N)
    gamma, beta = torch.rand(C) + 0.5, torch.randn(C) * 0.2
    ref = F.layer_norm(x, (C,), gamma, beta, 1e-5) @ w.t() + b
    w2, b2, g = ops.fold_layernorm(w, b, gamma, beta)
    mean, var = x.mean(-1, keepdim=True), x.var(-1, unbiased=False, keepdim=True)
    rstd = (var + 1e-5).rsqrt()
    got = rstd * (x @ w2.float().t() - mean * g[None, :]) + b2[None, :]
    assert torch.allclose(got, ref, atol=3e-2, rtol=2e-2)   # w2 is rounded to bf16; g is the row sum of the ROUNDED w2
    assert torch.equal(g, w2.float().sum(1))


def test_color_map_and_encode_seg_match_reference():
    """ldmseg/utils/utils.py:240-258 and trainers_ldm_cond.py:326-334 (what decode_latents returns without
    return_logits), against arrays of the real reference functions (tests/golden/make_golden_color_map.py)."""
    from video_latent_diffusion_panoptic_segmentation_b200.ldmseg.trainers import TrainerDiffusion
    from video_latent_diffusion_panoptic_segmentation_b200.ldmseg.utils import color_map
    z = np.load(os.path.join(G, "color_map.npz"))
    cm = color_map()
    assert cm.dtype == np.uint8 and np.array_equal(cm, z["cmap"])
    assert cm[:4].tolist() == [[0, 0, 0], [128, 0, 0], [0, 128, 0], [128, 128, 0]]   # the published PASCAL palette
    cn = color_map(normalized=True)
    assert cn.dtype == z["cmap_norm"].dtype and np.array_equal(cn, z["cmap_norm"])
    assert np.array_equal(color_map(N=19), z["cmap_19"])
    got = TrainerDiffusion.encode_seg(type("T", (), {"cmap": None})(), z["labels"])
    assert got.dtype == z["colours"].dtype and np.array_equal(got, z["colours"])
    two = np.stack([np.arange(256), np.arange(256)[::-1]], 1).astype(np.float32)     # a caller's own palette
    assert np.array_equal(TrainerDiffusion.encode_seg(None, z["labels"], cmap=two), two[z["labels"].astype(np.uint8)])


def test_load_unet_checkpoint_keeps_descriptor_mode_additions():
    """tools/main_ldm.py load_path branch: a checkpoint without object queries must not silently drop the ones
    define_learnable_embeddings added; one that carries them wins; 'remove' drops attn2 / norm2 again."""
    from video_latent_diffusion_panoptic_segmentation_b200.tools.main_ldm import load_unet_checkpoint
    cfg = dict(block_out_channels=(64, 128, 256, 256), cross_attention_dim=768)
    torch.manual_seed(1)
    sd = {k: v.clone() for k, v in UO.UNetOracle(**cfg).state_dict().items()}
    assert any(".attn2." in k for k in sd)
    m = UNet(device="cpu", **cfg)
    m.load_state_dict(sd)
    m.define_learnable_embeddings(16, 768)
    q0 = m.state_dict()["object_queries.weight"].clone()
    ckpt = {"module." + k: v + 1.0 for k, v in sd.items()}          # DDP-prefixed, different weights, no queries
    load_unet_checkpoint(m, ckpt, "learnable")
    got = m.state_dict()
    assert torch.equal(got["object_queries.weight"], q0) and m.has_cross_attention()
    assert torch.equal(got["conv_in.weight"], sd["conv_in.weight"] + 1.0)
    ckpt["module.object_queries.weight"] = torch.full_like(q0, 3.0)  # a checkpoint trained in that mode carries its own
    load_unet_checkpoint(m, ckpt, "learnable")
    assert torch.equal(m.state_dict()["object_queries.weight"], torch.full_like(q0, 3.0))
    m2 = UNet(device="cpu", **cfg)
    m2.load_state_dict(sd)
    m2.remove_cross_attention()
    load_unet_checkpoint(m2, {k: v for k, v in ckpt.items() if "object_queries" not in k}, "remove")
    assert not m2.has_cross_attention() and not any(".attn2." in k for k in m2.state_dict())
