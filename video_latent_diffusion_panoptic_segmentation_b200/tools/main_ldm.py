"""B200 mirror of the sampling entry point ``tools/main_ldm.py`` (reference :31-240), eval_only branch.

    python -m video_latent_diffusion_panoptic_segmentation_b200.tools.main_ldm base.eval_only=True \
        base.sampling_kwargs.num_inference_steps=50 [base.load_path=ckpt.pt] [key=value ...]
    torchrun --nproc-per-node N -m ...tools.main_ldm ...         (one process per GPU, NCCL)

Same config key names as tools/configs/base/base.yaml (``vae_model_kwargs``, ``model_kwargs``,
``noise_scheduler_kwargs``, ``sampling_kwargs``, ``eval_kwargs``) and the same ``main_worker(gpu, ngpus_per_node,
cfg_dist, p, name)`` signature (:73-79). Hydra is not available here, so dotted ``a.b.c=value`` overrides are parsed
directly (YAML scalars). Training (everything outside ``eval_only``) is out of scope. Datasets on disk are out of
scope too: without ``load_path`` the models are random-init and the validation loader is synthetic.
"""
import copy
import os
import sys

import numpy as np
import torch
import yaml

from ..ldmseg.models import GeneralVAEImage, GeneralVAESeg, UNet
from ..ldmseg.models import unet_init
from ..ldmseg.schedulers import DDIMNoiseScheduler
from ..ldmseg.trainers import TrainerDiffusion
from ..ldmseg.utils import is_main_process

# defaults of tools/configs/base/base.yaml (+ datasets/cityscapes.yaml:5-7) that the sampling path reads
BASE = {
    "eval_only": True,
    "load_path": None,
    "image_scaling_factor": 0.18215,
    "vae_model_kwargs": dict(in_channels=16, int_channels=256, out_channels=128, block_out_channels=[32, 64, 128, 256],
                             latent_channels=4, num_latents=2, num_upscalers=2, upscale_channels=256,
                             norm_num_groups=32, scaling_factor=0.2, parametrization="gaussian", act_fn="none",
                             clamp_output=False, freeze_codebook=False, num_mid_blocks=0, fuse_rgb=False,
                             resize_input=False, skip_encoder=False, pretrained_path=None),
    "model_kwargs": dict(in_channels=8, init_mode_seg="copy", init_mode_image="zero", cond_channels=0,
                         separate_conv=False, separate_encoder=False, add_adaptor=False, init_mode_adaptor="random"),
    "noise_scheduler_kwargs": dict(prediction_type="epsilon", beta_schedule="scaled_linear", num_train_timesteps=1000,
                                   beta_start=0.00085, beta_end=0.012, steps_offset=1, clip_sample=False,
                                   set_alpha_to_one=False, thresholding=False, dynamic_thresholding_ratio=0.995,
                                   clip_sample_range=1.0, sample_max_value=1.0, weight="none", max_snr=5.0),
    "train_kwargs": dict(image_descriptors="remove", self_condition=False, weight_dtype="float32", fp16=False,
                         freeze_layers=["time_embedding"]),
    "sampling_kwargs": dict(num_inference_steps=50, guidance_scale=7.5, seed=0),
    "eval_kwargs": dict(mask_th=0.5, count_th=512, overlap_th=0.5, batch_size=16),
    "num_classes": 128,
    "ignore_label": 127,
    # synthetic validation set (no dataset on disk): frames of `synthetic.size`, `synthetic.frames` of them
    # from_images: the loader yields RGB frames in [0, 1] and the RGB VAE encoder produces the latents (main_ldm.py:138-140,
    # trainers_ldm_cond.py:1234-1239) instead of handing the latents over directly
    "synthetic": dict(frames=8, size=[384, 1248], batch_size=8, seed=1234, from_images=False),
}
DIST = {"world_size": 1, "rank": 0, "dist_url": "tcp://127.0.0.1:54288", "dist_backend": "nccl",
        "multiprocessing_distributed": False}


def apply_overrides(cfg, overrides):
    """hydra-style ``base.a.b=value`` / ``a.b=value`` overrides."""
    for ov in overrides:
        if "=" not in ov:
            raise ValueError(f"override '{ov}' is not of the form key=value")
        key, val = ov.split("=", 1)
        parts = key.split(".")
        if parts[0] == "base":
            parts = parts[1:]
        node = cfg
        for part in parts[:-1]:
            node = node.setdefault(part, {})
        node[parts[-1]] = yaml.safe_load(val)
    return cfg


def build_models(p, device, seed=0):
    """main_ldm.py:138-176: seg-AE, UNet (+ remove_cross_attention, modify_encoder), scheduler."""
    vae = GeneralVAESeg(**{k: v for k, v in p["vae_model_kwargs"].items()}, device=device)
    if p["vae_model_kwargs"].get("pretrained_path") is None:
        vk = p["vae_model_kwargs"]
        vae.load_state_dict(unet_init.random_seg_decoder_state_dict(
            seed=seed + 1, out_channels=vk["out_channels"], int_channels=vk["int_channels"],
            latent_channels=vk["latent_channels"], num_upscalers=vk["num_upscalers"],
            upscale_channels=vk["upscale_channels"]))
    unet = UNet(device=device)
    desc = p["train_kwargs"].get("image_descriptors", "remove")
    # stands in for from_pretrained (SD-1.4 has cross_attention_dim 768; without it the attn2 weights are not drawn)
    unet.load_state_dict(unet_init.random_unet_state_dict(seed=seed, in_channels=4,
                                                          cross_attention_dim=None if desc == "remove" else 768))
    if desc == "remove":        # descriptors.py:93-95
        unet.remove_cross_attention()
    elif desc == "learnable":   # descriptors.py:89-91: 128 learnable object queries are the cross-attention context
        torch.manual_seed(seed + 2)
        unet.define_learnable_embeddings(128, 768)
    else:
        raise NotImplementedError(f"image_descriptors={desc!r} needs a CLIP text / vision encoder, which is outside "
                                  "this path (SURVEY section 8(f) rank 4 covers the UNet's cross-attention itself)")
    torch.manual_seed(seed)
    unet.modify_encoder(**p["model_kwargs"])
    unet.freeze_layers(p["train_kwargs"].get("freeze_layers", []))
    sched = DDIMNoiseScheduler(**p["noise_scheduler_kwargs"], device=device)
    return vae, unet, sched


def build_vae_image(p, device, seed=0):
    """main_ldm.py:138-140: the RGB VAE (decoder dropped); random-init SD-1.4 VAE encoder without a checkpoint."""
    vim = GeneralVAEImage.from_pretrained(state_dict=unet_init.random_vae_image_state_dict(seed=seed + 3), device=device)
    vim.set_scaling_factor(p["image_scaling_factor"])
    return vim


def synthetic_batches(p, rank=0, world=1):
    """Synthetic validation shard of this rank: random RGB latents (N(0,1)*0.18215) + Voronoi ground truth."""
    sy = p["synthetic"]
    H, W = sy["size"]
    frames = list(range(sy["frames"]))[rank::world]
    rng = np.random.default_rng(7)
    for i in range(0, len(frames), sy["batch_size"]):
        idx = frames[i:i + sy["batch_size"]]
        g = torch.Generator().manual_seed(sy["seed"] + i + 1000 * rank)
        gt = np.stack([_voronoi_semantic(rng, H, W) for _ in idx])
        batch = {"semseg": torch.from_numpy(gt), "mask": torch.ones((len(idx), H, W), dtype=torch.bool),
                 "meta": [{"im_size": (H, W), "image_id": int(j)} for j in idx]}
        if sy.get("from_images", False):
            batch["image"] = torch.rand((len(idx), 3, H, W), generator=g)
        else:
            batch["rgb_latents"] = p["image_scaling_factor"] * torch.randn((len(idx), 4, H // 8, W // 8), generator=g)
        yield batch


def _voronoi_semantic(rng, H, W, n_seeds=40, n_cls=19):
    ys, xs = rng.integers(0, H, n_seeds), rng.integers(0, W, n_seeds)
    cls = rng.integers(0, n_cls, n_seeds)
    yy, xx = np.mgrid[0:H, 0:W]
    owner = ((yy[None] - ys[:, None, None]) ** 2 + (xx[None] - xs[:, None, None]) ** 2).argmin(0)
    sem = cls[owner].astype(np.int64)
    sem[rng.random((H, W)) < 0.05] = 0
    return sem


def main_worker(gpu, ngpus_per_node, cfg_dist, p, name="b200"):
    """main_ldm.py:73-240 (eval_only branch :220-233)."""
    world = int(os.environ.get("WORLD_SIZE", cfg_dist.get("world_size", 1)))
    rank = int(os.environ.get("RANK", cfg_dist.get("rank", 0)))
    torch.cuda.set_device(gpu)
    if world > 1 and not torch.distributed.is_initialized():
        torch.distributed.init_process_group(backend=cfg_dist.get("dist_backend", "nccl"), world_size=world, rank=rank,
                                             init_method="env://" if "MASTER_ADDR" in os.environ else cfg_dist["dist_url"])
    if not p.get("eval_only", True):
        raise NotImplementedError("only base.eval_only=True (sampling + PQ) is built; training is out of scope")
    device = torch.device("cuda", gpu)
    vae, unet, sched = build_models(p, device)
    vae_image = build_vae_image(p, device) if p["synthetic"].get("from_images", False) else None
    trainer = TrainerDiffusion(p=p, vae_image=vae_image, vae_semseg=vae, unet_model=unet, noise_scheduler=sched,
                               args={"gpu": gpu})
    if p.get("load_path"):
        data = torch.load(p["load_path"], map_location="cpu")
        unet.load_state_dict(data["unet"])
        if p["train_kwargs"].get("image_descriptors", "remove") == "remove":
            unet.remove_cross_attention()
        if "vae_image" in data and vae_image is not None:
            vae_image.load_state_dict(data["vae_image"])
        if "vae_semseg" in data:
            vae.load_state_dict({k.replace("module.", ""): v for k, v in data["vae_semseg"].items()})
    res = trainer.compute_metrics(["pq"], threshold_output=True, save_images=False, seed=42,
                                  dataloader=list(synthetic_batches(p, rank, world)),
                                  num_inference_steps=p["sampling_kwargs"]["num_inference_steps"])
    if world > 1:
        torch.distributed.barrier()
    if is_main_process():
        print({k: v for k, v in res.items() if k != "per_class"})
    return res


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    p = apply_overrides(copy.deepcopy(BASE), argv)
    gpu = int(os.environ.get("LOCAL_RANK", 0))
    return main_worker(gpu, torch.cuda.device_count(), dict(DIST), p)


if __name__ == "__main__":
    main()
