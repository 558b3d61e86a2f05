"""B200 mirror of the sampling entry point ``tools/main_ldm.py`` (reference :31-240), eval_only branch.

    python -m video_latent_diffusion_panoptic_segmentation_b200.tools.main_ldm base.eval_only=True \
        base.sampling_kwargs.num_inference_steps=50 [base.load_path=ckpt.pt] [key=value ...]
    torchrun --nproc-per-node N -m ...tools.main_ldm ...         (one process per GPU, NCCL)

Same config key names as tools/configs/base/base.yaml (``vae_model_kwargs``, ``model_kwargs``,
``noise_scheduler_kwargs``, ``sampling_kwargs``, ``eval_kwargs``) and the same ``main_worker(gpu, ngpus_per_node,
cfg_dist, p, name)`` signature (:73-79). Hydra is not available here, so dotted ``a.b.c=value`` overrides are parsed
directly (YAML scalars); ``--config-dir <reference>/tools/configs`` reads the reference's own YAML tree
(base/base.yaml overlaid with datasets/cityscapes.yaml, as main_ldm.py:41 does) instead of the built-in defaults.
Training (everything outside ``eval_only``) is out of scope. Datasets on disk are out of scope too: without
``load_path`` the models are random-init (+ the "trained-like" recipe of unet_init.py) and the validation loader is
synthetic.
"""
import copy
import os
import re
import sys

import numpy as np
import torch
import yaml

from ..ldmseg.data import synthetic as SY
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
    # recipe: "trained_like" (unet_init.TRAINED_LIKE: the tail is not degenerate, segments survive the merge and the
    # ground truth is derived from a teacher prediction) or "random" (plain torch-default init: every pixel ends up void)
    "synthetic": dict(frames=8, size=[384, 1248], batch_size=8, seed=1234, from_images=False, recipe="trained_like"),
}
# the keys of BASE that exist in the reference's YAML tree (tests/test_host_cpu.py checks them against the parsed files)
YAML_KEYS = ("eval_only", "load_path", "image_scaling_factor", "vae_model_kwargs", "model_kwargs",
             "noise_scheduler_kwargs", "sampling_kwargs", "eval_kwargs", "num_classes", "ignore_label")
TRAIN_KWARGS_KEYS = ("image_descriptors", "self_condition", "weight_dtype", "fp16", "freeze_layers")
DIST = {"world_size": 1, "rank": 0, "dist_url": "tcp://127.0.0.1:54288", "dist_backend": "nccl",
        "multiprocessing_distributed": False}


def load_reference_config(config_dir):
    """The reference's config tree as main_ldm.py:31-41 assembles it: ``base/base.yaml`` overlaid with
    ``datasets/cityscapes.yaml`` (``cfg_base | cfg_dataset``) and ``distributed/local.yaml``. base.yaml:177-178 reads
    ``train_db_name:cityscapes`` / ``val_db_name:cityscapes`` without the space YAML requires after a key (PyYAML
    rejects the file as it stands); those two lines are repaired before parsing, nothing else is touched.
    Returns (p, cfg_dist)."""
    def read(rel):
        with open(os.path.join(config_dir, rel)) as f:
            text = re.sub(r"(?m)^(\w+):(?=\S)", r"\1: ", f.read())
        return yaml.safe_load(text) or {}
    base, dataset = read("base/base.yaml"), read("datasets/cityscapes.yaml")
    p = dict(base)
    p.update(dataset)
    dist_path = os.path.join(config_dir, "distributed", "local.yaml")
    cfg_dist = read("distributed/local.yaml") if os.path.exists(dist_path) else dict(DIST)
    return p, cfg_dist


def from_reference_config(config_dir):
    """BASE with every key the sampling path reads replaced by the reference YAML's value (+ the synthetic section)."""
    ref, cfg_dist = load_reference_config(config_dir)
    p = copy.deepcopy(BASE)
    for k in YAML_KEYS:
        if k in ref:
            p[k] = copy.deepcopy(ref[k])
    p["train_kwargs"] = dict(ref.get("train_kwargs", BASE["train_kwargs"]))
    if "transformation_kwargs" in ref:
        p["transformation_kwargs"] = copy.deepcopy(ref["transformation_kwargs"])
    return p, cfg_dist


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


def recipe_of(p):
    return p.get("synthetic", {}).get("recipe", "random")


def build_models(p, device, seed=0):
    """main_ldm.py:138-176: seg-AE, UNet (+ remove_cross_attention, modify_encoder), scheduler. Without checkpoints the
    weights are random (torch defaults); with ``synthetic.recipe == "trained_like"`` the gains / smoothing of
    unet_init.TRAINED_LIKE are applied on top (the classifier head is fitted later, see fit_trained_like_head)."""
    trained_like = recipe_of(p) == "trained_like"
    vae = GeneralVAESeg(**{k: v for k, v in p["vae_model_kwargs"].items()}, device=device)
    if p["vae_model_kwargs"].get("pretrained_path") is None:
        vk = p["vae_model_kwargs"]
        vsd = unet_init.random_seg_decoder_state_dict(
            seed=seed + 1, out_channels=vk["out_channels"], int_channels=vk["int_channels"],
            latent_channels=vk["latent_channels"], num_upscalers=vk["num_upscalers"],
            upscale_channels=vk["upscale_channels"])
        if trained_like:
            unet_init.trained_like_seg_decoder_(vsd)
        vae.load_state_dict(vsd)
    unet = UNet(device=device)
    desc = p["train_kwargs"].get("image_descriptors", "remove")
    # stands in for from_pretrained (SD-1.4 has cross_attention_dim 768; without it the attn2 weights are not drawn)
    usd = unet_init.random_unet_state_dict(seed=seed, in_channels=4,
                                           cross_attention_dim=None if desc == "remove" else 768)
    if trained_like:
        unet_init.trained_like_unet_(usd)
    unet.load_state_dict(usd)
    if desc == "remove":        # descriptors.py:93-95
        unet.remove_cross_attention()
    elif desc == "learnable":   # descriptors.py:89-91: 128 learnable object queries are the cross-attention context
        torch.manual_seed(seed + 2)
        unet.define_learnable_embeddings(128, 768)
    else:
        raise NotImplementedError(f"image_descriptors={desc!r} needs a CLIP text / vision encoder, which is outside "
                                  "this path (SURVEY section 8(f) rank 4 covers the UNet's cross-attention itself)")
    torch.manual_seed(seed)
    mk = dict(p["model_kwargs"])
    if trained_like:  # the prediction follows the image (a trained model's image weights are not zero)
        mk["init_mode_image"] = unet_init.TRAINED_LIKE_MODEL_KWARGS["init_mode_image"]
    unet.modify_encoder(**mk)
    unet.freeze_layers(p["train_kwargs"].get("freeze_layers", []))
    sched = DDIMNoiseScheduler(**p["noise_scheduler_kwargs"], device=device)
    return vae, unet, sched


@torch.no_grad()
def fit_trained_like_head(trainer, latent_hw, num_inference_steps, seed=42, data_seed=1234):
    """Step 4 of the trained-like recipe: sample global frame 0 of the synthetic clip alone (batch 1: the same launch
    shapes, hence the same head, on every rank and at every world size), take the decoder features in front of the last
    conv and fit the nearest-centroid head to them. Returns the teacher latents."""
    vae = trainer.vae_semseg
    h, w = latent_hw
    rgb0 = SY.trained_like_rgb_latents(1, h, w, seed=data_seed).to(trainer.device)
    lat = trainer.sample([""], num_inference_steps, seed=seed, rgb_latents=rgb0)
    feats = vae.decode_features(lat, scale=1.0 / vae.scaling_factor)
    weight, bias = unet_init.fit_seg_head(feats, out_channels=vae.out_channels)
    vae.load_state_dict(unet_init.set_seg_head_(vae.state_dict(), weight, bias))
    return lat


def load_unet_checkpoint(unet, state_dict, image_descriptors="remove"):
    """main_ldm.py:205-214: the checkpoint's UNet weights replace the freshly built ones. UNet.load_state_dict swaps the
    whole state dict, so what the constructor path ADDED for the descriptor mode (descriptors.py:89-95) is re-applied:
    'remove' drops attn2 / norm2 again; 'learnable' keeps the object queries / encoder_hid_proj defined before the load
    when the checkpoint does not carry its own (a checkpoint trained in that mode does, and then wins)."""
    added = {k: v for k, v in unet.state_dict().items()
             if k.startswith(("object_queries.", "encoder_hid_proj."))} if image_descriptors == "learnable" else {}
    sd = {k.replace("module.", ""): v for k, v in state_dict.items()}
    for k, v in added.items():
        sd.setdefault(k, v)
    unet.load_state_dict(sd)
    if image_descriptors == "remove":
        unet.remove_cross_attention()
    return unet


def build_vae_image(p, device, seed=0):
    """main_ldm.py:138-140: the RGB VAE (decoder dropped); random-init SD-1.4 VAE encoder without a checkpoint."""
    vim = GeneralVAEImage.from_pretrained(state_dict=unet_init.random_vae_image_state_dict(seed=seed + 3), device=device)
    vim.set_scaling_factor(p["image_scaling_factor"])
    return vim


def synthetic_batches(p, rank=0, world=1):
    """Synthetic validation shard of this rank (frames rank, rank + world, ... as a DistributedSampler deals them,
    trainers_ldm_cond.py:246-247). Latents handed over directly: the drifting-cell clip of ldmseg/data/synthetic.py
    (recipe "trained_like") or N(0,1) * 0.18215 ("random"); from_images: RGB frames in [0, 1] for the VAE encoder.
    The ground truth is a Voronoi map here; main_worker replaces it by one derived from a teacher prediction when the
    recipe is "trained_like" (a Voronoi map never matches a prediction: TP = 0)."""
    sy = p["synthetic"]
    H, W = sy["size"]
    frames = list(range(sy["frames"]))[rank::world]
    rng = np.random.default_rng(7)
    trained_like = recipe_of(p) == "trained_like" and not sy.get("from_images", False)
    for i in range(0, len(frames), sy["batch_size"]):
        idx = frames[i:i + sy["batch_size"]]
        g = torch.Generator().manual_seed(sy["seed"] + i + 1000 * rank)
        gt = np.stack([_voronoi_semantic(rng, H, W) for _ in idx])
        batch = {"semseg": torch.from_numpy(gt), "mask": torch.ones((len(idx), H, W), dtype=torch.bool),
                 "meta": [{"im_size": (H, W), "image_id": int(j)} for j in idx]}
        if sy.get("from_images", False):
            batch["image"] = torch.rand((len(idx), 3, H, W), generator=g)
        elif trained_like:
            batch["rgb_latents"] = torch.cat([SY.trained_like_rgb_latents(1, H // 8, W // 8, seed=sy["seed"],
                                                                          first_frame=int(j)) for j in idx])
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
        load_unet_checkpoint(unet, data["unet"], p["train_kwargs"].get("image_descriptors", "remove"))
        if "vae_image" in data and vae_image is not None:
            vae_image.load_state_dict(data["vae_image"])
        if "vae_semseg" in data:
            vae.load_state_dict({k.replace("module.", ""): v for k, v in data["vae_semseg"].items()})
    T = p["sampling_kwargs"]["num_inference_steps"]
    batches = list(synthetic_batches(p, rank, world))
    if recipe_of(p) == "trained_like" and not p.get("load_path") and not p["synthetic"].get("from_images", False):
        H, W = p["synthetic"]["size"]
        fit_trained_like_head(trainer, (H // 8, W // 8), T, seed=42, data_seed=p["synthetic"]["seed"])
        for data in batches:   # teacher pass: the ground truth is a coarse, partly mislabelled copy of the prediction
            lat = trainer.sample([""] * data["rgb_latents"].shape[0], T, seed=42,
                                 rgb_latents=data["rgb_latents"].to(device))
            _, cleaned, _ = trainer.panoptic_ids(lat)
            data["semseg"] = SY.teacher_ground_truth(cleaned, min_area=min(4096, H * W // 64)).cpu()
    res = trainer.compute_metrics(["pq"], threshold_output=True, save_images=False, seed=42,
                                  dataloader=batches, num_inference_steps=T)
    if world > 1:
        torch.distributed.barrier()
    if is_main_process():
        print({k: v for k, v in res.items() if k != "per_class"})
    return res


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    cfg_dist = dict(DIST)
    if "--config-dir" in argv:
        i = argv.index("--config-dir")
        p, cfg_dist = from_reference_config(argv[i + 1])
        p["eval_only"] = True  # the only branch built here (base.yaml:8 defaults to training)
        del argv[i:i + 2]
    else:
        p = copy.deepcopy(BASE)
    p = apply_overrides(p, argv)
    gpu = int(os.environ.get("LOCAL_RANK", 0))
    return main_worker(gpu, torch.cuda.device_count(), cfg_dist, p)


if __name__ == "__main__":
    main()
