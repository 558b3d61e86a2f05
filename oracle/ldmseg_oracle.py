"""ORACLE (test infrastructure only): CPU restatement of the reference's sampler path around the UNet.

Pinned against the real reference code, imported from /root/reference in the authoring container by
tests/golden/make_golden.py (fixtures under tests/golden/): the DDIM scheduler, the seg-AE decoder, the bit codec
(plus the reference's own sample_outputs/ PNGs). The sampling loop / merge bodies cannot be imported (the trainer
class needs datasets + DDP), so they are restated from the cited lines.

  DDIMOracle            ldmseg/schedulers/ddim_scheduler.py:51-95,119-136,218-269
  SegDecoderOracle      ldmseg/models/vae.py:124-173,268-272,310-323   (state-dict keys decoder.{0,2,3,5,6,8,10}.*)
  sample                ldmseg/trainers/trainers_ldm_cond.py:1048-1173
  decode_latents        ldmseg/trainers/trainers_ldm_cond.py:398-427
  logits_to_panoptic    ldmseg/trainers/trainers_ldm_cond.py:1286-1325
  encode/decode_bitmap  ldmseg/data/cityscapes.py:256-270 (kitti.py:292-306; coco.py:378-391 without the ==31 line)
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------------ scheduler
class DDIMOracle:
    def __init__(self, num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                 clip_sample=False, set_alpha_to_one=False, prediction_type="epsilon", clip_sample_range=1.0, **unused):
        if beta_schedule == "scaled_linear":
            betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        elif beta_schedule == "linear":
            betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        else:
            raise NotImplementedError(beta_schedule)
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.num_train_timesteps = num_train_timesteps
        self.prediction_type = prediction_type
        self.clip_sample, self.clip_sample_range = clip_sample, clip_sample_range
        self.init_noise_sigma = 1.0
        self.num_inference_steps = None
        self.timesteps = torch.arange(num_train_timesteps - 1, -1, -1, dtype=torch.int64)

    def set_timesteps_inference(self, num_inference_steps, tmin=0):
        ratio = self.num_train_timesteps // num_inference_steps
        self.num_inference_steps = num_inference_steps
        ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64) + (ratio - 1)
        ts = torch.from_numpy(ts)
        self.timesteps = ts[ts >= tmin]

    def coefficients(self, t):
        """(sqrt(1-a_t), sqrt(a_t), sqrt(a_prev), sqrt(1-a_prev)) as fp32 0-dim tensors, computed like :231-236."""
        t = int(t)
        prev_t = t - self.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        return (1 - a_t) ** 0.5, a_t ** 0.5, a_prev ** 0.5, (1 - a_prev) ** 0.5

    def step(self, model_output, timestep, sample, use_clipped_model_output=False):
        """ddim_scheduler.py:218-269 (epsilon prediction; clip_sample :253-257, use_clipped_model_output :259-261)."""
        assert self.prediction_type == "epsilon"
        s1m_at, s_at, s_ap, s1m_ap = self.coefficients(timestep)
        x0 = (sample - s1m_at * model_output) / s_at
        if self.clip_sample:
            x0 = x0.clamp(-self.clip_sample_range, self.clip_sample_range)
        if use_clipped_model_output:
            model_output = (sample - s_at * x0) / s1m_at
        prev = s_ap * x0 + s1m_ap * model_output
        return prev, x0


# ------------------------------------------------------------------------------------------------ seg-AE decoder
class LayerNorm2dOracle(nn.Module):
    def __init__(self, c, eps=1e-6):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(c))
        self.bias = nn.Parameter(torch.zeros(c))
        self.eps = eps

    def forward(self, x):
        u = x.mean(1, keepdim=True)
        s = (x - u).pow(2).mean(1, keepdim=True)
        x = (x - u) / torch.sqrt(s + self.eps)
        return self.weight[:, None, None] * x + self.bias[:, None, None]


class SegDecoderOracle(nn.Module):
    """decoder of GeneralVAESeg with num_mid_blocks=0 (base.yaml:14-33): indices 0..10 as in nn.Sequential."""

    def __init__(self, out_channels=128, int_channels=256, latent_channels=4, norm_num_groups=32, num_upscalers=2,
                 upscale_channels=256, scaling_factor=0.2, block_out_channels=(32, 64, 128, 256), **unused):
        super().__init__()
        layers = [nn.Conv2d(latent_channels, int_channels, 3, padding=1), nn.Identity()]
        dim = upscale_channels
        for i in range(num_upscalers):
            layers += [nn.ConvTranspose2d(int_channels if i == 0 else dim, dim, 2, stride=2), LayerNorm2dOracle(dim),
                       nn.SiLU()]
        layers += [nn.GroupNorm(norm_num_groups, dim), nn.SiLU(), nn.Conv2d(dim, out_channels, 3, padding=1)]
        self.decoder = nn.Sequential(*layers)
        self.scaling_factor = scaling_factor
        self.downsample_factor = 2 ** (len(block_out_channels) - 1)
        self.interpolation_factor = self.downsample_factor // (2 ** num_upscalers)

    def decode(self, z, interpolate=True):
        x = self.decoder(z)
        if interpolate:
            x = F.interpolate(x, scale_factor=self.interpolation_factor, mode="bilinear", align_corners=False)
        return x


class SegEncoderOracle(nn.Module):
    """encoder of GeneralVAESeg (vae.py:175-240: default topology, num_mid_blocks=0, gaussian posterior) with the
    reference's nn.Sequential indices, so its state_dict keys are the reference's ``encoder.N.*``. ``encode`` returns the
    moments; mean / logvar / std follow DiagonalGaussianDistribution (vae.py:385-392). Pinned bit-exactly to
    tests/golden/seg_encoder_small.npz (made by the real class)."""

    def __init__(self, in_channels=16, block_out_channels=(32, 64, 128, 256), int_channels=256, norm_num_groups=32,
                 latent_channels=4, num_latents=2, **unused):
        super().__init__()
        boc = block_out_channels
        layers = [nn.Conv2d(in_channels, boc[0], 3, padding=1), nn.SiLU()]
        for i in range(len(boc) - 1):
            layers += [nn.Conv2d(boc[i], boc[i], 3, padding=1), nn.Conv2d(boc[i], boc[i + 1], 3, padding=1, stride=2),
                       nn.SiLU()]
        layers += [nn.Conv2d(boc[-1], int_channels, 3, padding=1), nn.Identity(),
                   nn.GroupNorm(norm_num_groups, int_channels, eps=1e-6), nn.SiLU(),
                   nn.Conv2d(int_channels, latent_channels * num_latents, 3, padding=1)]
        self.encoder = nn.Sequential(*layers)

    def encode(self, semseg):
        return self.encoder(semseg)

    @staticmethod
    def posterior(moments):
        mean, logvar = torch.chunk(moments, 2, dim=1)
        logvar = torch.clamp(logvar, -30.0, 20.0)
        return mean, logvar, torch.exp(0.5 * logvar)


def build_seg_decoder(seed=0, **kw):
    torch.manual_seed(seed)
    return SegDecoderOracle(**kw).eval()


# ------------------------------------------------------------------------------------------------ sampler
@torch.no_grad()
def sample(unet, scheduler, rgb_latents, num_inference_steps=50, seed=None, self_condition=False,
           return_all_latents=False, noise=None, context=None, uncond_context=None, guidance_scale=7.5, trace=None):
    """trainers_ldm_cond.py:1048-1173. Default: image_descriptors=remove (no guidance, multiplier 1). With `context`
    [B,L,dim] the UNet gets encoder_hidden_states; with `uncond_context` too, the batch is doubled [uncond | text]
    and the prediction is uncond + guidance_scale * (text - uncond) (:1110-1122,1129,1147-1149).
    Noise is (B,4,h,w) from rgb_latents.shape[-2:], drawn from a CPU generator as in :1091-1095."""
    scheduler.set_timesteps_inference(num_inference_steps)
    B, _, h, w = rgb_latents.shape
    if noise is None:
        gen = torch.Generator().manual_seed(seed) if seed is not None else None
        noise = torch.randn((B, 4, h, w), generator=gen)
    latents = noise.to(rgb_latents.device) * scheduler.init_noise_sigma
    condition = torch.zeros_like(rgb_latents)
    steps = list(scheduler.timesteps)
    all_latents = []
    for i, t in enumerate(steps):
        if uncond_context is not None:
            assert not self_condition
            inp = torch.cat([torch.cat([latents] * 2), torch.cat([rgb_latents] * 2)], dim=1).float()
            eps2 = unet(inp, t, encoder_hidden_states=torch.cat([uncond_context, context]).float())
            eps_uncond, eps_text = eps2.chunk(2)
            eps = eps_uncond + guidance_scale * (eps_text - eps_uncond)
        else:
            parts = [latents, rgb_latents] + ([condition] if self_condition else [])
            eps = unet(torch.cat(parts, dim=1).float(), t, encoder_hidden_states=context)
        if trace is not None:  # (x_t, epsilon) of every step, for the per-step error curves of the parity tests
            trace.append((latents.clone(), eps.clone()))
        prev, x0 = scheduler.step(eps, t, latents)
        if self_condition:
            condition = x0
        latents = x0 if i == len(steps) - 1 else prev
        if return_all_latents:
            all_latents.append(latents)
    return torch.cat(all_latents, 0) if return_all_latents else latents


@torch.no_grad()
def decode_latents(vae, latents):
    """trainers_ldm_cond.py:423-427 with return_logits=True: logits [B, out, 8h, 8w] fp32."""
    return vae.decode(latents * (1.0 / vae.scaling_factor)).float()


def logits_to_panoptic(logits, mask_th=0.5, count_th=512, overlap_th=0.5, ignore_label=127):
    """One image: logits [C,H,W] fp32 -> (panoptic_pred int64 [H,W] before the merge, cleaned_pred int64 [H,W]
    with -1 = void, kept labels). trainers_ldm_cond.py:1286-1325 (threshold_mode 'max')."""
    pred = torch.argmax(logits, dim=0)
    probs = F.softmax(logits, dim=0).max(dim=0)[0]
    pred[probs < mask_th] = ignore_label
    pred = pred.cpu().numpy()
    sig = torch.sigmoid(logits).cpu().numpy()
    cleaned = pred.copy()
    kept = []
    for label, count in zip(*np.unique(pred, return_counts=True)):
        if count < count_th or label in {-1, ignore_label}:
            cleaned[cleaned == label] = -1
            continue
        original = sig[label] >= mask_th
        with np.errstate(divide="ignore"):
            if (pred == label).sum() / original.sum() < overlap_th:
                cleaned[cleaned == label] = -1
                continue
        kept.append(int(label))
    return pred, cleaned, kept


@torch.no_grad()
def pipeline(unet, vae, rgb_latents, num_inference_steps, seed, mask_th=0.5, count_th=512, overlap_th=0.5,
             ignore_label=127, scheduler=None, trace=None):
    """The whole reference path for one batch (trainers_ldm_cond.py:1222-1325 with identity resizes): sample ->
    decode_latents -> per image argmax / threshold / merge. Returns dict(latents, ids, cleaned) with int64 [B,H,W] maps."""
    scheduler = scheduler or DDIMOracle()
    latents = sample(unet, scheduler, rgb_latents, num_inference_steps=num_inference_steps, seed=seed, trace=trace)
    logits = decode_latents(vae, latents)
    ids, cleaned = [], []
    for b in range(logits.shape[0]):
        p, c, _ = logits_to_panoptic(logits[b], mask_th, count_th, overlap_th, ignore_label)
        ids.append(p)
        cleaned.append(c)
    return {"latents": latents, "ids": np.stack(ids), "cleaned": np.stack(cleaned)}


# ------------------------------------------------------------------------------------------------ bit codec
def encode_bitmap(x, n, ignore_label, fill_value=0.5):
    """x int64 [H,W] -> float [n,H,W] (cityscapes.py:256-261)."""
    ignore = x == ignore_label
    bits = torch.remainder(torch.bitwise_right_shift(x, torch.arange(n)[:, None, None]), 2).float()
    bits[:, ignore] = fill_value
    return bits, ignore


def decode_bitmap(x, quirk31=True):
    """x float [n,H,W] -> int64 [H,W] (cityscapes.py:263-270); quirk31=False is the coco.py variant."""
    b = (x > 0.).float()
    ids = (b * 2 ** torch.arange(b.shape[0])[:, None, None]).sum(0).long()
    if quirk31:
        ids[ids == 31] = 0
    return ids
