"""decode_latents without return_logits (trainers_ldm_cond.py:428-436): the argmax label map leaves as a colour image
through the reference's palette. (Sorted last on purpose: everything it builds on is tested before it.)"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_decode_latents_returns_the_palette_image():
    from video_latent_diffusion_panoptic_segmentation_b200.ldmseg.models import GeneralVAESeg, unet_init
    from video_latent_diffusion_panoptic_segmentation_b200.ldmseg.schedulers import DDIMNoiseScheduler
    from video_latent_diffusion_panoptic_segmentation_b200.ldmseg.trainers import TrainerDiffusion
    from video_latent_diffusion_panoptic_segmentation_b200.ldmseg.utils import color_map
    kw = dict(in_channels=16, int_channels=256, out_channels=128, num_upscalers=2, upscale_channels=256,
              scaling_factor=0.2)
    vae = GeneralVAESeg(**kw, device=DEV)
    vae.load_state_dict(unet_init.random_seg_decoder_state_dict(seed=1, out_channels=128, int_channels=256,
                                                                num_upscalers=2, upscale_channels=256))
    sched = DDIMNoiseScheduler(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012,
                               beta_schedule="scaled_linear", clip_sample=False, set_alpha_to_one=False,
                               prediction_type="epsilon", weight="none")
    tr = TrainerDiffusion(p={"eval_kwargs": {"mask_th": 0.5, "count_th": 16, "overlap_th": 0.0}}, vae_semseg=vae,
                          unet_model=None, noise_scheduler=sched, args={"gpu": 0})
    lat = (0.6 * torch.randn((2, 4, 8, 16), generator=torch.Generator().manual_seed(3))).to(DEV)
    img = tr.decode_latents(lat)
    logits = tr.decode_latents(lat, return_logits=True)
    assert isinstance(img, np.ndarray) and img.dtype == np.uint8 and img.shape == (2, 64, 128, 3)
    logits = logits.cpu()                    # the checker runs on the host (first-index ties, as the kernel)
    pred = torch.argmax(logits, dim=1)
    assert np.array_equal(img, color_map()[pred.numpy().astype(np.uint8)])
    # threshold_output (:430-433): pixels whose largest softmax probability is below mask_th take the ignore label
    img_t = tr.decode_latents(lat, threshold_output=True)
    probs = torch.softmax(logits, dim=1).max(dim=1)[0]
    pred_t = pred.clone()
    pred_t[probs < tr.mask_th] = tr.ignore_label
    want = color_map()[pred_t.numpy().astype(np.uint8)]
    # (the kernel forms the probability with IEEE expf / division in torch's order; a probability within one ulp of
    # the threshold may still land on the other side of it)
    assert float((img_t != want).any(axis=-1).mean()) < 1e-4
