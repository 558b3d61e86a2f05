"""Golden vectors of the scheduler's training-side helpers from the REAL reference DDIMNoiseScheduler (authoring container
only): add_noise / remove_noise (ddim_scheduler.py:155-216) and the loss weights (:97-117).

    python tests/golden/make_golden_scheduler_extra.py     -> tests/golden/scheduler_extra.npz
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref_stubs  # noqa: E402

ref_stubs.install()
from ldmseg.schedulers import DDIMNoiseScheduler  # noqa: E402

KW = dict(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", clip_sample=False,
          set_alpha_to_one=False, steps_offset=1, prediction_type="epsilon", thresholding=False, verbose=False)


def main():
    out = {}
    g = torch.Generator().manual_seed(77)
    x0, noise = torch.randn((3, 4, 6, 10), generator=g), torch.randn((3, 4, 6, 10), generator=g)
    t = torch.tensor([999, 400, 0])
    for mode in ("none", "max_clamp_snr", "inverse_log_snr", "fixed", "linear"):
        try:
            s = DDIMNoiseScheduler(**KW, weight=mode, max_snr=5.0)
        except Exception as e:  # a mode the reference does not implement
            print("weight mode", mode, "->", type(e).__name__)
            continue
        out[f"weights_{mode}"] = np.asarray(s.weights, dtype=np.float64) if s.weights is not None else np.zeros(0)
    s = DDIMNoiseScheduler(**KW, weight="none")
    noisy = s.add_noise(x0, noise, t)
    noisy_scaled = s.add_noise(x0, noise, t, scale=0.5)
    rec = s.remove_noise(noisy, noise, t)
    rec_scaled = s.remove_noise(noisy_scaled, noise, t, scale=0.5)
    out.update(x0=x0.numpy(), noise=noise.numpy(), t=t.numpy(), noisy=noisy.numpy(), noisy_scaled=noisy_scaled.numpy(),
               rec=rec.numpy(), rec_scaled=rec_scaled.numpy(), init_noise_sigma=np.float64(s.init_noise_sigma),
               final_alpha_cumprod=np.float64(float(s.final_alpha_cumprod)))
    np.savez_compressed(os.path.join(HERE, "scheduler_extra.npz"), **out)
    print(sorted(out))


if __name__ == "__main__":
    main()
