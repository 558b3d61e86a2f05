"""DDIM steps of the REAL reference scheduler with clip_sample on (the constructor default; base.yaml:55 turns it off)
and with use_clipped_model_output, for the kernel's clamp branch (ldm_ddim_step_clip).

    python tests/golden/make_golden_ddim_clip.py        (authoring container only: needs /root/reference)
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref_stubs  # noqa: E402

ref_stubs.install()
sys.path.insert(0, "/root/reference")
from ldmseg.schedulers import DDIMNoiseScheduler  # noqa: E402

KW = dict(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
          set_alpha_to_one=False, steps_offset=1, prediction_type="epsilon", thresholding=False, weight="none",
          verbose=False)


def main():
    g = torch.Generator().manual_seed(321)
    eps, x = torch.randn((2, 4, 6, 10), generator=g), 1.5 * torch.randn((2, 4, 6, 10), generator=g)
    out = {"eps": eps.numpy(), "x": x.numpy()}
    for rng in (1.0, 0.5):
        s = DDIMNoiseScheduler(clip_sample=True, clip_sample_range=rng, **KW)
        s.set_timesteps_inference(50)
        for t in (999, 499, 19):
            for ucm in (False, True):
                o = s.step(eps, torch.tensor(t), x, use_clipped_model_output=ucm)
                tag = f"r{rng}_t{t}_u{int(ucm)}"
                out["prev_" + tag] = o.prev_sample.numpy()
                out["x0_" + tag] = o.pred_original_sample.numpy()
    np.savez_compressed(os.path.join(HERE, "ddim_steps_clip.npz"), **out)
    print("wrote ddim_steps_clip.npz:", len(out), "arrays")


if __name__ == "__main__":
    main()
