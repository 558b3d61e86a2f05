"""Golden vectors of the seg-AE posterior from the REAL reference class (authoring container only):
ldmseg/models/vae.py:371-425 DiagonalGaussianDistribution -- mean / logvar / std / var / kl / mode / get_range / sample
for every act_fn and clamp_output setting.

    python tests/golden/make_golden_posterior.py     -> tests/golden/posterior.npz
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
from ldmseg.models.vae import DiagonalGaussianDistribution  # noqa: E402


def main():
    g = torch.Generator().manual_seed(91)
    params = 4.0 * torch.randn((2, 8, 5, 7), generator=g)
    params[0, 4:, 0, 0] = torch.tensor([-40.0, 30.0, 0.0, 19.9])  # logvar clamp at -30 / 20
    out = {"params": params.numpy()}
    for act in ("none", "sigmoid", "tanh", "clip"):
        for clamp in (False, True):
            d = DiagonalGaussianDistribution(params.clone(), clamp_output=clamp, act_fn=act)
            k = f"{act}_{int(clamp)}"
            out[k + "_mean"], out[k + "_logvar"] = d.mean.numpy(), d.logvar.numpy()
            out[k + "_std"], out[k + "_var"], out[k + "_kl"] = d.std.numpy(), d.var.numpy(), d.kl().numpy()
            out[k + "_sample"] = d.sample(generator=torch.Generator().manual_seed(5)).numpy()
            r = d.get_range()
            out[k + "_range"] = np.array([float(r.min), float(r.max)], np.float32)
            assert torch.equal(d.mode(), d.mean)
    np.savez_compressed(os.path.join(HERE, "posterior.npz"), **out)
    print(len(out), "arrays")


if __name__ == "__main__":
    main()
