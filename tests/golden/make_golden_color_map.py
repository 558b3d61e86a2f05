"""The reference's label palette and encode_seg on a random label batch (ldmseg/utils/utils.py:240-258,
trainers_ldm_cond.py:326-334), for the mirrors' CPU test.

    python tests/golden/make_golden_color_map.py        (authoring container only: needs /root/reference)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref_stubs  # noqa: E402

ref_stubs.install()
sys.path.insert(0, "/root/reference")
import ast  # noqa: E402

from ldmseg.utils.utils import color_map  # noqa: E402


def reference_encode_seg():
    """The reference's TrainerDiffusion.encode_seg, executed from its source: importing the trainer module needs
    detectron2's comm, which the stubs do not model; the method itself only needs numpy and color_map."""
    path = "/root/reference/ldmseg/trainers/trainers_ldm_cond.py"
    tree = ast.parse(open(path).read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "TrainerDiffusion")
    fn = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "encode_seg")
    ns = {"np": np, "color_map": color_map}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), path, "exec"), ns)
    return ns["encode_seg"]


def main():
    rng = np.random.default_rng(7)
    labels = rng.integers(0, 128, size=(2, 9, 13)).astype(np.int64)
    labels[0, :2] = 127          # the ignore label of the default config
    labels[1, 3, 4] = 255
    labels[1, 5, 6] = -1         # astype(uint8) wraps it to 255
    colours = reference_encode_seg()(None, labels)   # the method never reads self
    np.savez_compressed(os.path.join(HERE, "color_map.npz"), cmap=color_map(), cmap_norm=color_map(normalized=True),
                        cmap_19=color_map(N=19), labels=labels, colours=colours)
    print("wrote color_map.npz", colours.shape, colours.dtype)


if __name__ == "__main__":
    main()
