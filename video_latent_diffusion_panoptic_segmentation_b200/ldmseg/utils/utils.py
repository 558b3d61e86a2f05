"""Host-side helpers the sampler path uses from ldmseg/utils/utils.py: OutputDict (:26-31), rank helpers (:44-67),
gpu_gather (:76-81), color_map (:240-258). Training meters / LR schedules / visualisers are out of scope (SURVEY.md
section 2 row 9)."""
from collections import OrderedDict

import numpy as np
import torch
import torch.distributed as dist


class OutputDict(OrderedDict):
    """OrderedDict whose items are also attributes (``out.sample``), as in the reference."""

    def __setitem__(self, key, value):
        super().__setitem__(key, value)
        super().__setattr__(key, value)


def is_dist_avail_and_initialized() -> bool:
    return dist.is_available() and dist.is_initialized()


def get_world_size() -> int:
    return dist.get_world_size() if is_dist_avail_and_initialized() else 1


def get_rank() -> int:
    return dist.get_rank() if is_dist_avail_and_initialized() else 0


def is_main_process() -> bool:
    return get_rank() == 0


def gpu_gather(tensor: torch.Tensor) -> torch.Tensor:
    """all_gather along dim 0 (0-dim tensors are promoted to 1-dim first)."""
    if tensor.ndim == 0:
        tensor = tensor.clone()[None]
    if not is_dist_avail_and_initialized():
        return tensor.clone()
    outs = [torch.empty_like(tensor) for _ in range(dist.get_world_size())]
    dist.all_gather(outs, tensor.contiguous())
    return torch.cat(outs, dim=0)


def color_map(N: int = 256, normalized: bool = False) -> np.ndarray:
    """The PASCAL VOC label palette of ldmseg/utils/utils.py:240-258, [N, 3] uint8 (float32 in [0, 1] if normalized):
    bit 3j + k of the label index becomes bit 7 - j of channel k (k = 0, 1, 2 for r, g, b). All labels at once."""
    idx = np.arange(N, dtype=np.int64)
    cmap = np.zeros((N, 3), dtype=np.int64)
    for j in range(8):
        for k in range(3):
            cmap[:, k] |= ((idx >> (3 * j + k)) & 1) << (7 - j)
    if normalized:
        return (cmap.astype("float32") / 255).astype("float32")
    return cmap.astype("uint8")
