"""ORACLE support (authoring container only): minimal sys.modules stubs so that the REAL reference modules
(`ldmseg.schedulers`, `ldmseg.models.vae`, evaluators) import from /root/reference without diffusers / detectron2 /
termcolor / easydict, none of which are installed here. Nothing is restated in this file: the stubs only satisfy
import statements (ldmseg/utils/utils.py:20-23, ldmseg/models/unet.py:12-17, ldmseg/models/vae.py:14-19).
Used only by tests/golden/make_golden.py; /root/reference does not exist on the GPU box.
"""
import importlib.util
import sys
import types

REFERENCE = "/root/reference"


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install():
    import torch.nn as nn

    class _Dummy(nn.Module):
        def __init__(self, *a, **k):
            super().__init__()

    if "detectron2" not in sys.modules:
        _mod("detectron2")
        _mod("detectron2.utils")
        _mod("detectron2.utils.visualizer", Visualizer=object, _PanopticPrediction=object, ColorMode=object,
             _OFF_WHITE=(1.0, 1.0, 1.0), _create_text_labels=lambda *a, **k: [])
    if "diffusers" not in sys.modules:
        _mod("diffusers", UNet2DConditionModel=_Dummy, AutoencoderKL=_Dummy)
        _mod("diffusers.models")
        _mod("diffusers.models.unet_2d_blocks", UNetMidBlock2D=_Dummy)
        _mod("diffusers.training_utils", EMAModel=object)
    if "termcolor" not in sys.modules:
        _mod("termcolor", colored=lambda s, *a, **k: s)
    if "easydict" not in sys.modules:
        _mod("easydict", EasyDict=dict)
    if REFERENCE not in sys.path:
        sys.path.insert(0, REFERENCE)


def load_by_path(name, relpath):
    """Import a reference file that has no package dependencies (numpy/scipy only)."""
    spec = importlib.util.spec_from_file_location(name, f"{REFERENCE}/{relpath}")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
