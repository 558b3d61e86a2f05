"""CPU: the built library holds sm_100a code only and its SASS contains the Blackwell instructions the design claims
(B200_PROFILING.md: UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG = TMA tensor load / store, LDTM / STTM = tcgen05.ld / st,
UTCBAR = tcgen05.commit). No GPU needed: cuobjdump reads the cubins embedded in libldmseg_b200.so."""
import collections
import re
import shutil
import subprocess

import pytest

from video_latent_diffusion_panoptic_segmentation_b200 import _lib as L

CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"


def _run(*args):
    try:
        return subprocess.run([CUOBJDUMP, *args, L.LIB_PATH], check=True, capture_output=True, text=True).stdout
    except (OSError, subprocess.CalledProcessError) as e:
        pytest.skip(f"cuobjdump unavailable: {e}")


def test_library_is_sm100a_only():
    L.load()
    elfs = re.findall(r"ELF file\s+\d+:\s+(\S+)", _run("-lelf"))
    assert len(elfs) >= 8
    assert all(e.endswith(".sm_100a.cubin") for e in elfs), elfs


def test_sass_contains_tcgen05_tmem_and_tma_instructions():
    L.load()
    sass = _run("-sass")
    per_kernel, current = collections.defaultdict(collections.Counter), None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            current = m.group(1)
            continue
        for op in re.findall(r"\b(UTCHMMA|UTMALDG|UTMASTG|LDTM|STTM|UTCBAR|MUFU\.EX2|MUFU\.TANH)\b", line):
            per_kernel[current][op] += 1

    def kernels_with(op, substr):
        return [k for k, c in per_kernel.items() if substr in k and c[op] > 0]

    # the contraction kernel: tcgen05.mma fed by TMA loads, accumulators read back from TMEM, TMA stores in the epilogue
    assert len(kernels_with("UTCHMMA", "gemm_tc_kernel")) >= 8          # every (pair, epilogue) instantiation
    assert kernels_with("UTMALDG", "gemm_tc_kernel") and kernels_with("LDTM", "gemm_tc_kernel")
    assert kernels_with("UTMASTG", "gemm_tc_kernel")                    # staged bf16 outputs and the q / k head blocks
    # flash attention: S / P / O in TMEM (tcgen05.ld AND tcgen05.st: P goes back into TMEM), exponentials on MUFU
    for name in ("flash_attn40_kernel", "flash_attn_kernel"):
        for op in ("UTCHMMA", "UTMALDG", "LDTM", "STTM", "MUFU.EX2"):
            assert kernels_with(op, name), (name, op)
    # GroupNorm + SiLU through one MUFU op (tanh)
    assert kernels_with("MUFU.TANH", "gn_")
