"""Build libldmseg_b200.so (hand-written sm_100a CUDA + the C ABI of include/ldmseg_b200.h) in-tree with nvcc.

    python -m video_latent_diffusion_panoptic_segmentation_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libldmseg_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v",
]
# LDM_BUILD_DIAG=1: compile the diagnostic environment switches in (csrc/host_util.h: A/B timing by tools/gpu/*.sh). The
# product build reads no environment variable. The two flavours keep separate object directories.
DIAG = os.environ.get("LDM_BUILD_DIAG", "0") not in ("", "0")
if DIAG:
    FLAGS.append("-DLDM_DIAG")
    OBJ = os.path.join(HERE, "build_diag")


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps_mtime():
    m = 0.0
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for f in os.listdir(root):
            if f.endswith((".cuh", ".h")):
                m = max(m, os.path.getmtime(os.path.join(root, f)))
    return m


def _compile(src, force, verbose):
    obj = os.path.join(OBJ, src[:-3] + ".o")
    spath = os.path.join(CSRC, src)
    if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(spath), _deps_mtime()):
        return obj, ""
    cmd = [NVCC, *FLAGS, "-c", spath, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
    log = r.stdout + r.stderr
    with open(obj + ".ptxas.log", "w") as f:
        f.write(log)
    if verbose:
        print(log)
    return obj, log


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    srcs = sources()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(lambda s: _compile(s, force, verbose), srcs))
    objs = [o for o, _ in results]
    stamp = os.path.join(OBJ, ".linked")
    if force or not os.path.exists(LIB) or not os.path.exists(stamp) or \
            any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs) or os.path.getmtime(stamp) < os.path.getmtime(LIB):
        cmd = [NVCC, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
        for other in (os.path.join(HERE, "build", ".linked"), os.path.join(HERE, "build_diag", ".linked")):
            if os.path.exists(other):
                os.remove(other)   # the library now holds this flavour's objects
        open(stamp, "w").close()
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
