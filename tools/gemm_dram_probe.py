"""Probe of the short-K pointwise GEMM (59904 x 320 x 320): rotating buffers so that nothing is L2 resident."""
import os, sys, json, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from video_latent_diffusion_panoptic_segmentation_b200 import ops, _lib as L
L.lib()
M, C, NS = 59904, 320, 6
bf16, f32 = torch.bfloat16, torch.float32
rn = lambda *s: torch.randn(*s, device="cuda").to(bf16)
A = [rn(M, C) for _ in range(NS)]; R = [rn(M, C) for _ in range(NS)]; O = [rn(M, C) for _ in range(NS)]
w, bias = rn(C, C) * 0.05, torch.randn(C, device="cuda")
def run(res, iters=24):
    f = lambda i: ops.gemm(A[i % NS], w, O[i % NS], bias=bias, residual=R[i % NS] if res else None)
    f(0); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters): f(i)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / iters
    nbytes = M * C * 2 * (3 if res else 2)
    return round(us, 1), round(nbytes / us / 1e3)
# plain copy kernels of the same byte counts for reference
def copy_ref(res, iters=24):
    f = (lambda i: torch.add(A[i % NS], R[i % NS], out=O[i % NS])) if res else (lambda i: O[i % NS].copy_(A[i % NS]))
    f(0); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters): f(i)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / iters
    return round(us, 1), round(M * C * 2 * (3 if res else 2) / us / 1e3)
print(json.dumps({"debug": os.environ.get("LDM_GEMM_DEBUG", "none"), "gemm_nores(us,GB/s)": run(False), "gemm_res": run(True),
                  "torch_copy": copy_ref(False), "torch_add": copy_ref(True)}))
