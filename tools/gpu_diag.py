"""Run every per-kernel parity check in its own subprocess (a trapped kernel poisons only its own CUDA context)
and write one JSON line per check to gpurun_out/diag.jsonl.

    python tools/gpu_diag.py [name-substring ...]
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CHILD = r"""
import sys, json, traceback
sys.path.insert(0, {root!r}); sys.path.insert(0, {root!r} + "/tests")
import torch
import gpu_kernel_checks as K
name = sys.argv[1]
try:
    res = K.CHECKS[name]()
    torch.cuda.synchronize()
    print("RESULT " + json.dumps({{"check": name, "ok": True, "info": res}}, default=str))
except Exception as e:
    print("RESULT " + json.dumps({{"check": name, "ok": False, "error": (str(e) or repr(e))[:1500]}}))
"""


def main():
    import gpu_kernel_checks as K  # noqa: imports torch; fine in the parent (no CUDA init)
    names = list(K.CHECKS)
    if len(sys.argv) > 1:
        names = [n for n in names if any(s in n for s in sys.argv[1:])]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    out_path = os.path.join(ROOT, "gpurun_out", "diag.jsonl")
    n_ok = 0
    with open(out_path, "a") as f:
        for name in names:
            t0 = time.time()
            try:
                r = subprocess.run([sys.executable, "-c", CHILD.format(root=ROOT), name], capture_output=True,
                                   text=True, timeout=180)
                line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
                if line:
                    rec = json.loads(line[-1][7:])
                else:
                    rec = {"check": name, "ok": False, "error": "no result", "rc": r.returncode,
                           "stderr": r.stderr[-1500:], "stdout": r.stdout[-500:]}
            except subprocess.TimeoutExpired:
                rec = {"check": name, "ok": False, "error": "timeout"}
            rec["sec"] = round(time.time() - t0, 2)
            n_ok += bool(rec.get("ok"))
            f.write(json.dumps(rec) + "\n")
            f.flush()
            print(json.dumps(rec)[:600], flush=True)
    print(f"DIAG {n_ok}/{len(names)} ok")


if __name__ == "__main__":
    main()
