"""Golden fixture for the file-based DVPQ evaluation: writes a seeded synthetic clip as PNGs in the directory layout of
eval/eval_dvpq.py:153-184 and runs the REAL reference `eval(element)` (eval_dvpq.py:104-150) on every window, for
window lengths 1 and 2 and depth thresholds 0 (off), 0.1 and 0.5. Authoring container only (needs /root/reference).

    python tests/golden/make_golden_dvpq_files.py      -> tests/golden/dvpq_files.json
"""
import json
import os
import sys
import tempfile

import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import ref_stubs  # noqa: E402
from synth_dvpq import dvpq_clip  # noqa: E402

SEED, N_FRAMES = 31, 3


def write_clip(clip, root):
    gt_dir, pred_dir, depth_dir = (os.path.join(root, d) for d in ("gt", "pred", "depth"))
    for d in (gt_dir, pred_dir, depth_dir):
        os.makedirs(d)
    for f in range(len(clip["gt_cat"])):
        stem = f"000000_{f:06d}"
        Image.fromarray(clip["gt_cat"][f]).save(os.path.join(gt_dir, stem + "_gtFine_class.png"))
        Image.fromarray(clip["gt_ins"][f]).save(os.path.join(gt_dir, stem + "_gtFine_instance.png"))
        Image.fromarray(clip["depth_gt"][f]).save(os.path.join(gt_dir, stem + "_depth_718.8560180664062.png"))
        Image.fromarray(clip["pred_cat"][f]).save(os.path.join(pred_dir, stem + "_cat.png"))
        Image.fromarray(clip["pred_ins"][f]).save(os.path.join(pred_dir, stem + "_ins.png"))
        Image.fromarray(clip["depth_pred"][f]).save(os.path.join(depth_dir, stem + ".png"))
    return gt_dir, pred_dir, depth_dir


def windows(gt_dir, pred_dir, depth_dir, k):  # the listing logic of eval_dvpq.py:153-184, verbatim in spirit
    gt = sorted(os.path.join(gt_dir, n) for n in os.listdir(gt_dir) if "gtFine_class" in n)
    dgt = sorted(os.path.join(gt_dir, n) for n in os.listdir(gt_dir) if "depth" in n)
    dpr = sorted(os.path.join(depth_dir, n) for n in os.listdir(depth_dir))
    pr = [os.path.join(pred_dir, n) for n in os.listdir(pred_dir)]
    cat, ins = sorted(n for n in pr if n.endswith("cat.png")), sorted(n for n in pr if n.endswith("ins.png"))
    return [[cat[i:i + k], ins[i:i + k], gt[i:i + k], dpr[i:i + k], dgt[i:i + k]] for i in range(len(cat) - k + 1)]


def main():
    sys.argv = [sys.argv[0]]  # eval/eval_dvpq.py parses argv at import
    dv = ref_stubs.load_by_path("ref_eval_dvpq_files", "eval/eval_dvpq.py")
    clip = dvpq_clip(SEED, N_FRAMES)
    out = {"seed": SEED, "n_frames": N_FRAMES, "cases": []}
    with tempfile.TemporaryDirectory() as root:
        dirs = write_clip(clip, root)
        # the PNG round trip must preserve the sample types the reference's arithmetic depends on
        probe = np.array(Image.open(os.path.join(dirs[2], sorted(os.listdir(dirs[2]))[0])))
        out["depth_png_dtype"] = str(probe.dtype)
        for k in (1, 2):
            for thres in (0.0, 0.1, 0.5):
                dv.depth_thres = thres  # module-level global read by eval()
                rows = [dv.eval(w) for w in windows(*dirs, k)]
                out["cases"].append({"eval_frames": k, "depth_thres": thres,
                                     "rows": [[np.asarray(x, dtype=np.float64).tolist() if i < 4 else float(x)
                                               for i, x in enumerate(r)] for r in rows]})
    with open(os.path.join(HERE, "dvpq_files.json"), "w") as f:
        json.dump(out, f)
    print("wrote dvpq_files.json:", out["depth_png_dtype"], [(c["eval_frames"], c["depth_thres"], len(c["rows"])) for c in out["cases"]])


if __name__ == "__main__":
    main()
