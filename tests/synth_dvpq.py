"""Seeded synthetic DVPS-style clip (category / instance / depth maps per frame) shared by
tests/golden/make_golden_dvpq_files.py (which writes it as PNGs and runs the REAL eval/eval_dvpq.py on them) and the
tests (which re-create the same arrays from the seed)."""
import numpy as np

from synth import synth_panoptic


def dvpq_clip(seed, n_frames=3, H=48, W=64):
    """Returns per-frame lists: gt_cat, gt_ins (uint8), pred_cat, pred_ins (uint8), depth_gt, depth_pred (uint16).
    No void class in the ground truth (so the int32-overflow quirk of eval_dvpq.py:60 under numpy >= 2 cannot matter);
    depth_gt has holes (0 = no measurement); a tenth of the predicted depths is badly off, some below the ground truth
    (which the reference's uint16 arithmetic turns into a wrapped, huge error -- reproduced on purpose)."""
    rng = np.random.default_rng(seed)
    out = {k: [] for k in ("gt_cat", "gt_ins", "pred_cat", "pred_ins", "depth_gt", "depth_pred")}
    for f in range(n_frames):
        _, cat, ins = synth_panoptic(rng, H, W, n_seeds=14, n_cls=19, n_ins=6, void_frac=0.0)
        cat, ins = np.roll(cat, f, axis=1), np.roll(ins, f, axis=1)  # segments drift one pixel per frame
        pc, pi = np.roll(cat, (1, 1), axis=(0, 1)).copy(), np.roll(ins, (1, 1), axis=(0, 1)).copy()
        flip = rng.random((H, W)) < 0.04
        pc[flip], pi[flip] = 4, 9
        dg = rng.integers(500, 20000, (H, W)).astype(np.uint16)
        dg[rng.random((H, W)) < 0.15] = 0
        dp = (dg.astype(np.float64) * (1.0 + 0.05 * rng.standard_normal((H, W)))).clip(1, 65535)
        bad = rng.random((H, W)) < 0.10
        dp[bad] *= rng.choice([0.4, 1.9], size=int(bad.sum()))
        out["gt_cat"].append(cat.astype(np.uint8)); out["gt_ins"].append(ins.astype(np.uint8))
        out["pred_cat"].append(pc.astype(np.uint8)); out["pred_ins"].append(pi.astype(np.uint8))
        out["depth_gt"].append(dg); out["depth_pred"].append(dp.clip(0, 65535).astype(np.uint16))
    return out
