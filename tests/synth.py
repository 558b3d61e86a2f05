"""Seeded synthetic inputs shared by tests/golden/make_golden.py (which runs the real reference on them) and the
tests (which re-create the same inputs from the seeds stored in tests/golden/golden.json)."""
import numpy as np


def synth_panoptic(rng, H, W, n_seeds=24, n_cls=19, n_ins=32, void_frac=0.05, max_ins=2 ** 20):
    """Voronoi-cell panoptic map (SURVEY.md section 8d synthetic GT): id = class*max_ins + instance, 255 = void class."""
    ys, xs = rng.integers(0, H, n_seeds), rng.integers(0, W, n_seeds)
    cls, ins = rng.integers(0, n_cls, n_seeds), rng.integers(0, n_ins, n_seeds)
    yy, xx = np.mgrid[0:H, 0:W]
    d = (yy[None] - ys[:, None, None]) ** 2 + (xx[None] - xs[:, None, None]) ** 2
    owner = d.argmin(0)
    cat = cls[owner].astype(np.int32)
    inst = ins[owner].astype(np.int32)
    void = rng.random((H, W)) < void_frac
    cat[void] = 255
    inst[void] = 0
    return cat.astype(np.int64) * max_ins + inst, cat, inst


def vpq_case(seed, H, W):
    """(pred, gt) int64 panoptic maps in which TP, FN and FP all occur."""
    rng = np.random.default_rng(seed)
    gt, _, _ = synth_panoptic(rng, H, W)
    pred = np.roll(gt.copy(), (2, 3), axis=(0, 1))
    flip = rng.random((H, W)) < 0.03
    pred[flip] = 5 * 2 ** 20 + 7
    pred[pred // 2 ** 20 == 255] = 3 * 2 ** 20 + 1
    return pred, gt


def city_case(seed, H=64, W=96):
    """(pred_seg, gt_semseg) int64 maps for CityscapesPanopticEvaluator.add_image (-1 = void prediction)."""
    rng = np.random.default_rng(seed)
    _, cat, _ = synth_panoptic(rng, H, W, n_seeds=30, n_cls=19)
    gt_sem = cat.astype(np.int64)
    gt_sem[gt_sem == 255] = 0
    pred = np.roll(gt_sem, (1, 2), axis=(0, 1)).copy()
    pred[rng.random(pred.shape) < 0.02] = 13
    pred[rng.random(pred.shape) < 0.05] = -1
    return pred, gt_sem
