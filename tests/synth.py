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


def edge_cases_vpq():
    """name -> (pred, gt) int64 panoptic maps (id = cat * 2**20 + ins) for vpq_eval edge cases."""
    M = 2 ** 20
    rng = np.random.default_rng(31)
    out = {}
    gt, _, _ = synth_panoptic(rng, 40, 56, n_seeds=12)
    pred = gt.copy()
    pred[pred // M == 255] = 2 * M  # (a prediction with category 255 makes the reference index out of bounds)
    out["identical"] = (pred, gt)                                        # every segment a TP with IoU 1 (void excluded)
    out["pred_one_wrong_segment"] = (np.full_like(gt, 17 * M + 3), gt)   # one FP, every gt segment a FN
    out["gt_all_void"] = (pred.copy(), np.full_like(gt, 255 * M))        # predictions lie on ignored ground truth only
    out["single_pixel"] = (np.array([[5 * M + 1]], np.int64), np.array([[5 * M + 9]], np.int64))
    # more than 1 024 distinct (gt, pred) pairs: 4x4 blocks with their own instance ids, prediction shifted by (1, 2)
    yy, xx = np.mgrid[0:128, 0:192]
    blk = (yy // 4) * 48 + xx // 4
    g2 = (blk % 19) * M + blk
    p2 = np.roll(g2, (1, 2), axis=(0, 1))
    out["many_instances"] = (p2.astype(np.int64), g2.astype(np.int64))
    # IoU exactly 0.5 is NOT a match (strict >): gt 2x4 block, prediction covers half of it and nothing else
    g3 = np.full((4, 8), 2 * M, np.int64)
    g3[:2, :4] = 6 * M + 1
    p3 = np.full((4, 8), 2 * M, np.int64)
    p3[:2, :2] = 6 * M + 4
    out["iou_exactly_half"] = (p3, g3)
    return out


def edge_cases_city():
    """name -> (pred_seg, gt_semseg) int64 maps for CityscapesPanopticEvaluator.add_image edge cases."""
    rng = np.random.default_rng(32)
    _, cat, _ = synth_panoptic(rng, 48, 64, n_seeds=20, n_cls=19)
    gt = cat.astype(np.int64)
    gt[gt == 255] = 0
    out = {}
    out["pred_all_void"] = (np.full_like(gt, -1), gt)
    out["gt_all_ignore"] = (gt.copy(), np.zeros_like(gt))
    out["identical"] = (gt.copy(), gt)
    # checkerboard of a thing class: every pixel its own 4-connected component (512 components per map)
    yy, xx = np.mgrid[0:32, 0:32]
    chk = np.where((yy + xx) % 2 == 0, 11, 3).astype(np.int64)
    out["checkerboard_components"] = (chk.copy(), chk)
    # a thing segment split in two in the prediction (one gt component vs two pred components), stuff untouched
    g5 = np.full((16, 24), 5, np.int64)
    g5[4:12, 4:20] = 13
    p5 = g5.copy()
    p5[4:12, 11:13] = 5
    out["split_thing"] = (p5, g5)
    return out
