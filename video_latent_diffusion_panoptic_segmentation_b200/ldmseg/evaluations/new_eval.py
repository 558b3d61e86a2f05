"""B200 mirror of the DVPQ / VPQ statistics: ``vpq_eval`` (eval/eval_dvpq.py:25-101 == ldmseg/evaluations/new_eval.py:14-90;
eval/new_eval.py:2-75 is the max_ins=64 + union guard variant), the window / depth-threshold logic of ``eval``
(eval_dvpq.py:104-150) and the final aggregation (:190-210).

The three ``np.unique(..., return_counts=True)`` passes over the id maps (pred areas, gt areas, gt*2^30+pred joint
areas) collapse into ONE GPU joint-histogram kernel (ldm_joint_hist): the marginals are sums over its table. The
matching loops then run on the host over that small table, in the reference's iteration order (ascending joint id),
with the same float64 arithmetic, so the four per-class arrays are bit-identical.

Semantics note: with int32 id maps numpy >= 2 overflows int32 in ``_ign_id * offset + pred_id`` (eval_dvpq.py:60),
silently disabling the ignored-overlap FP filter; this mirror implements the non-overflowing (numpy 1.x / intended)
semantics, which is what tests/golden pins.
"""
import numpy as np
import torch

from ... import ops

i32 = torch.int32


def _dev_i32(a, device):
    if torch.is_tensor(a):
        return a.to(device=device, dtype=i32).contiguous()
    return torch.from_numpy(np.ascontiguousarray(a).astype(np.int32)).to(device)


def vpq_eval(element, max_ins=2 ** 20, ign_id=255, num_cat=20, guard_union=False, device="cuda"):
    """element = [pred_ids, gt_ids] (numpy or torch, any int dtype; values must fit int32) ->
    (iou_per_class, tp_per_class, fn_per_class, fp_per_class), float64[num_cat]."""
    pred_ids, gt_ids = element
    gv, pv, cv = ops.joint_hist(_dev_i32(gt_ids, device), _dev_i32(pred_ids, device))
    return vpq_from_hist(gv, pv, cv, max_ins=max_ins, ign_id=ign_id, num_cat=num_cat, guard_union=guard_union)


def vpq_from_hist(gv, pv, cv, max_ins=2 ** 20, ign_id=255, num_cat=20, guard_union=False):
    """The matching loops of vpq_eval (eval_dvpq.py:50-101) on a joint (gt, pred) id histogram sorted by (gt, pred)."""
    iou_c, tp_c = np.zeros(num_cat, np.float64), np.zeros(num_cat, np.float64)
    fn_c, fp_c = np.zeros(num_cat, np.float64), np.zeros(num_cat, np.float64)
    gt_area, pred_area, joint = {}, {}, {}
    for g, p, c in zip(gv.tolist(), pv.tolist(), cv.tolist()):  # ascending (gt, pred) == ascending gt*offset+pred
        gt_area[g] = gt_area.get(g, 0) + c
        pred_area[p] = pred_area.get(p, 0) + c
        joint[(g, p)] = c
    void_id = ign_id * max_ins
    ign_ids = [g for g in gt_area if g // max_ins == ign_id]
    gt_matched, pred_matched = set(), set()
    for (g, p), inter in joint.items():
        gcat, pcat = g // max_ins, p // max_ins
        if gcat != pcat:
            continue
        union = gt_area[g] + pred_area[p] - inter - joint.get((void_id, p), 0)
        iou = 0 if (guard_union and union <= 0) else inter / union
        if iou > 0.5:
            tp_c[gcat] += 1
            iou_c[gcat] += iou
            gt_matched.add(g)
            pred_matched.add(p)
    for g in sorted(gt_area):
        if g in gt_matched or g // max_ins == ign_id:
            continue
        fn_c[g // max_ins] += 1
    for p in sorted(pred_area):
        if p in pred_matched:
            continue
        if sum(joint.get((g, p), 0) for g in ign_ids) / pred_area[p] > 0.5:
            continue
        fp_c[p // max_ins] += 1
    return iou_c, tp_c, fn_c, fp_c


def eval_window(pred_cat, pred_ins, gt_cat, gt_ins, depth_pred=None, depth_gt=None, depth_thres=0.0,
                max_ins=2 ** 20, device="cuda"):
    """One sliding window of k frames (eval_dvpq.py:104-150): the frames' maps are concatenated along the width,
    pan = cat*max_ins + ins; predictions whose abs-rel depth error exceeds depth_thres become class 19.
    Each argument is a list of k [H,W] arrays. Returns (iou, tp, fn, fp, abs_rel)."""
    pc = np.concatenate([np.asarray(x) for x in pred_cat], axis=1).astype(np.int32)
    pi = np.concatenate([np.asarray(x) for x in pred_ins], axis=1).astype(np.int32)
    pred = pc * max_ins + pi
    gt = np.concatenate([np.asarray(c).astype(np.int32) * max_ins + np.asarray(i).astype(np.int32)
                         for c, i in zip(gt_cat, gt_ins)], axis=1)
    abs_rel = 0
    if depth_thres > 0:
        dp = np.concatenate([np.asarray(x) for x in depth_pred], axis=1)
        dg = np.concatenate([np.asarray(x) for x in depth_gt], axis=1)
        valid = dg > 0
        rel = np.abs(dp[valid] - dg[valid]) / dg[valid]
        abs_rel = np.mean(rel)
        sub = pred[:, :dp.shape[1]]
        vals = sub[valid]
        vals[rel > depth_thres] = 19 * max_ins
        sub[valid] = vals
        pred[:, :dp.shape[1]] = sub
    return vpq_eval([pred, gt], max_ins=max_ins, device=device) + (abs_rel,)


def aggregate(results, n_classes=19, n_things=8):
    """eval_dvpq.py:190-210: sum the per-window rows in window order, PQ = mean(sq*rq) over the first 19 classes."""
    iou = np.stack([r[0] for r in results]).sum(axis=0)[:n_classes]
    tp = np.stack([r[1] for r in results]).sum(axis=0)[:n_classes]
    fn = np.stack([r[2] for r in results]).sum(axis=0)[:n_classes]
    fp = np.stack([r[3] for r in results]).sum(axis=0)[:n_classes]
    abs_rel = np.stack([r[4] for r in results]).mean(axis=0) if len(results[0]) > 4 else 0
    eps = 1e-10
    sq = iou / (tp + eps)
    rq = tp / (tp + 0.5 * fn + 0.5 * fp + eps)
    pq = sq * rq
    return {"pq": pq.mean() * 100, "pq_things": pq[:n_things].mean() * 100, "pq_stuff": pq[n_things:].mean() * 100,
            "iou": iou, "tp": tp, "fn": fn, "fp": fp, "abs_rel": abs_rel}
