"""ORACLE (test infrastructure only): numpy restatement of the PQ / DVPQ statistics.

Pinned by tests/golden/make_golden.py against the real reference functions imported from /root/reference
(eval/eval_dvpq.py:25-101 ``vpq_eval``, eval/new_eval.py:2-75 (max_ins=64 variant),
ldmseg/evaluations/cityscapes_pap_eval.py:9-249 ``CityscapesPanopticEvaluator``), fixtures in tests/golden/.

  vpq_stats            eval/eval_dvpq.py:25-101
  dvpq_aggregate       eval/eval_dvpq.py:190-210
  depth_mask_pred      eval/eval_dvpq.py:123-145
  dvpq_window          eval/eval_dvpq.py:104-150 (pinned by tests/golden/dvpq_files.json, made by the real eval())
  CityscapesPQOracle   ldmseg/evaluations/cityscapes_pap_eval.py:64-249
"""
import numpy as np
from scipy import ndimage


def _areas(a):
    ids, counts = np.unique(a, return_counts=True)
    return dict(zip(ids.tolist(), counts.tolist()))


def vpq_stats(pred_ids, gt_ids, max_ins=2 ** 20, ign_id=255, offset=2 ** 30, num_cat=20, guard_union=False):
    """Returns (iou, tp, fn, fp) float64[num_cat]. guard_union=True is the eval/new_eval.py variant (iou=0 if union<=0)."""
    iou_c, tp_c = np.zeros(num_cat, np.float64), np.zeros(num_cat, np.float64)
    fn_c, fp_c = np.zeros(num_cat, np.float64), np.zeros(num_cat, np.float64)
    pred_area, gt_area = _areas(pred_ids), _areas(gt_ids)
    joint = _areas(gt_ids.astype(np.int64) * offset + pred_ids.astype(np.int64))
    void = ign_id * max_ins
    ignored_gt = [g for g in gt_area if g // max_ins == ign_id]
    gt_hit, pred_hit = set(), set()
    for key, inter in joint.items():  # ascending key order == np.unique order
        g, p = key // offset, key % offset
        gc, pc = g // max_ins, p // max_ins
        if gc != pc:
            continue
        union = gt_area[g] + pred_area[p] - inter - joint.get(void * offset + p, 0)
        if guard_union and union <= 0:
            iou = 0
        else:
            iou = inter / union
        if iou > 0.5:
            tp_c[gc] += 1
            iou_c[gc] += iou
            gt_hit.add(g)
            pred_hit.add(p)
    for g in gt_area:
        if g in gt_hit or g // max_ins == ign_id:
            continue
        fn_c[g // max_ins] += 1
    for p, area in pred_area.items():
        if p in pred_hit:
            continue
        ign_overlap = sum(joint.get(g * offset + p, 0) for g in ignored_gt)
        if ign_overlap / area > 0.5:
            continue
        fp_c[p // max_ins] += 1
    return iou_c, tp_c, fn_c, fp_c


def dvpq_aggregate(rows, n_classes=19, n_things=8):
    """rows: list of (iou, tp, fn, fp[, abs_rel]) per window, summed in window order (eval_dvpq.py:190-210).
    Returns dict(pq, pq_things, pq_stuff) in percent plus the summed per-class arrays."""
    iou = np.stack([r[0] for r in rows]).sum(axis=0)[:n_classes]
    tp = np.stack([r[1] for r in rows]).sum(axis=0)[:n_classes]
    fn = np.stack([r[2] for r in rows]).sum(axis=0)[:n_classes]
    fp = np.stack([r[3] for r in rows]).sum(axis=0)[:n_classes]
    eps = 1e-10
    sq = iou / (tp + eps)
    rq = tp / (tp + 0.5 * fn + 0.5 * fp + eps)
    pq = sq * rq
    return dict(pq=pq.mean() * 100, pq_things=pq[:n_things].mean() * 100, pq_stuff=pq[n_things:].mean() * 100,
                iou=iou, tp=tp, fn=fn, fp=fp)


def depth_mask_pred(pred, depth_pred, depth_gt, depth_thres, max_ins=2 ** 20):
    """eval_dvpq.py:123-145: predictions whose abs-rel depth error exceeds the threshold become class 19.
    Returns (masked pred copy, abs_rel)."""
    pred = pred.copy()
    if depth_thres <= 0:
        return pred, 0
    valid = depth_gt > 0
    rel = np.abs(depth_pred[valid] - depth_gt[valid]) / depth_gt[valid]
    sub = pred[:, :depth_pred.shape[1]]
    vals = sub[valid]
    vals[rel > depth_thres] = 19 * max_ins
    sub[valid] = vals
    pred[:, :depth_pred.shape[1]] = sub
    return pred, np.mean(rel)


def dvpq_window(pred_cat, pred_ins, gt_cat, gt_ins, depth_pred=None, depth_gt=None, depth_thres=0.0, max_ins=2 ** 20):
    """eval_dvpq.py:104-150 on decoded arrays (lists of k [H, W] maps, the PNG sample types untouched: the uint16
    depth arithmetic wraps exactly as in the reference). Returns (iou, tp, fn, fp, abs_rel)."""
    pc = np.concatenate(pred_cat, axis=1)
    pi = np.concatenate(pred_ins, axis=1)
    pred = pc.astype(np.int32) * max_ins + pi.astype(np.int32)
    gt = np.concatenate([c.astype(np.int32) * max_ins + i.astype(np.int32) for c, i in zip(gt_cat, gt_ins)], axis=1)
    abs_rel = 0
    if depth_thres > 0:
        pred, abs_rel = depth_mask_pred(pred, np.concatenate(depth_pred, axis=1), np.concatenate(depth_gt, axis=1),
                                        depth_thres, max_ins)
    return vpq_stats(pred.astype(np.int64), gt.astype(np.int64), max_ins=max_ins) + (abs_rel,)


class CityscapesPQOracle:
    def __init__(self, thing_ids=(11, 12, 13, 14, 15, 16, 17, 18), ignore_label=0, iou_thresh=0.5, max_ins=1 << 20):
        self.thing_ids, self.ignore_label = set(thing_ids), ignore_label
        self.iou_thresh, self.max_ins = iou_thresh, max_ins
        self.reset()

    def reset(self):
        self.TP = self.FP = self.FN = 0
        self.iou_sum = 0.0
        self.per_class = {}  # cat -> [tp, fp, fn, iou_sum]

    def _cat(self, i):
        return i // self.max_ins if i >= self.max_ins else i

    def panoptic_maps(self, pred_seg, gt_semseg):
        """The two int64 id maps whose overlaps add_image matches (cityscapes_pap_eval.py:71-110)."""
        pred_seg = pred_seg.copy()
        pred_seg[pred_seg == -1] = self.ignore_label
        gt_ins = np.zeros_like(gt_semseg)
        for t in self.thing_ids:  # set iteration order is irrelevant: masks are disjoint
            m = gt_semseg == t
            if m.any():
                lab, n = ndimage.label(m)
                if n > 0:
                    gt_ins[m] = lab[m]
        sem, ins = gt_semseg.astype(np.int64), gt_ins.astype(np.int64)
        gt_pan = np.where(np.isin(sem, list(self.thing_ids)), sem * self.max_ins + ins, sem)
        gt_pan[sem == self.ignore_label] = -1
        pred_pan = np.zeros_like(pred_seg)
        for label in np.unique(pred_seg):
            if label == self.ignore_label:
                continue
            if label in self.thing_ids:
                comp, n = ndimage.label(pred_seg == label)
                for j in range(1, n + 1):
                    pred_pan[comp == j] = label * self.max_ins + j
            else:
                pred_pan[pred_seg == label] = label
        pred_pan[gt_semseg == self.ignore_label] = -1
        pred_pan[pred_seg == self.ignore_label] = -1
        return pred_pan, gt_pan

    def match(self, gt_area, pred_area, joint):
        """Greedy matching from integer areas: gt ids ascending, best same-category pred by IoU (strict >),
        TP if IoU >= thresh (cityscapes_pap_eval.py:119-174). joint: {(gid, pid): pixels}."""
        gids = sorted(g for g in gt_area if g != -1)
        pids = sorted(p for p in pred_area if p != -1)
        matched = set()
        for g in gids:
            gc = self._cat(g)
            st = self.per_class.setdefault(gc, [0, 0, 0, 0.0])
            best, best_p = 0.0, None
            for p in pids:
                if self._cat(p) != gc:
                    continue
                inter = joint.get((g, p), 0)
                union = gt_area[g] + pred_area[p] - inter
                iou = 0.0 if union == 0 else inter / union
                if iou > best:
                    best, best_p = iou, p
            if best >= self.iou_thresh:
                self.TP += 1
                self.iou_sum += best
                matched.add(best_p)
                st[0] += 1
                st[3] += best
            else:
                self.FN += 1
                st[2] += 1
        self.FP += len(pids) - len(matched)
        for p in pids:
            if p not in matched:
                # NB reference quirk: FP for a category never seen in GT is recorded in FP_per_class only, and
                # evaluate() iterates TP_per_class keys, so such categories do not appear in per_class.
                self.per_class.setdefault(("fp_only", self._cat(p)), [0, 0, 0, 0.0])
                key = self._cat(p) if self._cat(p) in self.per_class else ("fp_only", self._cat(p))
                self.per_class[key][1] += 1

    def add_image(self, pred_seg, gt_semseg):
        pred_pan, gt_pan = self.panoptic_maps(pred_seg, gt_semseg)
        gt_area, pred_area = _areas(gt_pan), _areas(pred_pan)
        pairs, counts = np.unique(np.stack([gt_pan.ravel(), pred_pan.ravel()]), axis=1, return_counts=True)
        joint = {(int(g), int(p)): int(c) for (g, p), c in zip(pairs.T, counts)}
        self.match(gt_area, pred_area, joint)

    def evaluate(self):
        if self.TP == 0:
            sq = rq = pq = 0.0
        else:
            sq = self.iou_sum / self.TP
            rq = self.TP / (self.TP + 0.5 * (self.FP + self.FN))
            pq = sq * rq
        per_class = {}
        for cat, (tp, fp, fn, iou_sum) in self.per_class.items():
            if isinstance(cat, tuple):
                continue
            if tp == 0:
                c_sq = c_rq = c_pq = 0.0
            else:
                c_sq = iou_sum / tp
                denom = tp + 0.5 * (fp + fn)
                c_rq = tp / denom if denom > 0 else 0.0
                c_pq = c_sq * c_rq
            per_class[int(cat)] = {"pq": c_pq, "sq": c_sq, "rq": c_rq, "tp": tp, "fp": fp, "fn": fn}
        agg = {"thing": [0.0, 0.0, 0.0, 0], "stuff": [0.0, 0.0, 0.0, 0]}
        for cat, m in per_class.items():
            a = agg["thing" if cat in self.thing_ids else "stuff"]
            a[0] += m["pq"]; a[1] += m["sq"]; a[2] += m["rq"]; a[3] += 1
        out = {"pq": pq * 100, "sq": sq * 100, "rq": rq * 100, "tp": self.TP, "fp": self.FP, "fn": self.FN,
               "iou_sum": self.iou_sum, "per_class": per_class}
        for k, a in agg.items():
            n = a[3] if a[3] > 0 else 1
            out[f"{k}_pq"], out[f"{k}_sq"], out[f"{k}_rq"] = a[0] / n * 100, a[1] / n * 100, a[2] / n * 100
        return out
