"""B200 mirror of ``CityscapesPanopticEvaluator`` (ldmseg/evaluations/cityscapes_pap_eval.py:9-249).

The reference labels connected components with scipy on the CPU and then computes an O(#gt x #pred x H x W) table of
boolean-mask IoUs. Here the pixel work of a whole BATCH of images is two C-ABI calls --
``ldm_city_pan_maps`` (one 4-connected component labelling over all thing classes of all predictions and ground truths
at once, scipy numbering per class, and the composition of both panoptic id maps) and ``ldm_joint_hist_batch`` (one
(gt, pred) id histogram per image) -- and ONE device->host copy of the histogram tables, after which every mask
intersection / area is an integer in a table of a few hundred entries, and the reference's greedy matching (:119-174)
runs on those integers in the same order, with the same float64 divisions, so TP / FP / FN / iou_sum are bit-identical.
``panoptic_maps`` keeps the step-by-step form (ldm_ccl_label4 per class, ldm_pan_insert, ldm_id_mask) that the batched
kernels are tested against.
"""
import numpy as np
import torch

from ... import ops

i32 = torch.int32


def _to_dev_i32(a, device):
    if torch.is_tensor(a):
        return a.to(device=device, dtype=i32).contiguous()
    return torch.from_numpy(np.ascontiguousarray(a).astype(np.int32)).to(device)


class CityscapesPanopticEvaluator:
    def __init__(self, thing_ids={11, 12, 13, 14, 15, 16, 17, 18}, ignore_label=0, iou_thresh=0.5, max_ins=1 << 20,
                 device="cuda"):
        self.thing_ids = set(thing_ids)
        self.ignore_label = ignore_label
        self.iou_thresh = iou_thresh
        self.max_ins = max_ins
        self.device = torch.device(device)
        self.reset()

    def reset(self):
        self.TP = self.FP = self.FN = 0
        self.iou_sum = 0.0
        self.TP_per_class, self.FP_per_class, self.FN_per_class, self.iou_sum_per_class = {}, {}, {}, {}

    # ------------------------------------------------------------------ device part
    def panoptic_maps(self, pred_seg, gt_semseg):
        """int32 device maps (pred_pan, gt_pan) of :66-110. pred_seg: -1 = void; gt_semseg: semantic labels."""
        dev = self.device
        pred = _to_dev_i32(pred_seg, dev).clone()
        gt = _to_dev_i32(gt_semseg, dev)
        H, W = pred.shape[-2:]
        pred3, gt3 = pred.view(1, H, W), gt.view(1, H, W)
        ops.id_mask(pred, pred, -1, fill=self.ignore_label)          # pred_seg[pred_seg == -1] = ignore_label
        # which labels are present (np.unique of both maps) comes from one joint histogram
        pv, gv, _ = ops.joint_hist(pred, gt)
        pred_labels, gt_labels = set(pv.tolist()), set(gv.tolist())
        gt_pan = gt.clone()
        for t in sorted(self.thing_ids & gt_labels):                 # :76-84,46-48  sem*max_ins + component
            lab, _ = ops.ccl_label4(gt3, t)
            ops.pan_insert(gt, lab.view(H, W), t, self.max_ins, gt_pan)
        ops.id_mask(gt_pan, gt, self.ignore_label, fill=-1)          # :49,110
        pred_pan = pred.clone()
        ops.id_mask(pred_pan, pred, self.ignore_label, fill=0)       # np.zeros_like + skip ignore label (:91-93)
        for t in sorted(self.thing_ids & pred_labels):               # :96-103
            if t == self.ignore_label:
                continue
            lab, _ = ops.ccl_label4(pred3, t)
            ops.pan_insert(pred, lab.view(H, W), t, self.max_ins, pred_pan)
        ops.id_mask(pred_pan, gt, self.ignore_label, pred, self.ignore_label, fill=-1)  # :108-109
        return pred_pan, gt_pan

    # ------------------------------------------------------------------ host part (tiny integer tables)
    def _cat(self, i):
        return (i // self.max_ins) if i >= self.max_ins else i

    def _thing_slots(self):
        """int8 [2, 256] device table for ldm_city_pan_maps: row 0 prediction (the ignore label is never a thing there,
        :97-98), row 1 ground truth."""
        if getattr(self, "_slots", None) is None:
            things = sorted(t for t in self.thing_ids if 0 <= t < 256)
            if len(things) != len(self.thing_ids) or len(things) > 32:
                raise ValueError("thing ids must lie in [0, 256) and number at most 32")
            tab = np.full((2, 256), -1, np.int8)
            for i, t in enumerate(things):
                tab[1, t] = i
                if t != self.ignore_label:
                    tab[0, t] = i
            self._slots = (torch.from_numpy(tab).to(self.device), len(things))
        return self._slots

    def add_images(self, pred_segs, gt_semsegs):
        """add_image for a batch ([B,H,W] int32 device tensors, or anything add_image takes stacked along dim 0):
        two kernel sequences and one device->host copy for the whole batch."""
        dev = self.device
        pred = _to_dev_i32(pred_segs, dev)
        gt = _to_dev_i32(gt_semsegs, dev)
        if pred.dim() == 2:
            pred, gt = pred[None], gt[None]
        slots, n_things = self._thing_slots()
        pred_pan, gt_pan = ops.city_pan_maps(pred, gt, slots, n_things, self.ignore_label, self.max_ins)
        B, H, W = pred.shape
        for g, p, c in ops.joint_hist_batch(gt_pan, pred_pan, B, H * W, H * W):
            self._match(g, p, c)

    def add_image(self, pred_seg, gt_semseg):
        self.add_images(pred_seg, gt_semseg)

    def add_image_stepwise(self, pred_seg, gt_semseg):
        """The unbatched form (one labelling per thing class): the test reference for add_images."""
        pred_pan, gt_pan = self.panoptic_maps(pred_seg, gt_semseg)
        self._match(*ops.joint_hist(gt_pan, pred_pan))

    def _match(self, g, p, c):
        gt_area, pred_area, joint = {}, {}, {}
        for gi, pi, ci in zip(g.tolist(), p.tolist(), c.tolist()):
            gt_area[gi] = gt_area.get(gi, 0) + ci
            pred_area[pi] = pred_area.get(pi, 0) + ci
            joint[(gi, pi)] = ci
        gt_ids = sorted(k for k in gt_area if k != -1)
        pred_ids = sorted(k for k in pred_area if k != -1)
        matched_pred = set()
        for gid in gt_ids:
            gcat = self._cat(gid)
            if gcat not in self.TP_per_class:
                self.TP_per_class[gcat] = 0
                self.FP_per_class[gcat] = 0
                self.FN_per_class[gcat] = 0
                self.iou_sum_per_class[gcat] = 0.0
            best_iou, best_pid = 0.0, None
            for pid in pred_ids:
                if self._cat(pid) != gcat:
                    continue
                inter = joint.get((gid, pid), 0)
                union = gt_area[gid] + pred_area[pid] - inter
                iou = 0.0 if union == 0 else inter / union
                if iou > best_iou:
                    best_iou, best_pid = iou, pid
            if best_iou >= self.iou_thresh:
                self.TP += 1
                self.iou_sum += best_iou
                matched_pred.add(best_pid)
                self.TP_per_class[gcat] += 1
                self.iou_sum_per_class[gcat] += best_iou
            else:
                self.FN += 1
                self.FN_per_class[gcat] = self.FN_per_class.get(gcat, 0) + 1
        self.FP += len(pred_ids) - len(matched_pred)
        for pid in pred_ids:
            if pid not in matched_pred:
                pcat = self._cat(pid)
                self.FP_per_class[pcat] = self.FP_per_class.get(pcat, 0) + 1

    def stats_tensor(self):
        """[TP, FP, FN] int64 + iou_sum float64 for the cross-rank reduction (exact for the integer part)."""
        return torch.tensor([self.TP, self.FP, self.FN], dtype=torch.int64), torch.tensor([self.iou_sum],
                                                                                           dtype=torch.float64)

    def evaluate(self):
        if self.TP == 0:
            sq = rq = pq = 0.0
        else:
            sq = self.iou_sum / self.TP
            rq = self.TP / (self.TP + 0.5 * (self.FP + self.FN))
            pq = sq * rq
        per_class = {}
        for cat in self.TP_per_class.keys():
            tp, fp = self.TP_per_class.get(cat, 0), self.FP_per_class.get(cat, 0)
            fn, iou_sum = self.FN_per_class.get(cat, 0), self.iou_sum_per_class.get(cat, 0.0)
            if tp == 0:
                cat_sq = cat_rq = cat_pq = 0.0
            else:
                cat_sq = iou_sum / tp
                denom = tp + 0.5 * (fp + fn)
                cat_rq = tp / denom if denom > 0 else 0.0
                cat_pq = cat_sq * cat_rq
            per_class[int(cat)] = {"pq": cat_pq, "sq": cat_sq, "rq": cat_rq, "tp": tp, "fp": fp, "fn": fn}
        sums = {"thing": [0.0, 0.0, 0.0, 0], "stuff": [0.0, 0.0, 0.0, 0]}
        for cat, m in per_class.items():
            s = sums["thing" if cat in self.thing_ids else "stuff"]
            s[0] += m["pq"]; s[1] += m["sq"]; s[2] += m["rq"]; s[3] += 1
        res = {"pq": pq * 100, "sq": sq * 100, "rq": rq * 100, "tp": self.TP, "fp": self.FP, "fn": self.FN,
               "iou_sum": self.iou_sum, "per_class": per_class}
        for k, s in sums.items():
            n = s[3] if s[3] > 0 else 1
            res[f"{k}_pq"], res[f"{k}_sq"], res[f"{k}_rq"] = s[0] / n * 100, s[1] / n * 100, s[2] / n * 100
        return res
