"""B200 mirror of the reference's DVPQ evaluation script ``eval/eval_dvpq.py`` including its file formats.

    python -m video_latent_diffusion_panoptic_segmentation_b200.eval.eval_dvpq --pan_dir P --depth_dir D \
        --eval_frames k --depth_thres t [--gt_dir video_sequence/val]

Same command line (eval_dvpq.py:11-22; the ground-truth directory, hard-coded there as 'video_sequence/val', is an
option here with that default), same directory conventions (eval_dvpq.py:153-184) and the same printed line
``PQ PQ-things PQ-stuff`` (eval_dvpq.py:205-210):

  gt_dir     ``*gtFine_class*.png`` (category per pixel), the matching ``*gtFine_instance*.png`` (the reader replaces
             'class' by 'instance' in the path, eval_dvpq.py:113-119) and ``*depth*.png``
  pan_dir    ``*cat.png`` / ``*ins.png`` per frame; panoptic id = cat * 2^20 + ins (eval_dvpq.py:105-110)
  depth_dir  one predicted depth PNG per frame (every file of the directory, sorted)

A window of k consecutive frames is concatenated along the width (eval_dvpq.py:108-109,145). With depth_thres > 0 the
predictions whose abs-rel depth error exceeds the threshold become class 19 (eval_dvpq.py:123-143) -- that step and the
three area histograms of ``vpq_eval`` run on the GPU (ldm_depth_mask_pred, ldm_joint_hist); PNG decoding stays on the host.
``write_panoptic_pngs`` is the writer of the prediction format (the reference only reads it).
"""
import argparse
import os

import numpy as np
import torch
from PIL import Image

from .. import ops
from ..ldmseg.evaluations.new_eval import aggregate, vpq_eval

MAX_INS = 2 ** 20


def read_png(path):
    """np.array(Image.open(path)) as the reference does: uint8 for 8-bit, uint16 for 16-bit ('I;16'), int32 for 'I'."""
    return np.array(Image.open(path))


def write_panoptic_pngs(pan, out_dir, stem, max_ins=MAX_INS):
    """pan: int [H, W] panoptic ids (cat * max_ins + ins) -> ``<stem>_cat.png`` (8 bit) and ``<stem>_ins.png``
    (8 bit when every instance id < 256, else 16 bit). Returns the two paths."""
    pan = np.asarray(pan).astype(np.int64)
    cat, ins = pan // max_ins, pan % max_ins
    if cat.min() < 0 or cat.max() > 255:
        raise ValueError(f"write_panoptic_pngs: category range [{cat.min()}, {cat.max()}] does not fit 8 bits")
    if ins.max() > 65535:
        raise ValueError(f"write_panoptic_pngs: instance id {ins.max()} does not fit 16 bits")
    os.makedirs(out_dir, exist_ok=True)
    pc, pi = os.path.join(out_dir, stem + "_cat.png"), os.path.join(out_dir, stem + "_ins.png")
    Image.fromarray(cat.astype(np.uint8)).save(pc)
    Image.fromarray(ins.astype(np.uint8) if ins.max() < 256 else ins.astype(np.uint16)).save(pi)
    return pc, pi


def _sample_bits(a, b):
    if a.dtype != b.dtype:
        raise ValueError(f"depth maps of different sample types: {a.dtype} vs {b.dtype}")
    if a.dtype == np.uint8:
        return 8
    if a.dtype == np.uint16:
        return 16
    if a.dtype in (np.int32, np.int16):
        return 32
    raise ValueError(f"unsupported depth sample type {a.dtype}")


def eval_arrays(pred_cat, pred_ins, gt_cat, gt_ins, depth_pred=None, depth_gt=None, depth_thres=0.0, max_ins=MAX_INS,
                device="cuda"):
    """One window from decoded arrays (lists of k [H, W] maps) -> (iou, tp, fn, fp, abs_rel), eval_dvpq.py:104-150."""
    pc = np.concatenate([np.asarray(x) for x in pred_cat], axis=1).astype(np.int32)
    pi = np.concatenate([np.asarray(x) for x in pred_ins], axis=1).astype(np.int32)
    pred = torch.from_numpy(pc * max_ins + pi).to(device)
    gt = np.concatenate([np.asarray(c).astype(np.int32) * max_ins + np.asarray(i).astype(np.int32)
                         for c, i in zip(gt_cat, gt_ins)], axis=1)
    abs_rel = 0
    if depth_thres > 0:
        dp = np.concatenate([np.asarray(x) for x in depth_pred], axis=1)
        dg = np.concatenate([np.asarray(x) for x in depth_gt], axis=1)
        bits = _sample_bits(dp, dg)
        abs_rel = ops.depth_mask_pred(pred, torch.from_numpy(dp.astype(np.int32)).to(device),
                                      torch.from_numpy(dg.astype(np.int32)).to(device), bits, depth_thres, 19 * max_ins)
    return vpq_eval([pred, torch.from_numpy(gt).to(device)], max_ins=max_ins, device=device) + (abs_rel,)


def eval(element, depth_thres=0.0, device="cuda"):  # noqa: A001 (the reference's name)
    """element = (pred_cat paths, pred_ins paths, gt class paths, depth pred paths, depth gt paths) of one window."""
    pred_cat, pred_ins, gts, depth_preds, depth_gts = element
    gt_cat = [read_png(p) for p in gts]
    gt_ins = [read_png(p.replace("class", "instance")) for p in gts]
    dpl = [read_png(p) for p in depth_preds] if depth_thres > 0 else None
    dgl = [read_png(p) for p in depth_gts] if depth_thres > 0 else None
    return eval_arrays([read_png(p) for p in pred_cat], [read_png(p) for p in pred_ins], gt_cat, gt_ins, dpl, dgl,
                       depth_thres, device=device)


def collect(gt_dir, pred_dir, depth_dir, eval_frames):
    """The sliding windows of eval_dvpq.py:153-184 (sorted names, every run of eval_frames consecutive frames)."""
    gt_names = sorted(os.path.join(gt_dir, n) for n in os.listdir(gt_dir) if "gtFine_class" in n)
    depth_gt_names = sorted(os.path.join(gt_dir, n) for n in os.listdir(gt_dir) if "depth" in n)
    depth_pred_names = sorted(os.path.join(depth_dir, n) for n in os.listdir(depth_dir)) if depth_dir else []
    pred_names = [os.path.join(pred_dir, n) for n in os.listdir(pred_dir)]
    cat_names = sorted(n for n in pred_names if n.endswith("cat.png"))
    ins_names = sorted(n for n in pred_names if n.endswith("ins.png"))
    k = eval_frames
    return [[cat_names[i:i + k], ins_names[i:i + k], gt_names[i:i + k], depth_pred_names[i:i + k],
             depth_gt_names[i:i + k]] for i in range(len(cat_names) - k + 1)]


def run(gt_dir, pred_dir, depth_dir, eval_frames=1, depth_thres=0.0, device="cuda", rank=0, world=1):
    """Evaluate this rank's share of the windows (contiguous ranges; the caller reduces with reduce_rows)."""
    windows = collect(gt_dir, pred_dir, depth_dir, eval_frames)
    per = (len(windows) + world - 1) // world
    return [eval(w, depth_thres, device) for w in windows[rank * per:(rank + 1) * per]]


def main(argv=None):
    ap = argparse.ArgumentParser(description="DVPQ evaluation (mirror of eval/eval_dvpq.py)")
    ap.add_argument("--pan_dir", type=str, default="")
    ap.add_argument("--depth_dir", type=str, default="")
    ap.add_argument("--gt_dir", type=str, default="video_sequence/val")
    ap.add_argument("--eval_frames", type=int, default=1)
    ap.add_argument("--depth_thres", type=float, default=0)
    args = ap.parse_args(argv)
    rows = run(args.gt_dir, args.pan_dir, args.depth_dir, args.eval_frames, args.depth_thres)
    res = aggregate(rows)
    print("{:.1f} {:.1f} {:.1f}".format(res["pq"], res["pq_things"], res["pq_stuff"]))
    return res


if __name__ == "__main__":
    main()
