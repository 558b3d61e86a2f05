"""DVPQ of a clip whose frames are sharded over the ranks of one box, on device-resident id maps.

The reference evaluates DVPQ offline, from PNGs, in an ``mp.Pool`` (eval/eval_dvpq.py:153-210): every window of
``eval_frames`` consecutive frames is one ``eval`` call and the per-window rows are summed. Here the sampler's id maps
stay on the GPUs that produced them:

  frames      rank r owns the contiguous range ``shard_range(n_frames, r, world)`` (as DistributedSampler-style
              sharding of the reference's val set, trainers_ldm_cond.py:246-273, but contiguous so that windows are local);
  halo        a window of k frames that starts in a rank's range reaches into the next shard(s): every rank contributes
              its first k-1 frames to ONE ``all_gather`` (NCCL over NVLink / NVSwitch; 1.9 MB per id map at 384x1248)
              and takes the k-1 frames that follow its range from it;
  windows     a window is evaluated by the rank that owns its first frame: ``pan = cat * max_ins + ins``
              (``ldm_pan_combine``), optional depth masking (``ldm_depth_mask_pred``), the joint id histogram
              (``ldm_joint_hist``) and the reference's matching loops (``vpq_eval``). A histogram does not depend on
              the arrangement of the pixels, so the k frames are stacked, not concatenated along the width;
  statistics  TP / FN / FP per class: ``all_reduce(SUM)`` of int64 [3, 20] (exact). IoU sums and abs-rel: the
              per-window float64 rows are all-gathered and added in WINDOW order by the reference's own aggregation
              (eval_dvpq.py:190-210), so the result is bit-identical to a single process whatever the world size.
"""
import numpy as np
import torch

from ..ldmseg.evaluations.new_eval import aggregate, vpq_eval, vpq_from_hist
from .. import ops

i32 = torch.int32
NUM_CAT = 20


def shard_range(n_frames, rank, world):
    """Contiguous frame range [lo, hi) of `rank` (ceil(n / world) frames per rank, the last shards may be short/empty)."""
    per = (n_frames + world - 1) // world
    lo = min(rank * per, n_frames)
    return lo, min(lo + per, n_frames)


def _dist():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist
    return None


def exchange_halo(maps, halo, n_frames):
    """maps: list of tensors [n_local, H, W] (this rank's frames, same n_local for all). Returns the list extended to
    [n_local + h, H, W] with the h = min(halo, frames after this shard) frames that follow the shard. One all_gather of
    every rank's first `halo` frames (padded); tensors stay on their device (CUDA for NCCL, CPU for gloo)."""
    dist = _dist()
    if halo <= 0 or dist is None:
        return list(maps)
    rank, world = dist.get_rank(), dist.get_world_size()
    n_local = maps[0].shape[0]
    H, W = maps[0].shape[-2:]
    head = torch.zeros((len(maps), halo, H, W), dtype=maps[0].dtype, device=maps[0].device)
    n_head = min(halo, n_local)
    for j, m in enumerate(maps):
        if m.dtype != maps[0].dtype or m.shape != maps[0].shape:
            raise ValueError("exchange_halo: all maps must share dtype and shape")
        head[j, :n_head] = m[:n_head]
    heads = [torch.empty_like(head) for _ in range(world)]
    dist.all_gather(heads, head)
    lo, hi = shard_range(n_frames, rank, world)
    need = min(halo, n_frames - hi)
    parts, r = [[m] for m in maps], rank + 1
    while need > 0 and r < world:
        rlo, rhi = shard_range(n_frames, r, world)
        take = min(need, rhi - rlo, halo)
        if take > 0:
            for j in range(len(maps)):
                parts[j].append(heads[r][j, :take])
            need -= take
        r += 1
    return [torch.cat(p, dim=0) if len(p) > 1 else p[0] for p in parts]


def eval_window_device(pred_cat, pred_ins, gt_cat, gt_ins, depth_pred=None, depth_gt=None, depth_thres=0.0,
                       max_ins=2 ** 20, depth_bits=16):
    """One window (eval_dvpq.py:104-150) from int32 device tensors [k, H, W] -> (iou, tp, fn, fp, abs_rel)."""
    dev = pred_cat.device
    pred = ops.pan_combine(pred_cat.contiguous(), pred_ins.contiguous(), max_ins)
    gt = ops.pan_combine(gt_cat.contiguous(), gt_ins.contiguous(), max_ins)
    abs_rel = 0
    if depth_thres > 0:
        k, H, W = pred.shape
        abs_rel = ops.depth_mask_pred(pred.view(k * H, W), depth_pred.contiguous().view(k * H, W),
                                      depth_gt.contiguous().view(k * H, W), depth_bits, depth_thres, 19 * max_ins)
    return vpq_eval([pred, gt], max_ins=max_ins, device=dev) + (abs_rel,)


def reduce_rows(rows, n_windows, device=None):
    """rows: this rank's per-window (iou[20], tp[20], fn[20], fp[20], abs_rel) tuples, in window order, for the windows
    that start in its shard. Returns the reference's aggregate (eval_dvpq.py:190-210) over ALL windows, identical on
    every rank: counts all-reduced as integers, float rows gathered and summed in window order."""
    if n_windows <= 0:
        raise ValueError("the clip has fewer frames than eval_frames: no window to evaluate")
    dist = _dist()
    if dist is None:
        res = aggregate(rows)
        res["n_windows"] = len(rows)
        return res
    world = dist.get_world_size()
    dev = torch.device(device) if (device is not None and dist.get_backend() == "nccl") else torch.device("cpu")
    width = 4 * NUM_CAT + 1
    n_local = torch.tensor([len(rows)], dtype=torch.int64, device=dev)
    ns = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(ns, n_local)                                        # windows per rank (also sizes the row buffer)
    ns = [int(t.item()) for t in ns]
    cap = max(1, max(ns))
    local_h = np.zeros((cap, width), dtype=np.float64)   # assembled on the host, one transfer each
    counts_h = np.zeros((3, NUM_CAT), dtype=np.int64)
    for i, r in enumerate(rows):
        local_h[i] = np.concatenate([np.asarray(r[0], np.float64), np.asarray(r[1], np.float64),
                                     np.asarray(r[2], np.float64), np.asarray(r[3], np.float64),
                                     np.asarray([r[4]], np.float64)])
        for j in range(3):
            counts_h[j] += np.asarray(r[1 + j]).astype(np.int64)
    local = torch.from_numpy(local_h).to(dev)
    counts = torch.from_numpy(counts_h).to(dev)
    dist.all_reduce(counts, op=dist.ReduceOp.SUM)                       # exact integer statistics
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)                                    # float rows, added in window order below
    all_rows = []
    for r in range(world):                                              # contiguous shards: rank order == window order
        g = gathered[r].cpu().numpy()
        for i in range(ns[r]):
            all_rows.append((g[i, 0:20], g[i, 20:40], g[i, 40:60], g[i, 60:80], g[i, 80]))
    if len(all_rows) != n_windows:
        raise RuntimeError(f"reduce_rows: gathered {len(all_rows)} windows, expected {n_windows}")
    res = aggregate(all_rows)
    c = counts.cpu().numpy().astype(np.float64)
    for j, name in enumerate(("tp", "fn", "fp")):
        if not np.array_equal(res[name], c[j, :len(res[name])]):
            raise RuntimeError(f"reduce_rows: all-reduced {name} counts differ from the gathered rows")
    res["n_windows"] = n_windows
    return res


def dvpq_clip_sharded(pred_cat, pred_ins, gt_cat, gt_ins, n_frames, eval_frames=1, depth_pred=None, depth_gt=None,
                      depth_thres=0.0, max_ins=2 ** 20, depth_bits=16, eval_fn=None):
    """DVPQ of a clip of `n_frames` frames. Every argument holds THIS rank's frames ``shard_range(n_frames, rank,
    world)`` as an int32 tensor [n_local, H, W]. Returns the aggregate dict (pq, pq_things, pq_stuff, iou, tp, fn, fp,
    abs_rel, n_windows), identical on all ranks. `eval_fn(pred_cat, pred_ins, gt_cat, gt_ins, depth_pred, depth_gt)`
    evaluates one window given [k, H, W] tensors (default: the GPU kernels)."""
    dist = _dist()
    rank, world = (dist.get_rank(), dist.get_world_size()) if dist is not None else (0, 1)
    k = int(eval_frames)
    lo, hi = shard_range(n_frames, rank, world)
    if pred_cat.shape[0] != hi - lo:
        raise ValueError(f"rank {rank} holds {pred_cat.shape[0]} frames, its shard of {n_frames} is [{lo}, {hi})")
    maps = [pred_cat, pred_ins, gt_cat, gt_ins]
    with_depth = depth_thres > 0
    if with_depth:
        maps += [depth_pred, depth_gt]
    ext = exchange_halo(maps, k - 1, n_frames)
    n_win = max(0, min(hi, n_frames - k + 1) - lo)
    if eval_fn is None and not with_depth and n_win > 0:
        # all windows of this rank at once: pan ids of every frame, ONE launch for the n_win joint histograms (window
        # i = frames [i, i + k) of the frame-major maps), ONE device->host copy, then the reference's matching loops
        H, W = ext[0].shape[-2:]
        pred = ops.pan_combine(ext[0].contiguous(), ext[1].contiguous(), max_ins)
        gt = ops.pan_combine(ext[2].contiguous(), ext[3].contiguous(), max_ins)
        rows = [vpq_from_hist(g, p, c, max_ins=max_ins) + (0,)
                for g, p, c in ops.joint_hist_batch(gt, pred, n_win, k * H * W, H * W)]
        return reduce_rows(rows, max(0, n_frames - k + 1), device=pred_cat.device)
    if eval_fn is None:
        def eval_fn(pc, pi, gc, gi, dp, dg):
            return eval_window_device(pc, pi, gc, gi, dp, dg, depth_thres, max_ins, depth_bits)
    rows = []
    for i in range(lo, hi):
        if i + k > n_frames:
            break
        a, b = i - lo, i - lo + k
        rows.append(eval_fn(ext[0][a:b], ext[1][a:b], ext[2][a:b], ext[3][a:b],
                            ext[4][a:b] if with_depth else None, ext[5][a:b] if with_depth else None))
    return reduce_rows(rows, max(0, n_frames - k + 1), device=pred_cat.device)
