"""Synthetic stand-ins for the datasets of the reference (ldmseg/data/{kitti,cityscapes}.py read files that do not exist
here; SURVEY section 8d defines the synthetic inputs): image latents of a drifting clip, a ground truth that is
correlated with a prediction so that TP, FP and FN all occur, the (category, instance) split the DVPQ statistics need,
and an order-independent digest of id maps for comparing runs at different world sizes. Plain torch / numpy: this is
input generation and book-keeping outside the timed path, not part of the library."""
import hashlib

import numpy as np
import torch

from ..models.unet_init import TRAINED_LIKE


def trained_like_rgb_latents(n_frames, h, w, seed=1234, first_frame=0, amplitude=TRAINED_LIKE["rgb_amplitude"],
                             regions=TRAINED_LIKE["regions"], drift=TRAINED_LIKE["drift"]):
    """Image latents [n_frames, 4, h, w] f32 of a synthetic clip: `regions` Voronoi cells, each with one constant
    4-vector ~ N(0, 1) * amplitude; the cell centres drift by at most `drift` latent pixels per frame (4 image pixels
    at drift 0.5) and bounce off the frame border, so consecutive frames overlap however long the clip is.
    Frame i depends only on (seed, first_frame + i): any sharding of the clip over ranks sees the same frames."""
    rng = np.random.default_rng(seed)
    cy, cx = rng.uniform(0, h, regions), rng.uniform(0, w, regions)
    vy, vx = rng.uniform(-drift, drift, regions), rng.uniform(-drift, drift, regions)
    vec = rng.standard_normal((regions, 4)).astype(np.float32) * amplitude
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)

    def bounce(p, size):  # triangle wave: reflect at 0 and size
        p = np.mod(p, 2.0 * size)
        return np.where(p > size, 2.0 * size - p, p)

    out = np.empty((n_frames, 4, h, w), np.float32)
    for i in range(n_frames):
        f = first_frame + i
        py, px = bounce(cy + f * vy, h), bounce(cx + f * vx, w)
        d = (yy[None] - py[:, None, None]) ** 2 + (xx[None] - px[:, None, None]) ** 2
        out[i] = vec[d.argmin(0)].transpose(2, 0, 1)
    return torch.from_numpy(out)


def block_majority_tiles(ids, block=8, void=-1):
    """[B, H, W] integer ids (labels >= 0, `void` = -1) -> [B, ceil(H/block), ceil(W/block)] int64: for every
    block x block tile the value that covers most of its pixels (ties go to the smaller value, void included).
    Coarse-graining is neutral: the tile grid knows nothing about where the prediction's boundaries are."""
    B, H, W = ids.shape
    hb, wb = (H + block - 1) // block, (W + block - 1) // block
    lab = ids.long() - void                                     # void -> 0, label l -> l + 1
    n = int(lab.max().item()) + 1
    ty = torch.arange(H, device=ids.device) // block
    tx = torch.arange(W, device=ids.device) // block
    tile = (ty[:, None] * wb + tx[None, :]).expand(B, H, W)
    key = (torch.arange(B, device=ids.device)[:, None, None] * (hb * wb) + tile) * n + lab
    hist = torch.bincount(key.reshape(-1), minlength=B * hb * wb * n).view(B, hb * wb, n)
    return (hist.argmax(dim=2) + void).view(B, hb, wb)          # first maximum = smallest value


def block_majority_labels(ids, block=8, void=-1):
    """block_majority_tiles expanded back to [B, H, W] (same dtype as `ids`)."""
    B, H, W = ids.shape
    t = block_majority_tiles(ids, block, void)
    return t.repeat_interleave(block, 1).repeat_interleave(block, 2)[:, :H, :W].to(ids.dtype).contiguous()


def teacher_ground_truth(ids, block=8, min_area=4096, relabel_every=4, relabel_offset=128,
                         thing_ids=(11, 12, 13, 14, 15, 16, 17, 18)):
    """Synthetic ground-truth labels [B, H, W] (same dtype / device as `ids`; 0 = ignore as in the Cityscapes evaluator,
    cityscapes_pap_eval.py:108-110) derived from a teacher prediction `ids` (-1 = void, labels >= 0):
      * block majority on a `block`-pixel grid (boundaries move by up to block / 2 pixels: matched IoUs land well
        inside (0.5, 1), not at 1);
      * labels that cover fewer than `min_area` pixels of a frame become ignore, and so do the 4-connected components
        of the thing classes (which the evaluator matches one by one, cityscapes_pap_eval.py:76-84) below that size:
        an annotation does not contain segments near the size at which a 0.5 % id difference decides a match;
      * every label with ``label % relabel_every == relabel_every - 1`` is renamed ``label + relabel_offset``: a false
        negative for the ground truth and a false positive for the prediction.
    Label 0 of the prediction is void for the evaluator (pred == 0 is its ignore label too), hence ignore here. The
    filters work on the tile map (48 x 156 tiles for a 384 x 1248 frame) on the host: input generation, not the path."""
    from scipy import ndimage
    B, H, W = ids.shape
    tiles = block_majority_tiles(ids, block).cpu().numpy()       # [B, hb, wb] int64
    px = block * block
    out = np.zeros_like(tiles)
    for b in range(B):
        t = tiles[b]
        keep = np.zeros_like(t, dtype=bool)
        for lab in np.unique(t):
            if lab <= 0:
                continue
            m = t == lab
            if int(m.sum()) * px < min_area:
                continue
            if int(lab) in thing_ids:
                cc, n = ndimage.label(m)                         # 4-connectivity: the evaluator's components
                sizes = np.bincount(cc.ravel(), minlength=n + 1) * px
                m = m & (sizes[cc] >= min_area)
            keep |= m
        g = np.where(keep, t, 0)
        if relabel_every > 0:
            g = np.where((g > 0) & (g % relabel_every == relabel_every - 1), g + relabel_offset, g)
        out[b] = g
    gt = torch.from_numpy(out).to(ids.device)
    return gt.repeat_interleave(block, 1).repeat_interleave(block, 2)[:, :H, :W].to(ids.dtype).contiguous()


def split_cat_ins(ids, n_cat=19, void_cat=19, void=-1, ignore=None, ignore_cat=255):
    """The (category, instance) pair the DVPQ statistics work on (eval/eval_dvpq.py: pan = cat * 2**20 + ins) for the
    class-agnostic labels of the LDMSeg head: cat = label % n_cat, ins = label // n_cat. `void` pixels go to
    (void_cat, 0) -- class 19 is outside the 19 evaluated classes, as the depth-masked pixels of eval_dvpq.py:141 --
    and, for a ground truth, `ignore` pixels to (255, 0), the ignored region of vpq_eval."""
    is_void = ids == void
    cat = torch.where(is_void, torch.full_like(ids, void_cat), ids % n_cat)
    ins = torch.where(is_void, torch.zeros_like(ids), ids // n_cat)
    if ignore is not None:
        ign = ids == ignore
        cat = torch.where(ign, torch.full_like(ids, ignore_cat), cat)
        ins = torch.where(ign, torch.zeros_like(ids), ins)
    return cat, ins


def ids_digest(ids):
    """sha256 over the bytes of int32 id maps [n, H, W] in frame order (hex, first 16 digits): equal digests at N = 1 and
    N > 1 for the same global frames mean identical ids."""
    a = np.ascontiguousarray(ids.detach().to("cpu", torch.int32).numpy())
    return hashlib.sha256(a.tobytes()).hexdigest()[:16]
