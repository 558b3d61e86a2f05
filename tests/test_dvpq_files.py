"""DVPQ file formats and the depth term (reference eval/eval_dvpq.py:104-184).

tests/golden/dvpq_files.json was produced by the REAL reference `eval(element)` on PNGs of the seeded clip of
tests/synth_dvpq.py (tests/golden/make_golden_dvpq_files.py). CPU tests pin the oracle restatement and the host logic
(PNG writer / reader, window listing); the GPU test runs the mirror end to end on files."""
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

from synth_dvpq import dvpq_clip  # noqa: E402

GOLD = json.load(open(os.path.join(HERE, "golden", "dvpq_files.json")))


def _write_clip(clip, root):
    from PIL import Image
    gt_dir, pred_dir, depth_dir = (os.path.join(root, d) for d in ("gt", "pred", "depth"))
    for d in (gt_dir, pred_dir, depth_dir):
        os.makedirs(d)
    for f in range(len(clip["gt_cat"])):
        stem = f"000000_{f:06d}"
        Image.fromarray(clip["gt_cat"][f]).save(os.path.join(gt_dir, stem + "_gtFine_class.png"))
        Image.fromarray(clip["gt_ins"][f]).save(os.path.join(gt_dir, stem + "_gtFine_instance.png"))
        Image.fromarray(clip["depth_gt"][f]).save(os.path.join(gt_dir, stem + "_depth_718.8560180664062.png"))
        Image.fromarray(clip["depth_pred"][f]).save(os.path.join(depth_dir, stem + ".png"))
    return gt_dir, pred_dir, depth_dir


def test_oracle_window_matches_reference_eval():
    from oracle import eval_oracle as EO
    clip = dvpq_clip(GOLD["seed"], GOLD["n_frames"])
    assert GOLD["depth_png_dtype"] == "uint16" == str(clip["depth_gt"][0].dtype)
    for case in GOLD["cases"]:
        k, thres = case["eval_frames"], case["depth_thres"]
        for w, row in enumerate(case["rows"]):
            sl = slice(w, w + k)
            got = EO.dvpq_window(clip["pred_cat"][sl], clip["pred_ins"][sl], clip["gt_cat"][sl], clip["gt_ins"][sl],
                                 clip["depth_pred"][sl], clip["depth_gt"][sl], thres)
            for a, b in zip(got[:4], row[:4]):
                assert np.array_equal(a, np.asarray(b)), (k, thres, w)
            assert got[4] == row[4], (k, thres, w)  # same numpy mean: bit-identical


def test_png_writer_reader_and_window_listing(tmp_path):
    from video_latent_diffusion_panoptic_segmentation_b200.eval import eval_dvpq as E
    clip = dvpq_clip(GOLD["seed"], GOLD["n_frames"])
    gt_dir, pred_dir, depth_dir = _write_clip(clip, str(tmp_path))
    for f in range(GOLD["n_frames"]):
        pan = clip["pred_cat"][f].astype(np.int64) * E.MAX_INS + clip["pred_ins"][f]
        pc, pi = E.write_panoptic_pngs(pan, pred_dir, f"000000_{f:06d}")
        assert np.array_equal(E.read_png(pc), clip["pred_cat"][f]) and np.array_equal(E.read_png(pi), clip["pred_ins"][f])
    # 16-bit instance ids survive too
    big = np.array([[3 * E.MAX_INS + 300, 7 * E.MAX_INS + 65535]], dtype=np.int64)
    pc, pi = E.write_panoptic_pngs(big, str(tmp_path / "big"), "x")
    assert E.read_png(pi).dtype == np.uint16 and E.read_png(pi).tolist() == [[300, 65535]]
    with pytest.raises(ValueError):
        E.write_panoptic_pngs(np.array([[256 * E.MAX_INS]]), str(tmp_path / "bad"), "x")
    for k in (1, 2):
        win = E.collect(gt_dir, pred_dir, depth_dir, k)
        assert len(win) == GOLD["n_frames"] - k + 1
        assert all(len(part) == k for w in win for part in w)
        assert [os.path.basename(p) for p in win[0][0]] == [f"000000_{f:06d}_cat.png" for f in range(k)]
        assert all("gtFine_class" in p for w in win for p in w[2]) and all("depth" in p for w in win for p in w[4])
    assert str(E.read_png(win[0][4][0]).dtype) == "uint16"


@pytest.mark.gpu
def test_eval_dvpq_on_files_matches_reference(tmp_path):
    import torch
    from video_latent_diffusion_panoptic_segmentation_b200.eval import eval_dvpq as E
    clip = dvpq_clip(GOLD["seed"], GOLD["n_frames"])
    gt_dir, pred_dir, depth_dir = _write_clip(clip, str(tmp_path))
    for f in range(GOLD["n_frames"]):
        E.write_panoptic_pngs(clip["pred_cat"][f].astype(np.int64) * E.MAX_INS + clip["pred_ins"][f], pred_dir,
                              f"000000_{f:06d}")
    for case in GOLD["cases"]:
        k, thres = case["eval_frames"], case["depth_thres"]
        rows = E.run(gt_dir, pred_dir, depth_dir, k, thres)
        assert len(rows) == len(case["rows"])
        for got, ref in zip(rows, case["rows"]):
            for a, b in zip(got[:4], ref[:4]):  # TP / FN / FP counts and IoU sums: bit-exact
                assert np.array_equal(a, np.asarray(b)), (k, thres)
            # abs-rel: the reference's np.mean sums pairwise, the kernel by CTA -> agreement to rounding, not bitwise
            assert abs(got[4] - ref[4]) <= 1e-12 * max(1.0, abs(ref[4])), (k, thres, got[4], ref[4])
    torch.cuda.synchronize()
    # the CLI prints the reference's line
    res = E.main(["--gt_dir", gt_dir, "--pan_dir", pred_dir, "--depth_dir", depth_dir, "--eval_frames", "2",
                  "--depth_thres", "0.5"])
    assert set(res) >= {"pq", "pq_things", "pq_stuff", "abs_rel"}
