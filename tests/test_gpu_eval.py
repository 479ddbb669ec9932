"""GPU parity of the evaluation path (f2): the mirror functions of oriented_object_detection_b200.evaluate
against the reference's own results (tests/golden/eval_golden.json) and against the oracle on a larger
seeded case."""
import io
import json
import os
from contextlib import redirect_stdout

import numpy as np
import pytest

from oracle import evaluation as E

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "eval_golden.json")


@pytest.fixture()
def world(cuda_dev, monkeypatch):
    from oriented_object_detection_b200 import detect, evaluate
    with open(GOLD) as fh:
        g = json.load(fh)
    for rec in g["images"].values():
        rec["dets"] = [tuple(d) for d in rec["dets"]]
        for gt in rec["gts"]:
            gt["pts"] = [tuple(p) for p in gt["pts"]]
    monkeypatch.setattr(evaluate, "_load_gt_as_pixels", lambda p: [dict(x) for x in g["images"][p]["gts"]])
    monkeypatch.setattr(detect, "all_dets_per_image", {k: list(v["dets"]) for k, v in g["images"].items()})
    if hasattr(detect, "all_dets_per_image_map"):
        monkeypatch.delattr(detect, "all_dets_per_image_map")
    return g, detect, evaluate


def test_match_counts_equal_reference(world):
    g, detect, ev = world
    for key, want in g["match"].items():
        name, thr = key.split("@")
        rec = g["images"][name]
        assert list(detect._match_dets_to_gts_pixel(rec["dets"], rec["gts"], iou_thr=float(thr))) == want
    assert detect._match_dets_to_gts_pixel([], g["images"]["a.png"]["gts"]) == (0, 0, len(g["images"]["a.png"]["gts"]))
    assert detect._match_dets_to_gts_pixel(g["images"]["a.png"]["dets"], []) == (0, len(g["images"]["a.png"]["dets"]), 0)


def test_pr_curves_map_and_center_hit_equal_reference(world):
    g, detect, ev = world
    images = list(g["images"].keys())
    for key, want in g["pr"].items():
        cid, thr = key.split("@")
        dets, gts = detect.gather_detections_and_gts(detect.all_dets_per_image, images, int(cid))
        p, r, ap, tp, fp, fn = detect.compute_pr_for_class(dets, gts, iou_thr=float(thr))
        assert (tp, fp, fn) == (want["tp"], want["fp"], want["fn"])
        assert np.array_equal(p, np.asarray(want["precision"])) and np.array_equal(r, np.asarray(want["recall"]))
        assert ap == want["ap"]
    m = detect.evaluate_map(detect.all_dets_per_image, images)
    assert m["mAP@0.5"] == g["map_default"]["mAP@0.5"] and m["mAP@[0.5:0.95]"] == g["map_default"]["mAP@[0.5:0.95]"]
    assert {repr(k): v for k, v in m["per_iou"].items()} == g["map_default"]["per_iou"]
    soft = detect.evaluate_map(detect.all_dets_per_image, images, iou_list=[0.30, 0.40, 0.50, 0.60, 0.70])
    assert {repr(k): v for k, v in soft["per_iou"].items()} == g["map_soft"]
    buf = io.StringIO()
    with redirect_stdout(buf):
        for thr, want in g["dataset"].items():
            assert list(detect.evaluate_center_hit(images, conf_thr=float(thr))) == want["center_hit"]
            assert list(detect._evaluate_dataset(images, conf_thr=float(thr), iou_thr=0.25)) == want["prf"]
    assert buf.getvalue().splitlines() == g["center_hit_prints"]


def test_larger_random_case_against_oracle(cuda_dev):
    """~2.5k detections x ~1.5k general-quad GTs in 6 segments, 10 thresholds in one launch."""
    from oriented_object_detection_b200 import evaluate as ev
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from make_eval_golden import make_case
    rng = np.random.default_rng(99)
    det_rows, det_off, gt_rows, gt_off, cases = [], [0], [], [0], []
    for _ in range(6):
        dets, gts = make_case(rng, 300, 1, 1500.0)
        dets = sorted(dets, key=lambda d: d[9], reverse=True)
        cases.append((dets, gts))
        det_rows += [d[:8] for d in dets]; det_off.append(len(det_rows))
        gt_rows += [E.flat(x["pts"]) for x in gts]; gt_off.append(len(gt_rows))
    thr = [0.3, 0.5, 0.55, 0.6, 0.65, 0.7, 0.75, 0.8, 0.85, 0.9]
    m = ev.match_segments(det_rows, gt_rows, det_off, gt_off, thr)
    for s, (dets, gts) in enumerate(cases):
        for t, th in enumerate(thr[:4]):
            want = E.match_image(dets, gts, th)[1]
            assert m[t, det_off[s]:det_off[s + 1]].tolist() == want
    cls = [0] * len(det_rows)
    hit = ev.center_hit_segments(det_rows, cls, gt_rows, [0] * len(gt_rows), det_off, gt_off)
    for s, (dets, gts) in enumerate(cases):
        tp, fp, fn = E.center_hit({"i": dets}, {"i": gts}, ["i"], conf_thr=0.0)
        assert int((hit[det_off[s]:det_off[s + 1]] >= 0).sum()) == tp


def test_concave_ground_truth_equal_reference(cuda_dev, monkeypatch):
    """Labels that are concave simple quads (valid for shapely): IoU matching, PR curves and centre hits equal the lifted
    reference's results (tests/golden/eval_concave_golden.json)."""
    from oriented_object_detection_b200 import detect, evaluate
    with open(os.path.join(os.path.dirname(GOLD), "eval_concave_golden.json")) as fh:
        g = json.load(fh)
    rec = g["images"]["concave.png"]
    rec["dets"] = [tuple(d) for d in rec["dets"]]
    for gt in rec["gts"]:
        gt["pts"] = [tuple(p) for p in gt["pts"]]
    monkeypatch.setattr(evaluate, "_load_gt_as_pixels", lambda p: [dict(x) for x in rec["gts"]])
    monkeypatch.setattr(detect, "all_dets_per_image", {"concave.png": list(rec["dets"])})
    if hasattr(detect, "all_dets_per_image_map"):
        monkeypatch.delattr(detect, "all_dets_per_image_map")
    for thr, want in g["match"].items():
        assert list(detect._match_dets_to_gts_pixel(rec["dets"], rec["gts"], iou_thr=float(thr))) == want
    for key, want in g["pr"].items():
        cid, thr = key.split("@")
        dets, gts = detect.gather_detections_and_gts(detect.all_dets_per_image, ["concave.png"], int(cid))
        p, r, ap, tp, fp, fn = detect.compute_pr_for_class(dets, gts, iou_thr=float(thr))
        assert (tp, fp, fn) == (want["tp"], want["fp"], want["fn"]) and ap == want["ap"]
    with redirect_stdout(io.StringIO()):
        for thr, want in g["dataset"].items():
            assert list(detect.evaluate_center_hit(["concave.png"], conf_thr=float(thr))) == want["center_hit"]
            assert list(detect._evaluate_dataset(["concave.png"], conf_thr=float(thr), iou_thr=0.25)) == want["prf"]


def test_concave_quads_through_the_public_iou_and_merge(cuda_dev):
    """compute_polygon_iou / rotated_iou_pairs / merge_detections on concave simple quads against the oracle."""
    import torch
    from oracle import geometry as G
    from oriented_object_detection_b200 import detect, ops
    rng = np.random.default_rng(5)
    A, B = [], []
    while len(A) < 300:
        c = rng.uniform(100, 5000, 2)
        qs = []
        for ctr in (c, c + rng.normal(0, 15, 2)):
            ang = np.sort(rng.uniform(0, 2 * np.pi, 4)); r = rng.uniform(10, 60, 4)
            qs.append((ctr + np.stack([r * np.cos(ang), r * np.sin(ang)], 1)).ravel())
        kinds = [G.quad_classify([tuple(v) for v in q.reshape(4, 2)])[0] for q in qs]
        if 2 in kinds and 0 not in kinds:
            A.append(qs[0]); B.append(qs[1])
    A, B = np.array(A), np.array(B)
    want = np.array([G.quad_iou(a, b) for a, b in zip(A, B)])
    assert (want > 0.05).sum() > 50
    got = ops.rotated_iou_pairs(torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()).cpu().numpy()
    assert np.abs(got - want).max() < 1e-6                       # float64 piecewise clip, stored as float32
    for a, b, w in list(zip(A, B, want))[:20]:
        assert abs(detect.compute_polygon_iou(list(a), list(b)) - w) < 1e-12
    dets = [tuple(float(v) for v in q) + (0, float(np.float32(rng.uniform(0.3, 1))), 0.0) for q in np.concatenate([A, B])]
    ref = G.merge_detections(list(dets), 0.4)
    assert detect.merge_detections(list(dets), 0.4) == ref and len(ref) < len(dets)
