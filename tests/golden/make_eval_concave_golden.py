"""Golden vectors of the evaluation path with CONCAVE ground-truth quads, from the REFERENCE ITSELF (lifted functions,
float64 Polygon stand-in whose ``is_valid`` / ``intersection`` / ``contains`` follow shapely for simple concave rings:
Detect_OBB.py:148-151, :631-636).

  python tests/golden/make_eval_concave_golden.py        (build container only: needs /root/reference)

eval_concave_golden.json: one image whose labels are ~45 % concave "arrow" quads (one corner of a rotated box pulled
inside past the diagonal), plus what the reference returns for _match_dets_to_gts_pixel, compute_pr_for_class,
evaluate_center_hit and _evaluate_dataset on it.
"""
import io
import json
import os
import sys
from contextlib import redirect_stdout

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from oracle import geometry as G  # noqa: E402
from oracle import lift_reference as LR  # noqa: E402
from make_eval_golden import rbox  # noqa: E402


def make_concave_case(rng, n_obj, n_cls, extent):
    dets, gts = [], []
    n_concave = 0
    for _ in range(n_obj):
        cx, cy = rng.uniform(0, extent, 2)
        w, h, th = rng.uniform(20, 90), rng.uniform(18, 80), rng.uniform(-0.7, 2.3)
        cls = int(rng.integers(0, n_cls))
        q = rbox(cx, cy, w, h, th)
        if rng.random() < 0.45:
            k = int(rng.integers(0, 4))                         # pull corner k towards (and past) the centre
            q[k] = np.array([cx, cy]) + (q[k] - np.array([cx, cy])) * rng.uniform(-0.35, 0.3)
        if rng.random() < 0.3:
            q = q[::-1]
        pts = [(float(x), float(y)) for x, y in q]
        n_concave += G.quad_classify(pts)[0] == 2
        gts.append({"cls": cls, "pts": pts})
        for _ in range(int(rng.integers(0, 3))):
            b = rbox(cx + rng.normal(0, 4), cy + rng.normal(0, 4), w * rng.uniform(0.7, 1.2), h * rng.uniform(0.7, 1.2),
                     th + rng.normal(0, 0.1)).astype(np.float32)
            c = cls if rng.random() < 0.9 else int(rng.integers(0, n_cls))
            dets.append(tuple(float(v) for v in b.reshape(-1)) + (c, float(np.float32(rng.uniform(0.001, 1.0))), 0.0))
    order = rng.permutation(len(dets))
    return [dets[i] for i in order], gts, int(n_concave)


def main():
    ref = LR.load_detect(3)
    rng = np.random.default_rng(777)
    dets, gts, n_concave = make_concave_case(rng, 160, 3, 900.0)
    images = {"concave.png": {"dets": dets, "gts": gts}}
    ref._load_gt_as_pixels = lambda p: [dict(g) for g in images[p]["gts"]]
    ref.all_dets_per_image = {k: list(v["dets"]) for k, v in images.items()}
    names = list(images)
    gold = {"images": images, "n_concave_gt": n_concave, "match": {}, "pr": {}, "dataset": {}}
    for thr in (0.25, 0.5, 0.75):
        gold["match"][repr(thr)] = list(ref._match_dets_to_gts_pixel(dets, gts, iou_thr=thr))
    for cid in range(3):
        d, g = ref.gather_detections_and_gts(ref.all_dets_per_image, names, cid)
        for thr in (0.3, 0.5):
            p, r, ap, tp, fp, fn = ref.compute_pr_for_class(d, g, iou_thr=thr)
            gold["pr"][f"{cid}@{thr}"] = {"precision": np.asarray(p).tolist(), "recall": np.asarray(r).tolist(),
                                          "ap": float(ap), "tp": int(tp), "fp": int(fp), "fn": int(fn)}
    with redirect_stdout(io.StringIO()):
        for thr in (0.25, 0.5):
            gold["dataset"][repr(thr)] = {"center_hit": list(ref.evaluate_center_hit(names, conf_thr=thr)),
                                          "prf": list(ref._evaluate_dataset(names, conf_thr=thr, iou_thr=0.25))}
    # the same run with concave quads declared invalid (the round-1 restatement) must differ, or the vectors pin nothing
    with open(os.path.join(HERE, "eval_concave_golden.json"), "w") as fh:
        json.dump(gold, fh)
    print("concave GT:", n_concave, "of", len(gts), "| match:", gold["match"], "| dataset:", gold["dataset"])


if __name__ == "__main__":
    main()
