"""Golden vectors of the evaluation path (Detect_OBB.py:456-648) from the REFERENCE ITSELF: the lifted
functions run on seeded synthetic detections / ground truths (float64 Polygon stand-in for shapely).

  python tests/golden/make_eval_golden.py        (build container only: needs /root/reference)

eval_golden.json: per image the detections (11-tuples) and GTs ({cls, pts}), and what the reference
returns for _match_dets_to_gts_pixel, compute_pr_for_class, evaluate_map (default and "soft" lists),
evaluate_center_hit and _evaluate_dataset.
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

from oracle import lift_reference as LR  # noqa: E402


def rbox(cx, cy, w, h, th):
    c, s = np.cos(th), np.sin(th)
    v1, v2 = np.array([w / 2 * c, w / 2 * s]), np.array([-h / 2 * s, h / 2 * c])
    ctr = np.array([cx, cy])
    return np.stack([ctr + v1 + v2, ctr + v1 - v2, ctr - v1 - v2, ctr - v1 + v2])


def make_case(rng, n_obj, n_cls, extent):
    dets, gts = [], []
    for _ in range(n_obj):
        cx, cy = rng.uniform(0, extent, 2)
        w, h, th = rng.uniform(14, 90), rng.uniform(12, 80), rng.uniform(-0.7, 2.3)
        cls = int(rng.integers(0, n_cls))
        r = rng.random()
        if r < 0.8:                                   # labelled object: a general quad near the true box
            q = rbox(cx, cy, w, h, th) + rng.normal(0, 1.5, (4, 2))
            if rng.random() < 0.04:
                q = q[[0, 2, 1, 3]]                   # bow-tie: invalid polygon
            if rng.random() < 0.03:
                q[:] = q[0]                           # collapsed: zero area
            if rng.random() < 0.3:
                q = q[::-1]                           # clockwise ring
            gts.append({"cls": cls, "pts": [(float(x), float(y)) for x, y in q]})
            if rng.random() < 0.1:                    # duplicate label (equal IoU ties)
                gts.append({"cls": cls, "pts": [(float(x), float(y)) for x, y in q]})
        n_det = int(rng.integers(0, 3)) if r < 0.8 else 1
        for _ in range(n_det):                        # detections: jittered rectangles, sometimes the wrong class
            b = rbox(cx + rng.normal(0, 3), cy + rng.normal(0, 3), w * rng.uniform(0.8, 1.2), h * rng.uniform(0.8, 1.2),
                     th + rng.normal(0, 0.08)).astype(np.float32)
            c = cls if rng.random() < 0.9 else int(rng.integers(0, n_cls))
            conf = float(np.float32(rng.uniform(0.001, 1.0)))
            if rng.random() < 0.08:
                conf = 0.5                            # score ties: stable order decides
            dets.append(tuple(float(v) for v in b.reshape(-1)) + (c, conf, 0.0))
    order = rng.permutation(len(dets))
    return [dets[i] for i in order], gts


def main():
    ref = LR.load_detect(3)
    rng = np.random.default_rng(4242)
    images = {}
    for name, (n_obj, n_cls, extent) in {"a.png": (140, 4, 1200.0), "b.png": (90, 4, 700.0), "c.png": (60, 3, 400.0),
                                         "empty_gt.png": (0, 1, 10.0)}.items():
        dets, gts = make_case(rng, n_obj, n_cls, extent)
        images[name] = {"dets": dets, "gts": gts}
    images["empty_gt.png"]["dets"] = make_case(rng, 10, 2, 300.0)[0]
    ref._load_gt_as_pixels = lambda p: [dict(g) for g in images[p]["gts"]]
    ref.all_dets_per_image = {k: list(v["dets"]) for k, v in images.items()}
    all_images = list(images.keys())
    gold = {"images": images, "match": {}, "pr": {}, "dataset": {}}
    for name, rec in images.items():
        for thr in (0.25, 0.5, 0.75):
            gold["match"][f"{name}@{thr}"] = list(ref._match_dets_to_gts_pixel(rec["dets"], rec["gts"], iou_thr=thr))
    for cid in range(4):
        dets, gts = ref.gather_detections_and_gts(ref.all_dets_per_image, all_images, cid)
        for thr in (0.3, 0.5, 0.85):
            p, r, ap, tp, fp, fn = ref.compute_pr_for_class(dets, gts, iou_thr=thr)
            gold["pr"][f"{cid}@{thr}"] = {"precision": np.asarray(p).tolist(), "recall": np.asarray(r).tolist(),
                                          "ap": float(ap), "tp": int(tp), "fp": int(fp), "fn": int(fn)}
    m = ref.evaluate_map(ref.all_dets_per_image, all_images)
    gold["map_default"] = {"mAP@0.5": m["mAP@0.5"], "mAP@[0.5:0.95]": m["mAP@[0.5:0.95]"],
                           "per_iou": {repr(k): v for k, v in m["per_iou"].items()}}
    soft = [0.30, 0.40, 0.50, 0.60, 0.70]
    m = ref.evaluate_map(ref.all_dets_per_image, all_images, iou_list=soft)
    gold["map_soft"] = {repr(k): v for k, v in m["per_iou"].items()}
    buf = io.StringIO()
    with redirect_stdout(buf):
        for thr in (0.25, 0.5):
            gold["dataset"][repr(thr)] = {"center_hit": list(ref.evaluate_center_hit(all_images, conf_thr=thr)),
                                          "prf": list(ref._evaluate_dataset(all_images, conf_thr=thr, iou_thr=0.25))}
    gold["center_hit_prints"] = buf.getvalue().splitlines()
    with open(os.path.join(HERE, "eval_golden.json"), "w") as fh:
        json.dump(gold, fh)
    print("eval golden written:", {k: (len(v) if hasattr(v, "__len__") else v) for k, v in gold.items()})
    print(gold["map_default"]["mAP@0.5"], gold["dataset"])


if __name__ == "__main__":
    main()
