"""CPU restatement of the reference's evaluation path.  TEST INFRASTRUCTURE ONLY.

Follows (file:line of the reference repository):
  * greedy detection -> GT matching per image        Detect_OBB.py:456-480
  * precision / recall / F1                          Detect_OBB.py:482-486
  * AP from a PR curve                               Detect_OBB.py:489-499
  * per-class PR over all images, score-sorted       Detect_OBB.py:512-565
  * mAP over IoU thresholds                          Detect_OBB.py:574-607
  * centre-hit metric                                Detect_OBB.py:609-648
The polygon arithmetic (shapely in the reference) is the float64 restatement of oracle/geometry.py.
Pinned against the reference's own functions through tests/golden/eval_golden.json
(tests/golden/make_eval_golden.py runs the lifted reference; tests/test_oracle_eval.py compares).
Pure-Python loops: small cases only.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np

from .geometry import Point, Polygon, quad_iou


def flat(pts) -> List[float]:
    return [c for pt in pts for c in pt]


def match_image(dets: Sequence[tuple], gts: Sequence[dict], iou_thr: float = 0.5):
    """(TP, FP, FN) and the GT index taken by each detection (-1 = none).  Detect_OBB.py:456-480."""
    used = [False] * len(gts)
    taken = []
    for det in dets:
        best_iou, best_j = 0.0, -1
        for j, g in enumerate(gts):
            if used[j] or int(det[8]) != g["cls"]:
                continue
            iou = quad_iou(det[:8], flat(g["pts"]))
            if iou > best_iou:
                best_iou, best_j = iou, j
        if best_iou >= iou_thr and best_j >= 0:
            used[best_j] = True
            taken.append(best_j)
        else:
            taken.append(-1)
    tp = sum(1 for t in taken if t >= 0)
    return (tp, len(dets) - tp, used.count(False)), taken


def prec_rec_f1(tp, fp, fn):
    P = tp / (tp + fp + 1e-9)
    R = tp / (tp + fn + 1e-9)
    return P, R, 2 * P * R / (P + R + 1e-9)


def ap_from_pr(recall, precision) -> float:
    mrec = np.concatenate(([0.0], recall, [1.0]))
    mpre = np.concatenate(([0.0], precision, [0.0]))
    for i in range(mpre.size - 2, -1, -1):
        mpre[i] = max(mpre[i], mpre[i + 1])
    idx = np.where(mrec[1:] != mrec[:-1])[0]
    return float(np.sum((mrec[idx + 1] - mrec[idx]) * mpre[idx + 1]))


def pr_for_class(dets: List[dict], gts: Dict[str, list], iou_thr: float = 0.5):
    """Detect_OBB.py:512-565.  dets: {"image_id", "score", "bbox"}; gts: image -> list of 8-float quads."""
    npos = sum(len(v) for v in gts.values())
    if npos == 0:
        return np.array([0.0]), np.array([0.0]), 0.0, 0.0, 0.0, 0, 0, 0
    ds = sorted(dets, key=lambda d: d["score"], reverse=True)
    if not ds:
        return np.array([0.0]), np.array([0.0]), 0.0, 0.0, 0.0, 0, 0, npos
    tp, fp = np.zeros(len(ds)), np.zeros(len(ds))
    matched = {img: [False] * len(v) for img, v in gts.items()}
    for i, d in enumerate(ds):
        best_iou, best_j = 0.0, -1
        for j, g in enumerate(gts.get(d["image_id"], [])):
            if matched[d["image_id"]][j]:
                continue
            iou = quad_iou(d["bbox"], g)
            if iou > best_iou:
                best_iou, best_j = iou, j
        if best_iou >= iou_thr and best_j >= 0:
            tp[i] = 1
            matched[d["image_id"]][best_j] = True
        else:
            fp[i] = 1
    tpc, fpc = np.cumsum(tp), np.cumsum(fp)
    recall = tpc / (npos + 1e-9)
    precision = tpc / (tpc + fpc + 1e-9)
    return precision, recall, ap_from_pr(recall, precision), int(tpc[-1]), int(fpc[-1]), npos - int(tpc[-1])


def gather(dets_source: Dict[str, list], gts_by_image: Dict[str, list], images, cls_id: int, min_score: float = 0.001):
    """Detect_OBB.py:501-510 with the label files replaced by ``gts_by_image``."""
    dets, gts = [], {}
    for img in images:
        for d in dets_source.get(img, []):
            if int(d[8]) == cls_id and d[9] >= min_score:
                dets.append({"image_id": img, "score": float(d[9]), "bbox": d[:8]})
        gts[img] = [flat(g["pts"]) for g in gts_by_image[img] if g["cls"] == cls_id]
    return dets, gts


def evaluate_map(dets_source, gts_by_image, images, iou_list=None, min_score: float = 0.001):
    """Detect_OBB.py:574-607."""
    if iou_list is None:
        iou_list = [0.5] + [round(0.5 + 0.05 * i, 2) for i in range(1, 10)]
    class_ids = sorted({int(g["cls"]) for img in images for g in gts_by_image[img]})
    per_iou = {}
    for iou in iou_list:
        aps = [pr_for_class(*gather(dets_source, gts_by_image, images, cid, min_score), iou_thr=iou)[2] for cid in class_ids]
        per_iou[iou] = float(np.mean(aps)) if aps else 0.0
    return {"mAP@0.5": per_iou.get(0.5, 0.0),
            "mAP@[0.5:0.95]": float(np.mean([per_iou[i] for i in iou_list])) if iou_list else 0.0, "per_iou": per_iou}


def center_hit(dets_source, gts_by_image, images, conf_thr: float = 0.5):
    """(TP, FP, FN).  Detect_OBB.py:609-648."""
    tp = fp = fn = 0
    for img in images:
        gts = gts_by_image[img]
        used = [False] * len(gts)
        for d in (d for d in dets_source.get(img, []) if d[9] >= conf_thr):
            cx = (d[0] + d[2] + d[4] + d[6]) / 4.0
            cy = (d[1] + d[3] + d[5] + d[7]) / 4.0
            hit = False
            for j, g in enumerate(gts):
                if used[j] or g["cls"] != int(d[8]):
                    continue
                poly = Polygon(g["pts"])
                if poly.is_valid and poly.contains(Point(cx, cy)):
                    tp += 1
                    used[j] = True
                    hit = True
                    break
            if not hit:
                fp += 1
        fn += used.count(False)
    return tp, fp, fn
