"""Evaluation path of the reference (Detect_OBB.py:425-743) on the GPU: detection x ground-truth
matching, PR / AP / mAP, centre-hit metric, class-wise report.

Same function names, arguments and return values as the reference.  What changes is where the
work is done: the float64 rotated IoU of every (detection, GT) candidate pair and the sequential
greedy matching run in ``libgeomap_b200.so`` (``csrc/eval.cu``); for ``evaluate_map`` the IoU
values of a class are computed ONCE and the matching is replayed for all IoU thresholds in one
launch (the reference recomputes every IoU for each of its 10 thresholds, Detect_OBB.py:584-598).
Curve arithmetic (cumulative sums, AP integration) is the reference's numpy arithmetic on
arrays of a few thousand elements and stays on the host.  There is no CPU fallback for the
matching: without the library / an sm_100 GPU the calls raise.

Config globals (``MAP_MIN_SCORE``, ``CLASS_NAMES``, ``all_dets_per_image`` ...) live in
``detect`` like in the reference's single script and are read at call time.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib as L
from . import ops


def _cfg():
    from . import detect
    return detect


# ----------------------------------------------------------------------------- device matching primitives

def _dev_f64(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64).reshape(-1, 8)).to(dev)


def _dev_i32(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).to(dev)


def _dev_i64(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.int64)).to(dev)


def match_segments(det_boxes, gt_boxes, det_off, gt_off, thresholds: Sequence[float],
                   det_cls=None, gt_cls=None) -> np.ndarray:
    """Greedy matching of detections to GTs inside independent segments, for several IoU thresholds.

    Segment s holds detections ``det_off[s]:det_off[s+1]`` (already in processing order) and GTs
    ``gt_off[s]:gt_off[s+1]``.  Rule (Detect_OBB.py:466-478 and :536-552): a detection takes the unused
    GT of strictly largest IoU > 0 (the first on ties; same class only when classes are given) and is
    matched iff that IoU >= threshold.  Returns int32 [len(thresholds), n_det]: the segment-local GT
    index taken, or -1.
    """
    ops._require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device())
    det_off = np.asarray(det_off, dtype=np.int64)
    gt_off = np.asarray(gt_off, dtype=np.int64)
    ns = len(det_off) - 1
    n_det, n_gt = int(det_off[-1]), int(gt_off[-1])
    nt = len(thresholds)
    if n_det == 0 or nt == 0:
        return np.full((nt, n_det), -1, dtype=np.int32)
    mat_off = np.zeros(ns + 1, dtype=np.int64)
    np.cumsum((det_off[1:] - det_off[:-1]) * (gt_off[1:] - gt_off[:-1]), out=mat_off[1:])
    n_pairs = int(mat_off[-1])
    d_det = _dev_f64(det_boxes, dev)
    d_gt = _dev_f64(gt_boxes, dev) if n_gt else torch.empty((0, 8), dtype=torch.float64, device=dev)
    d_dc = _dev_i32(det_cls, dev) if det_cls is not None else None
    d_gc = _dev_i32(gt_cls, dev) if gt_cls is not None else None
    d_do, d_go, d_mo = _dev_i64(det_off, dev), _dev_i64(gt_off, dev), _dev_i64(mat_off, dev)
    iou = torch.empty(max(n_pairs, 1), dtype=torch.float64, device=dev)
    st = ops._stream()
    L.check(L.lib.gm_eval_iou_segments(ops._ptr(d_det), ops._ptr(d_dc), ops._ptr(d_gt), ops._ptr(d_gc), ops._ptr(d_do),
                                       ops._ptr(d_go), ops._ptr(d_mo), ns, n_pairs, ops._ptr(iou), st),
            "gm_eval_iou_segments")
    thr = torch.tensor([float(t) for t in thresholds], dtype=torch.float64, device=dev)
    used = torch.empty(max(nt * n_gt, 1), dtype=torch.uint8, device=dev)
    match = torch.full((nt, n_det), -1, dtype=torch.int32, device=dev)
    L.check(L.lib.gm_eval_match_greedy(ops._ptr(iou), ops._ptr(d_do), ops._ptr(d_go), ops._ptr(d_mo), ns, n_det, n_gt,
                                       ops._ptr(thr), nt, ops._ptr(used), ops._ptr(match), st),
            "gm_eval_match_greedy")
    return match.cpu().numpy()


def center_hit_segments(det_boxes, det_cls, gt_boxes, gt_cls, det_off, gt_off) -> np.ndarray:
    """Per detection: segment-local index of the first unused same-class valid GT polygon that strictly
    contains the detection's centre, or -1 (Detect_OBB.py:623-641)."""
    ops._require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device())
    det_off = np.asarray(det_off, dtype=np.int64)
    gt_off = np.asarray(gt_off, dtype=np.int64)
    n_det, n_gt = int(det_off[-1]), int(gt_off[-1])
    if n_det == 0:
        return np.zeros(0, dtype=np.int32)
    if n_gt == 0:
        return np.full(n_det, -1, dtype=np.int32)
    match = torch.full((n_det,), -1, dtype=torch.int32, device=dev)
    used = torch.empty(n_gt, dtype=torch.uint8, device=dev)
    d = [_dev_f64(det_boxes, dev), _dev_i32(det_cls, dev), _dev_f64(gt_boxes, dev), _dev_i32(gt_cls, dev),
         _dev_i64(det_off, dev), _dev_i64(gt_off, dev)]
    L.check(L.lib.gm_eval_center_hit(*[ops._ptr(t) for t in d], len(det_off) - 1, ops._ptr(used), ops._ptr(match),
                                     ops._stream()), "gm_eval_center_hit")
    return match.cpu().numpy()


# ----------------------------------------------------------------------------- label files (Detect_OBB.py:425-454)

def _label_path_for_image(image_path):
    """<image dir>/<stem>.txt, else <image dir>/Labels/<stem>.txt, else None."""
    stem = os.path.splitext(os.path.basename(image_path))[0] + ".txt"
    here = os.path.dirname(image_path)
    for cand in (os.path.join(here, stem), os.path.join(here, "Labels", stem)):
        if os.path.exists(cand):
            return cand
    return None


def _load_gt_as_pixels(image_path):
    """YOLO-OBB label rows (cls + 8 normalised coordinates) scaled to pixels of the image."""
    import cv2
    gts = []
    lp = _label_path_for_image(image_path)
    if lp is None or not os.path.exists(lp):
        return gts
    img = cv2.imread(image_path)
    if img is None:
        return gts
    h, w = img.shape[:2]
    with open(lp, "r") as fh:
        for line in fh:
            parts = line.strip().split()
            if len(parts) != 9:
                continue
            vals = [float(v) for v in parts[1:]]
            gts.append({"cls": int(parts[0]), "pts": [(vals[i] * w, vals[i + 1] * h) for i in range(0, 8, 2)]})
    return gts


def _flat(pts):
    return [c for pt in pts for c in pt]


# ----------------------------------------------------------------------------- matching metrics

def _match_dets_to_gts_pixel(dets, gts, iou_thr=0.5):
    """Detect_OBB.py:456-480 -> (TP, FP, FN) of one image."""
    if not dets:
        return 0, 0, len(gts)
    if not gts:
        return 0, len(dets), 0
    m = match_segments([d[:8] for d in dets], [_flat(g["pts"]) for g in gts], [0, len(dets)], [0, len(gts)],
                       [iou_thr], det_cls=[int(d[8]) for d in dets], gt_cls=[int(g["cls"]) for g in gts])[0]
    tp = int((m >= 0).sum())
    return tp, len(dets) - tp, len(gts) - tp


def _prec_rec_f1(tp, fp, fn):
    """Detect_OBB.py:482-486."""
    P = tp / (tp + fp + 1e-9)
    R = tp / (tp + fn + 1e-9)
    return P, R, 2 * P * R / (P + R + 1e-9)


def compute_ap_from_pr(recall, precision):
    """Detect_OBB.py:489-499: area under the monotone precision envelope."""
    mrec = np.concatenate(([0.0], recall, [1.0]))
    mpre = np.concatenate(([0.0], precision, [0.0]))
    mpre = np.maximum.accumulate(mpre[::-1])[::-1]
    idx = np.where(mrec[1:] != mrec[:-1])[0]
    return np.sum((mrec[idx + 1] - mrec[idx]) * mpre[idx + 1])


def gather_detections_and_gts(dets_source, all_images, cls_id):
    """Detect_OBB.py:501-510."""
    cfg = _cfg()
    dets, gts = [], {}
    for img_path in all_images:
        for d in dets_source.get(img_path, []):
            if int(d[8]) == cls_id and d[9] >= cfg.MAP_MIN_SCORE:
                dets.append({"image_id": img_path, "score": float(d[9]), "bbox": d[:8]})
        gts[img_path] = [_flat(g["pts"]) for g in _load_gt_as_pixels(img_path) if g["cls"] == cls_id]
    return dets, gts


def _pr_curves(dets, gts, iou_list):
    """compute_pr_for_class for several IoU thresholds with one IoU evaluation.  Returns a list of the
    reference's 8-tuples, one per threshold."""
    npos = sum(len(v) for v in gts.values())
    if npos == 0:
        return [(np.array([0.0]), np.array([0.0]), 0.0, 0.0, 0.0, 0, 0, 0) for _ in iou_list]
    order = sorted(range(len(dets)), key=lambda i: dets[i]["score"], reverse=True)       # stable, like sorted()
    if not order:
        return [(np.array([0.0]), np.array([0.0]), 0.0, 0.0, 0.0, 0, 0, npos) for _ in iou_list]
    # segments = images; inside an image the detections keep the global score order
    images = list(gts.keys())
    per_img = {img: [] for img in images}
    extra = []
    for rank, i in enumerate(order):
        img = dets[i]["image_id"]
        (per_img[img] if img in per_img else extra).append(rank)      # an image without a gts entry can only yield FPs
    det_rows, det_off, gt_rows, gt_off, rank_of_row = [], [0], [], [0], []
    for img in images:
        for rank in per_img[img]:
            det_rows.append(dets[order[rank]]["bbox"])
            rank_of_row.append(rank)
        det_off.append(len(det_rows))
        gt_rows.extend(gts[img])
        gt_off.append(len(gt_rows))
    m = match_segments(det_rows, gt_rows, det_off, gt_off, iou_list) if det_rows else np.zeros((len(iou_list), 0), np.int32)
    out = []
    rank_of_row = np.asarray(rank_of_row, dtype=np.int64)
    for t in range(len(iou_list)):
        tp = np.zeros(len(order))
        tp[rank_of_row[m[t] >= 0]] = 1
        fp = 1.0 - tp
        tp_cum, fp_cum = np.cumsum(tp), np.cumsum(fp)
        recall = tp_cum / (npos + 1e-9)
        precision = tp_cum / (tp_cum + fp_cum + 1e-9)
        ap = compute_ap_from_pr(recall, precision)
        total_tp, total_fp = int(tp_cum[-1]), int(fp_cum[-1])
        out.append((precision, recall, ap, total_tp, total_fp, npos - total_tp))
    return out


def compute_pr_for_class(dets, gts, iou_thr=0.5):
    """Detect_OBB.py:512-565 -> (precision, recall, ap, TP, FP, FN)."""
    return _pr_curves(dets, gts, [iou_thr])[0]


def _gt_class_ids(all_images):
    """Detect_OBB.py:567-572."""
    return sorted({int(g["cls"]) for img in all_images for g in _load_gt_as_pixels(img)})


def evaluate_map(dets_source, all_images, iou_list=None):
    """Detect_OBB.py:574-607 (the wide ``all_dets_per_image_map`` source is preferred when it exists)."""
    cfg = _cfg()
    if iou_list is None:
        iou_list = [0.5] + [round(0.5 + 0.05 * i, 2) for i in range(1, 10)]
    iou_list = list(iou_list)
    use_source = getattr(cfg, "all_dets_per_image_map", dets_source)
    aps = {iou: [] for iou in iou_list}
    for cid in _gt_class_ids(all_images):
        dets, gts = gather_detections_and_gts(use_source, all_images, cid)
        for iou, res in zip(iou_list, _pr_curves(dets, gts, iou_list)):
            aps[iou].append(res[2])
    per_iou_map = {iou: (float(np.mean(v)) if v else 0.0) for iou, v in aps.items()}
    return {"mAP@0.5": per_iou_map.get(0.5, 0.0),
            "mAP@[0.5:0.95]": float(np.mean([per_iou_map[i] for i in iou_list])) if iou_list else 0.0,
            "per_iou": per_iou_map}


def evaluate_center_hit(all_images, conf_thr=0.5):
    """Detect_OBB.py:609-648."""
    cfg = _cfg()
    det_rows, det_cls, det_off, gt_rows, gt_cls, gt_off = [], [], [0], [], [], [0]
    for img_path in all_images:
        for d in cfg.all_dets_per_image.get(img_path, []):
            if d[9] >= conf_thr:
                det_rows.append(d[:8]); det_cls.append(int(d[8]))
        det_off.append(len(det_rows))
        for g in _load_gt_as_pixels(img_path):
            gt_rows.append(_flat(g["pts"])); gt_cls.append(int(g["cls"]))
        gt_off.append(len(gt_rows))
    m = center_hit_segments(det_rows, det_cls, gt_rows, gt_cls, det_off, gt_off)
    tp = int((m >= 0).sum())
    fp, fn = len(det_rows) - tp, len(gt_rows) - tp
    P, R, F1 = _prec_rec_f1(tp, fp, fn)
    print(f"[Center-Hit @ conf≥{conf_thr:.2f}] P={P:.3f} R={R:.3f} F1={F1:.3f} (TP={tp}, FP={fp}, FN={fn})")
    return P, R, F1


def _dataset_counts(dets_source, all_images, conf_thr, iou_thr, cls_id: Optional[int] = None):
    """TP/FP/FN over a set of images in ONE batched matching (segments = images)."""
    det_rows, det_cls, det_off, gt_rows, gt_cls, gt_off = [], [], [0], [], [], [0]
    for img_path in all_images:
        for d in dets_source.get(img_path, []):
            if d[9] >= conf_thr and (cls_id is None or int(d[8]) == cls_id):
                det_rows.append(d[:8]); det_cls.append(int(d[8]))
        det_off.append(len(det_rows))
        for g in _load_gt_as_pixels(img_path):
            if cls_id is None or g["cls"] == cls_id:
                gt_rows.append(_flat(g["pts"])); gt_cls.append(int(g["cls"]))
        gt_off.append(len(gt_rows))
    if not det_rows:
        return 0, 0, len(gt_rows)
    m = match_segments(det_rows, gt_rows, det_off, gt_off, [iou_thr], det_cls=det_cls, gt_cls=gt_cls)[0]
    tp = int((m >= 0).sum())
    return tp, len(det_rows) - tp, len(gt_rows) - tp


def _evaluate_dataset(all_images, conf_thr, iou_thr):
    """Detect_OBB.py:650-658."""
    return _prec_rec_f1(*_dataset_counts(_cfg().all_dets_per_image, all_images, conf_thr, iou_thr))


def _classwise_report(dets_source, all_images, conf_thr, iou_thr):
    """Detect_OBB.py:660-686: per-class TP/FP/FN/P/R/F1, written to Output/fusion_classwise_metrics.xlsx."""
    cfg = _cfg()
    cids = sorted({int(d[8]) for dets in dets_source.values() for d in dets})
    rows = []
    for cid in cids:
        tp, fp, fn = _dataset_counts(dets_source, all_images, conf_thr, iou_thr, cls_id=cid)
        P, R, F1 = _prec_rec_f1(tp, fp, fn)
        rows.append([cid, cfg.CLASS_NAMES.get(cid, str(cid)), tp, fp, fn, P, R, F1])
    columns = ["cls_id", "class", "TP", "FP", "FN", "Precision", "Recall", "F1"]
    out_path = os.path.join(getattr(cfg, "output_dir", "Output"), "fusion_classwise_metrics.xlsx")
    os.makedirs(os.path.dirname(out_path) or ".", exist_ok=True)
    df = None
    try:
        import pandas as pd
        df = pd.DataFrame(rows, columns=columns)
        df.to_excel(out_path, index=False)
    except (ImportError, ModuleNotFoundError, ValueError):
        cfg._write_xlsx(out_path, columns, rows)
    print(f"[Saved] {out_path}")
    return df if df is not None else rows


def run_fusion_eval(input_dir, iou_thr=None):
    """Detect_OBB.py:688-743: the printed evaluation report over a directory of images."""
    cfg = _cfg()
    iou_thr = cfg.iou_thr if iou_thr is None else iou_thr
    all_images = [os.path.join(input_dir, f) for f in os.listdir(input_dir)
                  if f.lower().endswith((".png", ".jpg", ".jpeg", ".tif", ".tiff"))]
    if not all_images:
        print("[Eval] No images found for evaluation.")
        return
    print(f"Tile size: {cfg.tile_sizes}, Overlap: {cfg.overlaps}")
    single = len(cfg.models) == 1 or len(cfg.tile_sizes) == 1
    thr = float(iou_thr)
    if single:
        print(f"[Single-scale] Using manual threshold: {thr:.2f}")
    else:
        print("[Fusion] scale-agnostic merge (late fusion).")
        print(f"[Fusion] Using manual threshold: {thr:.2f}")
    P, R, F1 = _evaluate_dataset(all_images, conf_thr=thr, iou_thr=iou_thr)
    tag = "Report" if single else "Fusion"
    print(f"[{tag} @ {thr:.2f}] Precision={P:.3f} | Recall={R:.3f} | F1={F1:.3f}")
    _classwise_report(cfg.all_dets_per_image, all_images, conf_thr=thr, iou_thr=iou_thr)
    evaluate_center_hit(all_images, conf_thr=thr)
    maps = evaluate_map(cfg.all_dets_per_image, all_images, iou_list=list(np.arange(0.5, 0.96, 0.05)))
    print("[mAP Results]")
    print(f"mAP@0.5 = {maps['mAP@0.5']:.4f}")
    print(f"mAP@[0.5:0.95] = {maps['mAP@[0.5:0.95]']:.4f}")
    soft = [0.30, 0.40, 0.50, 0.60, 0.70]
    maps_soft = evaluate_map(cfg.all_dets_per_image, all_images, iou_list=soft)
    print("[mAP (soft) Results]")
    print(f"mAP@0.3 = {maps_soft['per_iou'][0.30]:.4f}")
    print(f"mAP@[0.3:0.7] = {float(np.mean([maps_soft['per_iou'][i] for i in soft])):.4f}")
