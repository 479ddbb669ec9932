"""Host-side mirror of the reference's detection interface, backed by the CUDA library.

Same function names, argument meaning, return types and error behaviour as the module-level
functions of the reference ``Detect_OBB.py`` (numpy arrays and lists of 11-tuples in, the
same out), so a script written against the reference runs unchanged on top of this module.
Every function cites the reference lines it replaces.  Nothing here computes on the CPU:
pixel and geometry work is done by ``libgeomap_b200.so``; Python only marshals.

Config globals mirror Detect_OBB.py:23-72 and are read at call time (edit them the way the
reference README tells users to edit the script).
"""
from __future__ import annotations

import ctypes as C
import os
import time
import zipfile
from typing import Dict, List, Sequence
from xml.sax.saxutils import escape

import numpy as np
import torch

from . import _lib as L
from . import ops

# ----------------------------------------------------------------------------- config (Detect_OBB.py:23-72)
calculate_metrics = False
tile_sizes = [128, 416]
overlaps = [30, 100]
models: list = []                # filled by the entry script (YOLO checkpoints or stand-ins)

channels = 3                     # 3 or 4
MS_SIGMAS = (0, 0.6, 1.2, 2.4)
DT_BIN_METHOD = "percentile"
DT_P_HI, DT_P_LO = 90, 65
DT_MORPH_OPEN = 1

MAP_MIN_SCORE = 0.001
iou_thr = 0.25                   # metrics
iou_threshold = 0.4              # merge

APPLY_BORDER_FILTER = True
MARGIN_128 = 10
MARGIN_416 = 20

all_dets_per_image: Dict[str, list] = {}
output_tag = ""                  # appended to the output file stems; the entry script sets it in offline-model mode so that
                                 # deliverables produced by a random-init network cannot be mistaken for real ones

CLASS_NAMES = {0: "Landslide 1", 1: "Strike", 2: "Spring 1", 3: "Minepit 1", 4: "Hillside", 5: "Feuchte",
               6: "Torf", 7: "Bergsturz", 8: "Landslide 2", 9: "Spring 2", 10: "Spring 3", 11: "Minepit 2"}
CLASS_COLORS = {0: (255, 0, 0), 1: (0, 255, 0), 2: (0, 0, 255), 3: (255, 255, 0), 4: (255, 0, 255),
                5: (0, 255, 255), 6: (0, 0, 0), 7: (240, 34, 0), 8: (50, 20, 60), 9: (60, 50, 20),
                10: (200, 150, 80), 11: (100, 200, 150)}

CONS_IOU_PARTNER, CONS_LOW, CONS_HIGH = 0.40, 0.25, 0.70      # Detect_OBB.py:349-351


def _device() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("geomap_b200 needs an sm_100 CUDA device; there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _check_dtedge_limits(h: int, w: int, morph_open: int) -> None:
    """The two limits of the device DT-Edge builder the reference does not have, reported before the launch with what
    to do about them (the C ABI answers GM_ERANGE): the chamfer scan holds one tile row in a CTA, so a crop may be at
    most GM_MAX_TILE = 1024 px on a side (the reference's tiled path uses 128 / 416; an un-tiled map must be tiled),
    and the cross open is implemented for DT_MORPH_OPEN in 0..8 (the reference's configured value is 1)."""
    if max(h, w) > L.GM_MAX_TILE:
        raise ValueError(f"4-channel DT-Edge build of a {h}x{w} crop: the device builder supports crops up to "
                         f"{L.GM_MAX_TILE} px on a side - tile the map (tile_sizes / need_cropping) instead of passing it whole")
    if not 0 <= int(morph_open) <= 8:
        raise NotImplementedError(f"DT_MORPH_OPEN={morph_open}: the device builder implements 0 to 8 iterations of the cross open")


def _params(layout: int = 0, sigmas=None, p_hi=None, morph_open=None) -> L.gm_dtedge_params:
    return L.make_params(MS_SIGMAS if sigmas is None else sigmas, DT_P_HI if p_hi is None else p_hi,
                         DT_MORPH_OPEN if morph_open is None else morph_open, layout,
                         flags=L.bin_method_flags(DT_BIN_METHOD))


def _strike_class() -> int:
    for k, v in CLASS_NAMES.items():
        if v == "Strike":
            return int(k)
    return -1


# ----------------------------------------------------------------------------- a2/a3

def build_multich(bgr: np.ndarray, out_channels: int = None) -> np.ndarray:
    """Detect_OBB.py:87-133.  3 -> contiguous BGR copy; 4 -> HWC [R,G,B,DT-Edge] uint8."""
    if out_channels is None:
        out_channels = channels
    assert out_channels in (3, 4), f"Unsupported out_channels={out_channels}"
    src = np.ascontiguousarray(bgr, dtype=np.uint8)
    h, w = src.shape[:2]
    out = np.empty((h, w, out_channels), dtype=np.uint8)
    if out_channels == 4:
        _check_dtedge_limits(h, w, DT_MORPH_OPEN)
    p = _params(0) if out_channels == 4 else None
    L.check(L.lib.gm_build_multich_host(src.ctypes.data_as(C.c_void_p), h, w, out_channels,
                                        C.byref(p) if p is not None else None, out.ctypes.data_as(C.c_void_p)),
            "gm_build_multich_host")
    return out


def dt_edge_channel_from_bgr(bgr: np.ndarray, sigmas=(0, 0.6, 1.2, 2.4), bin_method: str = "percentile",
                             p_hi: int = 90, p_lo: int = 65, morph_open: int = 1) -> np.ndarray:
    """Train_OBB.py:615-653: the DT-Edge plane alone (uint8 [h,w]); bin_method "percentile" or "otsu" (:633-638)."""
    return build_4ch_CHW_from_bgr_dtedge(bgr, sigmas=sigmas, bin_method=bin_method, p_hi=p_hi, p_lo=p_lo,
                                         morph_open=morph_open)[3]


def build_4ch_CHW_from_bgr_dtedge(bgr: np.ndarray, sigmas=(0, 0.8, 1.6, 3.2), **kwargs) -> np.ndarray:
    """Train_OBB.py:655-664: (4,H,W) = [R,G,B,DT-Edge]."""
    src = np.ascontiguousarray(bgr, dtype=np.uint8)
    h, w = src.shape[:2]
    out = np.empty((4, h, w), dtype=np.uint8)
    _check_dtedge_limits(h, w, kwargs.get("morph_open", 1))
    p = L.make_params(sigmas, kwargs.get("p_hi", 90), kwargs.get("morph_open", 1), layout=1,
                      flags=L.bin_method_flags(kwargs.get("bin_method", "percentile")))
    L.check(L.lib.gm_build_multich_host(src.ctypes.data_as(C.c_void_p), h, w, 4, C.byref(p),
                                        out.ctypes.data_as(C.c_void_p)), "gm_build_multich_host")
    return out


def run_inference_on_crop(crop_bgr: np.ndarray, model):
    """Detect_OBB.py:76-85."""
    net_input = build_multich(crop_bgr, out_channels=channels)
    with torch.no_grad():
        return model(net_input, conf=0.001 if calculate_metrics else 0.25)


# ----------------------------------------------------------------------------- a7-a10 scalar helpers

def compute_angle_from_bbox(points) -> float:
    """Detect_OBB.py:135-142 (scalar helper kept for API parity; the batched path computes it on the GPU)."""
    x1, y1, _, _, _, _, x4, y4 = points
    a = float(np.arctan2(x4 - x1, y4 - y1) * (180.0 / np.pi))
    return 180 - a if a > 0 else abs(a)


def compute_polygon_iou(box1, box2) -> float:
    """Detect_OBB.py:144-154: rotated IoU of two 8-float corner lists; 0.0 for invalid input."""
    a = (C.c_double * 8)(*[float(v) for v in box1[:8]])
    b = (C.c_double * 8)(*[float(v) for v in box2[:8]])
    out = C.c_double()
    L.check(L.lib.gm_polygon_iou_host(a, b, C.byref(out)), "gm_polygon_iou_host")
    return float(out.value)


def margin_for(tile_size: int) -> int:
    """Detect_OBB.py:156-157."""
    return MARGIN_128 if tile_size <= 128 else MARGIN_416


def box_center_from_xyxyxyxy(points8):
    """Detect_OBB.py:159-165."""
    return ((points8[0] + points8[2] + points8[4] + points8[6]) / 4.0,
            (points8[1] + points8[3] + points8[5] + points8[7]) / 4.0)


def center_inside_safe_region(points8, crop_x0, crop_y0, crop_w, crop_h, margin_px: int) -> bool:
    """Detect_OBB.py:167-174."""
    cx, cy = box_center_from_xyxyxyxy(points8)
    return (margin_px <= cx - crop_x0 <= crop_w - margin_px) and (margin_px <= cy - crop_y0 <= crop_h - margin_px)


# ----------------------------------------------------------------------------- a11 / a12 on lists of tuples

def _arrays(dets: Sequence[tuple], dev, thresholds: Sequence[float] = ()):
    """List of 11-tuples -> device arrays (boxes float64 [n,8], cls int32, conf float32) + the thresholds to use.

    One bulk conversion (no per-detection Python work beyond building the row list).  The reference sorts and
    thresholds on Python floats (float64).  Pipeline confidences are float32 values, which survive the narrowing
    unchanged; for callers that pass genuine float64 confidences the values are replaced by their dense rank
    (order and ties preserved exactly, 2^24 distinct values fit a float32) and every confidence threshold by the
    rank of the first value that reaches it, so ``>=`` decisions and sort order equal the float64 ones
    (float32(0.7) < 0.70 would otherwise drop a detection the reference keeps, Detect_OBB.py:183, :401)."""
    n = len(dets)
    rows = np.asarray([d[:10] for d in dets], dtype=np.float64).reshape(n, 10)
    boxes = np.ascontiguousarray(rows[:, :8])
    cls = rows[:, 8].astype(np.int32)
    conf64 = rows[:, 9]
    conf = conf64.astype(np.float32)
    thr = [float(t) for t in thresholds]
    if n and not np.array_equal(conf.astype(np.float64), conf64):
        uniq, inv = np.unique(conf64, return_inverse=True)
        if uniq.size >= (1 << 24):
            raise ValueError("more than 2^24 distinct float64 confidences")
        conf = inv.astype(np.float32)
        thr = [float(np.searchsorted(uniq, t, side="left")) for t in thr]
    return (torch.from_numpy(boxes).to(dev), torch.from_numpy(cls).to(dev), torch.from_numpy(conf).to(dev), thr)


def merge_detections(detections: list, iou_threshold: float = 0.5) -> list:
    """Detect_OBB.py:176-200: sorts ``detections`` in place (stable, conf desc), returns the kept members."""
    if not detections:
        return []
    dev = _device()
    boxes, cls, conf, _ = _arrays(detections, dev)
    cmin, cmax = int(cls.min().item()), int(cls.max().item())
    order, _, kept = ops.nms_global(boxes, cls - cmin, conf, iou_threshold, max_class=cmax - cmin)
    src = list(detections)
    detections[:] = [src[i] for i in order.cpu().tolist()]
    return [src[i] for i in kept.cpu().tolist()]


def cross_scale_consensus_filter(dets_by_scale: Dict[int, list]) -> list:
    """Detect_OBB.py:347-423."""
    scales = sorted(dets_by_scale.keys())
    if len(scales) == 1:
        return list(dets_by_scale[scales[0]])
    flat, sid = [], []
    for k, s in enumerate(scales):
        flat.extend(dets_by_scale[s])
        sid.extend([k] * len(dets_by_scale[s]))
    if not flat:
        return []
    dev = _device()
    boxes, cls, conf, (low, high) = _arrays(flat, dev, (CONS_LOW, CONS_HIGH))
    cmin, cmax = int(cls.min().item()), int(cls.max().item())
    kept = ops.fuse_scales(boxes, cls - cmin, conf, torch.tensor(sid, dtype=torch.int32, device=dev), len(scales),
                           max_class=cmax - cmin, iou_partner=CONS_IOU_PARTNER, conf_low=low, conf_high=high)
    return [flat[i] for i in kept.cpu().tolist()]


# ----------------------------------------------------------------------------- a1-a9: tiled detection

def detect_symbols(image: np.ndarray, model, tile_size: int, overlap: int) -> list:
    """Detect_OBB.py:202-266.  Output: list of (x1,y1,x2,y2,x3,y3,x4,y4, cls_id, conf, angle).

    All tiles of the map are cut (and, for 4 channels, given their DT-Edge plane) in one batched
    launch; the model is then called once per tile with the same ``(ndarray, conf=)`` protocol
    the reference uses (``results[0].obb`` items with ``.xyxyxyxy/.cls/.conf``), or once per
    batch when it offers ``predict_tiles``; remap, border filter, strike angle and the per-tile
    NMS run batched on the GPU afterwards.
    """
    dev = _device()
    img = np.ascontiguousarray(image[:, :, :3], dtype=np.uint8)
    H, W = img.shape[:2]
    plan = ops.make_plan(H, W, tile_size, overlap, device=dev)
    map_dev = torch.from_numpy(img).to(dev)
    nch = channels
    assert nch in (3, 4), f"Unsupported out_channels={nch}"
    packed = ops.tile_gather(map_dev, plan) if nch == 3 else ops.dtedge_build(map_dev, plan, _params(0))
    conf_thr = 0.001 if calculate_metrics else 0.25

    per_tile_bound = 0            # detections one tile can have, when known: lets the per-tile NMS take its one-CTA-per-tile form
    if hasattr(model, "predict_tiles"):
        # device-resident batched predictor: (packed tiles, plan, channels, conf) -> per-tile lists
        local, cls, conf, tile_id = model.predict_tiles(packed, plan, nch, conf_thr)
        per_tile_bound = int(getattr(model, "max_det", 0) or 0)
    else:
        host = packed.cpu().numpy()
        rows_l, cls_l, conf_l, tid_l = [], [], [], []
        with torch.no_grad():
            for ti, t in enumerate(plan.tiles):
                h, w, off = int(t["h"]), int(t["w"]), int(t["px_off"])
                crop = host[nch * off: nch * (off + h * w)].reshape(h, w, nch)
                results = model(crop, conf=conf_thr)
                for det in results[0].obb:
                    rows_l.append([float(v) for v in det.xyxyxyxy[0].flatten().tolist()])
                    cls_l.append(int(det.cls[0]))
                    conf_l.append(float(det.conf[0]))
                    tid_l.append(ti)
        if not rows_l:
            return []
        local = torch.tensor(rows_l, dtype=torch.float32, device=dev)
        cls = torch.tensor(cls_l, dtype=torch.int32, device=dev)
        conf = torch.tensor(conf_l, dtype=torch.float32, device=dev)
        tile_id = torch.tensor(tid_l, dtype=torch.int32, device=dev)
        per_tile_bound = int(np.bincount(np.asarray(tid_l, dtype=np.int64)).max())
    if local.shape[0] == 0:
        return []
    margin = margin_for(tile_size) if APPLY_BORDER_FILTER else 0
    cmin = int(cls.min().item())
    cmax = int(cls.max().item())
    strike = _strike_class()
    out = ops.tile_postprocess(local, cls - cmin, conf, tile_id, plan, margin, strike - cmin, iou_threshold,
                               max_class=cmax - cmin, max_per_tile=per_tile_bound if 0 < per_tile_bound <= 320 else 0)
    # bulk conversion: ndarray.tolist() yields Python floats / ints (float32 confidences widen exactly)
    b = out["boxes"].cpu().numpy().tolist()
    c = (out["cls"] + cmin).cpu().numpy().tolist()
    f = out["conf"].cpu().numpy().astype(np.float64).tolist()
    a = out["angle"].cpu().numpy().tolist()
    return [(*bi, ci, fi, ai) for bi, ci, fi, ai in zip(b, c, f, a)]


# ----------------------------------------------------------------------------- a13 + outputs

def _write_xlsx(path: str, columns: List[str], rows: List[list]) -> None:
    """Minimal one-sheet workbook with inline strings (pandas has no Excel engine here)."""
    def cell(ref, v):
        if isinstance(v, str):
            return f'<c r="{ref}" t="inlineStr"><is><t>{escape(v)}</t></is></c>'
        return f'<c r="{ref}" t="n"><v>{repr(float(v)) if not isinstance(v, int) else v}</v></c>'

    def col(i):
        s = ""
        i += 1
        while i:
            i, r = divmod(i - 1, 26)
            s = chr(65 + r) + s
        return s

    lines = []
    for r, row in enumerate([columns] + rows, start=1):
        lines.append(f'<row r="{r}">' + "".join(cell(f"{col(c)}{r}", v) for c, v in enumerate(row)) + "</row>")
    sheet = ('<?xml version="1.0" encoding="UTF-8" standalone="yes"?>'
             '<worksheet xmlns="http://schemas.openxmlformats.org/spreadsheetml/2006/main"><sheetData>'
             + "".join(lines) + "</sheetData></worksheet>")
    with zipfile.ZipFile(path, "w", zipfile.ZIP_DEFLATED) as z:
        z.writestr("[Content_Types].xml",
                   '<?xml version="1.0" encoding="UTF-8" standalone="yes"?>'
                   '<Types xmlns="http://schemas.openxmlformats.org/package/2006/content-types">'
                   '<Default Extension="rels" ContentType="application/vnd.openxmlformats-package.relationships+xml"/>'
                   '<Default Extension="xml" ContentType="application/xml"/>'
                   '<Override PartName="/xl/workbook.xml" ContentType="application/vnd.openxmlformats-officedocument.spreadsheetml.sheet.main+xml"/>'
                   '<Override PartName="/xl/worksheets/sheet1.xml" ContentType="application/vnd.openxmlformats-officedocument.spreadsheetml.worksheet+xml"/>'
                   '</Types>')
        z.writestr("_rels/.rels",
                   '<?xml version="1.0" encoding="UTF-8" standalone="yes"?>'
                   '<Relationships xmlns="http://schemas.openxmlformats.org/package/2006/relationships">'
                   '<Relationship Id="rId1" Type="http://schemas.openxmlformats.org/officeDocument/2006/relationships/officeDocument" Target="xl/workbook.xml"/>'
                   '</Relationships>')
        z.writestr("xl/workbook.xml",
                   '<?xml version="1.0" encoding="UTF-8" standalone="yes"?>'
                   '<workbook xmlns="http://schemas.openxmlformats.org/spreadsheetml/2006/main" '
                   'xmlns:r="http://schemas.openxmlformats.org/officeDocument/2006/relationships">'
                   '<sheets><sheet name="Sheet1" sheetId="1" r:id="rId1"/></sheets></workbook>')
        z.writestr("xl/_rels/workbook.xml.rels",
                   '<?xml version="1.0" encoding="UTF-8" standalone="yes"?>'
                   '<Relationships xmlns="http://schemas.openxmlformats.org/package/2006/relationships">'
                   '<Relationship Id="rId1" Type="http://schemas.openxmlformats.org/officeDocument/2006/relationships/worksheet" Target="worksheets/sheet1.xml"/>'
                   '</Relationships>')
        z.writestr("xl/worksheets/sheet1.xml", sheet)


XLSX_COLUMNS = ["Class", "X1", "Y1", "X2", "Y2", "X3", "Y3", "X4", "Y4", "Confidence", "Angle"]   # Detect_OBB.py:328


def process_image(image_path: str, output_dir: str):
    """Detect_OBB.py:268-345: per-scale detection -> fusion -> global NMS -> JPG + XLSX."""
    import cv2
    t0 = time.time()
    image = cv2.imread(image_path)
    if image is None:
        print(f"[Warn] Could not read image: {image_path}")
        return
    dets_by_scale = {}
    for ts, ov, model in zip(tile_sizes, overlaps, models):
        dets_by_scale[ts] = detect_symbols(image, model, ts, ov)
    merged_for_map = None
    if calculate_metrics:
        pooled = [d for s in dets_by_scale for d in dets_by_scale[s]]
        merged_for_map = merge_detections(pooled, iou_threshold)
        consensus = cross_scale_consensus_filter({s: list(v) for s, v in dets_by_scale.items()})
    else:
        consensus = cross_scale_consensus_filter(dets_by_scale)
    merged = merge_detections(consensus, iou_threshold)
    print(f"--- {time.time() - t0:.3f} seconds ---")

    canvas = image.copy()
    name = os.path.basename(image_path)
    Hh, Ww = canvas.shape[:2]
    rows = []
    for (x1, y1, x2, y2, x3, y3, x4, y4, cls_id, conf, angle) in merged:
        color = tuple(int(c) for c in CLASS_COLORS.get(cls_id, (0, 255, 255)))
        label = CLASS_NAMES.get(cls_id, f"Class{cls_id}")
        quad = np.array([[x1, y1], [x2, y2], [x3, y3], [x4, y4]], dtype=np.int32)
        cv2.polylines(canvas, [quad], isClosed=True, color=color, thickness=2)
        tx = int(max(0, min(Ww - 1, round(min(x1, x2, x3, x4)))))
        ty = int(max(0, min(Hh - 1, round(min(y1, y2, y3, y4) - 10))))
        cv2.putText(canvas, f"{label} {conf:.2f}", (tx, ty), cv2.FONT_HERSHEY_SIMPLEX, 0.5, color, 2,
                    lineType=cv2.LINE_AA)
        rows.append([label, x1, y1, x2, y2, x3, y3, x4, y4, conf, angle])
    cv2.imwrite(os.path.join(output_dir, name.replace(".jpg", f"_detected{output_tag}.jpg")
                             .replace(".png", f"_detected{output_tag}.jpg")), canvas)
    xlsx = os.path.join(output_dir, name.replace(".jpg", f"{output_tag}.xlsx").replace(".png", f"{output_tag}.xlsx"))
    try:
        import pandas as pd
        pd.DataFrame(rows, columns=XLSX_COLUMNS).to_excel(xlsx, index=False)
    except (ImportError, ModuleNotFoundError, ValueError):
        _write_xlsx(xlsx, XLSX_COLUMNS, rows)
    if calculate_metrics:
        g = globals()
        g.setdefault("all_dets_per_image_pr", {})[image_path] = merged
        g.setdefault("all_dets_per_image_map", {})[image_path] = merged_for_map
    all_dets_per_image[image_path] = merged


# ----------------------------------------------------------------------------- evaluation path (Detect_OBB.py:425-743)
output_dir = "Output"
from .evaluate import (_label_path_for_image, _load_gt_as_pixels, _match_dets_to_gts_pixel, _prec_rec_f1,   # noqa: E402,F401
                       compute_ap_from_pr, gather_detections_and_gts, compute_pr_for_class, _gt_class_ids,
                       evaluate_map, evaluate_center_hit, _evaluate_dataset, _classwise_report, run_fusion_eval)
