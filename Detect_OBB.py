# -*- coding: utf-8 -*-
"""Multi-channel OBB detection - drop-in for the reference's ``Detect_OBB.py`` on top of geomap_b200.

Same config block, same function names at module level, same outputs (``Output/<name>_detected.jpg``,
``Output/<name>.xlsx``, optional evaluation report).  Every hot function is the CUDA-backed mirror in
``oriented_object_detection_b200.detect`` / ``.evaluate``; this file only holds the configuration and the
main loop (the reference runs them at import; here they run under ``__main__`` so the module can be
imported).  Edit the config block the way the reference's README describes.
"""
import os
import time

from oriented_object_detection_b200 import detect as _gm
from oriented_object_detection_b200.detect import (  # noqa: F401  (the reference's module-level API)
    build_multich, run_inference_on_crop, compute_angle_from_bbox, compute_polygon_iou, margin_for,
    box_center_from_xyxyxyxy, center_inside_safe_region, merge_detections, detect_symbols, process_image,
    cross_scale_consensus_filter, _label_path_for_image, _load_gt_as_pixels, _match_dets_to_gts_pixel, _prec_rec_f1,
    compute_ap_from_pr, gather_detections_and_gts, compute_pr_for_class, _gt_class_ids, evaluate_map,
    evaluate_center_hit, _evaluate_dataset, _classwise_report, run_fusion_eval, CLASS_NAMES, CLASS_COLORS)

# =========================
# Config (Detect_OBB.py:20-72 of the reference)
# =========================
calculate_metrics = False
tile_sizes = [128, 416]
overlaps = [30, 100]
MODEL_FILES = ["best128.pt", "best416.pt"]

channels = 3        # 3 or 4
MS_SIGMAS = (0, 0.6, 1.2, 2.4)
DT_BIN_METHOD = "percentile"
DT_P_HI, DT_P_LO = 90, 65
DT_MORPH_OPEN = 1

MAP_MIN_SCORE = 0.001
iou_thr = 0.25   # Metrics
iou_threshold = 0.4  # Merge

APPLY_BORDER_FILTER = True
MARGIN_128 = 10
MARGIN_416 = 20

input_dir = "Input"
output_dir = "Output"


OFFLINE_TAG = "_OFFLINE-RANDOM-INIT"


def load_models():
    """The reference's ``[YOLO("best128.pt"), YOLO("best416.pt")]`` (Detect_OBB.py:26).  Like the reference, a missing
    Ultralytics install or checkpoint is an error: ``ImportError`` / ``FileNotFoundError``.

    Offline opt-in (the checkpoints are Google-Drive downloads): ``GM_OFFLINE_MODEL=yolo11n`` runs a random-init
    YOLO11n-OBB (the real architecture restated in PyTorch, BatchNorm statistics taken from the first batch of tiles)
    and ``GM_OFFLINE_MODEL=standin`` a small stand-in CNN, both behind the device-resident batched predictor.  Their
    detections are meaningless - the data path is the real one - so every output file of such a run carries the
    ``_OFFLINE-RANDOM-INIT`` suffix."""
    offline = os.environ.get("GM_OFFLINE_MODEL", "")
    if not offline:
        missing = [f for f in MODEL_FILES if not os.path.exists(f)]
        if missing:
            raise FileNotFoundError(f"checkpoint(s) not found: {missing}; set GM_OFFLINE_MODEL=yolo11n|standin to run the "
                                    "pipeline with a random-init network (outputs are tagged)")
        try:
            from ultralytics import YOLO
        except ImportError as e:
            raise ImportError("ultralytics is required to load the checkpoints; set GM_OFFLINE_MODEL=yolo11n|standin to "
                              "run the pipeline with a random-init network (outputs are tagged)") from e
        _gm.output_tag = ""
        return [YOLO(f) for f in MODEL_FILES]
    if offline not in ("yolo11n", "standin"):
        raise ValueError(f"GM_OFFLINE_MODEL={offline!r}: expected 'yolo11n' or 'standin'")
    import torch
    from oriented_object_detection_b200.predictor import StandInOBBNet, TilePredictor
    torch.manual_seed(0)
    _gm.output_tag = OFFLINE_TAG
    if offline == "standin":
        print(f"[Info] GM_OFFLINE_MODEL=standin: random-init stand-in predictor; outputs carry the suffix {OFFLINE_TAG}")
        return [TilePredictor(StandInOBBNet(channels, len(CLASS_NAMES)), ts) for ts in tile_sizes]
    from oriented_object_detection_b200.yolo11_obb import random_init_yolo11_obb
    print(f"[Info] GM_OFFLINE_MODEL=yolo11n: random-init YOLO11n-OBB behind the batched predictor; outputs carry the "
          f"suffix {OFFLINE_TAG}")
    return [TilePredictor(random_init_yolo11_obb("n", len(CLASS_NAMES), channels, ts, seed=i), ts)
            for i, ts in enumerate(tile_sizes)]


def _push_config():
    for name in ("calculate_metrics", "tile_sizes", "overlaps", "channels", "MS_SIGMAS", "DT_BIN_METHOD", "DT_P_HI",
                 "DT_P_LO", "DT_MORPH_OPEN", "MAP_MIN_SCORE", "iou_thr", "iou_threshold", "APPLY_BORDER_FILTER",
                 "MARGIN_128", "MARGIN_416", "output_dir"):
        setattr(_gm, name, globals()[name])


def main():
    start_time = time.time()
    _push_config()
    _gm.models = load_models()
    os.makedirs(output_dir, exist_ok=True)
    for fname in os.listdir(input_dir):
        if fname.lower().endswith((".jpg", ".png", ".jpeg", ".tif", ".tiff")):
            print(f"Processing {fname}...")
            process_image(os.path.join(input_dir, fname), output_dir)
            print(f"Results saved for {fname}")
    print(f"--- {time.time() - start_time:.2f} seconds ---")
    if calculate_metrics:
        try:
            run_fusion_eval(input_dir, iou_thr=iou_thr)
        except Exception as e:      # the reference reports and carries on
            print(f"[Eval] Skipped due to error: {e}")


if __name__ == "__main__":
    main()
