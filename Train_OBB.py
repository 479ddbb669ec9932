# -*- coding: utf-8 -*-
"""Dataset preparation for multi-channel OBB training - drop-in for the data half of the reference's
``Train_OBB.py`` on top of geomap_b200.

The reference's script tiles the labelled maps, optionally converts the tiles to 4-channel
[R, G, B, DT-Edge] TIFFs and then calls Ultralytics' trainer.  The tilers, the label tables and the DT-Edge
conversion are the CUDA-backed mirrors in ``oriented_object_detection_b200.train`` (same names, arguments and
files).  The training call itself (``YOLO(...).train``, Train_OBB.py:792-841) needs Ultralytics and is outside
this repository's scope; ``main`` prepares the dataset and stops there.
"""
import os

from oriented_object_detection_b200 import train as _tr
from oriented_object_detection_b200.train import (  # noqa: F401  (the reference's module-level API)
    enumerate_and_save_nonempty_tiles, crop_images_and_labels, read_labels_or_empty, update_txt_file,
    save_tiff_multipage_from_chw, convert_folder_to_4ch_tiff_dtedge, dt_edge_channel_from_bgr,
    build_4ch_CHW_from_bgr_dtedge)

# Config (Train_OBB.py:19-37 of the reference)
CHANNELS = 3               # 3 or 4
need_cropping = True
TILE_SIZE = 416
overlap = 100
object_boundary_threshold = 0.1
R_TARGET = 4
MS_SIGMAS = (0, 0.6, 1.2, 2.4)

DATA_ROOT = "datasets/GeoMap"


def main():
    _tr.object_boundary_threshold, _tr.R_TARGET = object_boundary_threshold, R_TARGET
    for split in ("train", "val"):
        src_img, src_lbl = os.path.join(DATA_ROOT, "images", split), os.path.join(DATA_ROOT, "labels", split)
        if not os.path.isdir(src_img):
            print(f"[Info] {src_img} not found: nothing to tile")
            continue
        out_img = os.path.join(DATA_ROOT, "cropped", "images", split)
        out_lbl = os.path.join(DATA_ROOT, "cropped", "labels", split)
        if need_cropping:
            crop_images_and_labels(src_img, src_lbl, out_img, out_lbl, os.path.join(DATA_ROOT, f"{split}.txt"),
                                   os.path.join(DATA_ROOT, f"{split}_cropped.txt"), tile_size=TILE_SIZE, overlap=overlap,
                                   keep_empty_fraction=None, split_name=split, boundary_threshold=object_boundary_threshold)
        if CHANNELS == 4:
            convert_folder_to_4ch_tiff_dtedge(out_img, os.path.join(DATA_ROOT, "cropped4ch", "images", split),
                                              sigmas=MS_SIGMAS, bin_method="percentile", p_hi=90, p_lo=65, morph_open=1)
    print("[Info] dataset prepared; run Ultralytics' trainer on it (Train_OBB.py:792-841 of the reference)")


if __name__ == "__main__":
    main()
