"""Host-side mirror of the reference's training-data tilers and 4-channel conversion (Train_OBB.py),
backed by the CUDA library: same function names, arguments, files written and return values.

What runs where: the per-tile label tables (anchor-in-tile test, bounding-box coverage filter, shift /
clip / normalise - O(tiles x labels) pandas work per image in the reference, Train_OBB.py:93-110 and
:339-355) come from ``gm_train_label_tiles`` for all tiles of an image at once; the DT-Edge plane of the
4-channel TIFF conversion is the same device build as detection's.  Image decode / JPEG / TIFF encode and
the text files are host IO, like in the reference.  Training itself (Ultralytics) is out of scope.
"""
from __future__ import annotations

import ctypes as C
import json
import os
from typing import Dict, List, Tuple

import numpy as np
import torch

from . import _lib as L
from . import ops
from .detect import build_4ch_CHW_from_bgr_dtedge, dt_edge_channel_from_bgr   # noqa: F401  (Train_OBB.py:615-664)

# ----------------------------------------------------------------------------- config (Train_OBB.py:19-37)
CHANNELS = 3
TILE_SIZE = 416
overlap = 100
object_boundary_threshold = 0.1
R_TARGET = 4
IMG_EXTS = (".jpg", ".jpeg", ".png")

LABEL_COLUMNS = ["class", "x1", "y1", "x2", "y2", "x3", "y3", "x4", "y4"]


def read_labels_or_empty(label_path: str, img_w: int, img_h: int) -> np.ndarray:
    """Train_OBB.py:228-261: YOLO-OBB rows -> float64 [n, 9] (class, 8 pixel coordinates); bad rows are skipped.
    (The reference returns a DataFrame with these nine columns; callers here only need the table.)"""
    if (not os.path.exists(label_path)) or os.path.getsize(label_path) == 0:
        return np.zeros((0, 9), dtype=np.float64)
    rows = []
    with open(label_path, "r") as fh:
        for line in fh:
            line = line.split("#", 1)[0].split()
            if len(line) < 9:
                continue
            try:
                vals = [float(v) for v in line[:9]]
            except ValueError:
                continue
            if any(np.isnan(vals)):
                continue
            rows.append(vals)
    if not rows:
        return np.zeros((0, 9), dtype=np.float64)
    t = np.asarray(rows, dtype=np.float64)
    t[:, 1::2] *= float(img_w)
    t[:, 2::2] *= float(img_h)
    return t


def full_tile_grid(H: int, W: int, tile_size: int, ov: int) -> Tuple[int, int, int]:
    rows, cols, span = C.c_int32(), C.c_int32(), C.c_int32()
    n = L.lib.gm_train_tile_grid(int(H), int(W), int(tile_size), int(ov), C.byref(rows), C.byref(cols), C.byref(span))
    if n < 0:
        raise AssertionError("overlap must be < tile_size")
    return int(rows.value), int(cols.value), int(span.value)


def tile_label_tables(labels: np.ndarray, H: int, W: int, tile_size: int, ov: int,
                      threshold: float) -> Dict[int, np.ndarray]:
    """tile_id -> float64 [k, 9] label table of that tile (class + normalised, clipped coordinates), rows in the
    order of ``labels``; only non-empty tiles appear."""
    ops._require_cuda()
    n = int(labels.shape[0])
    rows, cols, span = full_tile_grid(H, W, tile_size, ov)
    if n == 0 or rows * cols == 0:
        return {}
    dev = torch.device("cuda", torch.cuda.current_device())
    per = span * span
    d_lab = torch.from_numpy(np.ascontiguousarray(labels[:, 1:9], dtype=np.float64)).to(dev)
    flag = torch.empty(n * per, dtype=torch.uint8, device=dev)
    tid = torch.empty(n * per, dtype=torch.int32, device=dev)
    coords = torch.empty((n * per, 8), dtype=torch.float64, device=dev)
    L.check(L.lib.gm_train_label_tiles(ops._ptr(d_lab), n, int(H), int(W), int(tile_size), int(ov), float(threshold),
                                       ops._ptr(flag), ops._ptr(tid), ops._ptr(coords), ops._stream()),
            "gm_train_label_tiles")
    sel = torch.nonzero(flag).squeeze(1)
    lab = (sel // per).to(torch.int64)
    key = tid[sel].to(torch.int64) * n + lab
    order = torch.argsort(key)
    sel, lab = sel[order], lab[order]
    t_host = tid[sel].cpu().numpy()
    l_host = lab.cpu().numpy()
    c_host = coords[sel].cpu().numpy()
    out: Dict[int, np.ndarray] = {}
    if len(t_host):
        cuts = np.nonzero(np.diff(t_host))[0] + 1
        for a, b in zip(np.concatenate(([0], cuts)), np.concatenate((cuts, [len(t_host)]))):
            out[int(t_host[a])] = np.concatenate([labels[l_host[a:b], 0:1], c_host[a:b]], axis=1)
    return out


def _write_label_table(path: str, table: np.ndarray, int_class: bool) -> None:
    """DataFrame.to_csv(sep=' ', header=False, index=False): shortest round-trip floats, class as read."""
    with open(path, "w") as fh:
        for row in table:
            cls = repr(int(row[0])) if int_class else repr(float(row[0]))
            fh.write(" ".join([cls] + [repr(float(v)) for v in row[1:]]) + "\n")


def update_txt_file(txt_file, new_paths):
    """Train_OBB.py:263-269."""
    with open(txt_file, "w") as fh:
        for p in new_paths:
            fh.write(f"{p}\n")


def _image_files(image_dir, exts):
    return [f for f in os.listdir(image_dir) if f.lower().endswith(exts)]


def enumerate_and_save_nonempty_tiles(image_dir, label_dir, output_image_dir, output_label_dir, out_list_txt,
                                      tile_size=128, overlap=50, rng_seed=42, split_name="train",
                                      empty_meta_path="datasets/GeoMap/_empty_meta_train.json"):
    """Train_OBB.py:44-146: saves every non-empty full tile (JPG + label table), records the empty ones."""
    import cv2
    os.makedirs(output_image_dir, exist_ok=True)
    os.makedirs(output_label_dir, exist_ok=True)
    stride = tile_size - overlap
    assert stride > 0, "overlap must be < tile_size"
    new_paths, empty_meta = [], []
    P_total = E_total = 0
    for image_file in _image_files(image_dir, (".jpg", ".jpeg", ".png")):
        img = cv2.imread(os.path.join(image_dir, image_file))
        if img is None:
            print(f"[WARN] cannot read: {image_file}")
            continue
        H, W = img.shape[:2]
        stem = os.path.splitext(image_file)[0]
        labels = read_labels_or_empty(os.path.join(label_dir, stem + ".txt"), img_w=W, img_h=H)
        int_class = bool(len(labels)) and bool(np.all(labels[:, 0] == np.floor(labels[:, 0])))
        tables = tile_label_tables(labels, H, W, tile_size, overlap, object_boundary_threshold)
        rows, cols, _ = full_tile_grid(H, W, tile_size, overlap)
        pos = emp = 0
        for tile_id in range(rows * cols):
            y, x = (tile_id // cols) * stride, (tile_id % cols) * stride
            if tile_id in tables:
                op_img = os.path.join(output_image_dir, f"{stem}_tile_{tile_id}.jpg")
                cv2.imwrite(op_img, img[y:y + tile_size, x:x + tile_size])
                _write_label_table(os.path.join(output_label_dir, f"{stem}_tile_{tile_id}.txt"), tables[tile_id], int_class)
                new_paths.append(op_img)
                pos += 1
            else:
                empty_meta.append({"image_file": image_file, "tile_id": int(tile_id), "x": int(x), "y": int(y),
                                   "tile_size": int(tile_size)})
                emp += 1
        P_total += pos
        E_total += emp
        print(f"[TILED] {image_file} -> tiles: {pos + emp} (positives saved: {pos}, empties enumerated: {emp})")
    update_txt_file(out_list_txt, new_paths)
    with open(empty_meta_path, "w") as fh:
        json.dump({"image_dir": image_dir, "output_image_dir": output_image_dir, "output_label_dir": output_label_dir,
                   "empty": empty_meta}, fh)
    print(f"[{split_name}] PASS-1 done. Positives saved: {P_total:,} | Empty enumerated: {E_total:,}")
    return {"P_total": P_total, "E_total": E_total, "empty_meta_path": empty_meta_path}


def crop_images_and_labels(image_dir, label_dir, output_image_dir, output_label_dir, txt_file, cropped_txt_file,
                           tile_size=512, overlap=0, keep_empty_fraction=None, rng_seed=42, split_name="train",
                           boundary_threshold=None):
    """Train_OBB.py:290-428: all non-empty full tiles plus a seeded fraction of the empty ones."""
    import cv2
    if boundary_threshold is None:
        boundary_threshold = object_boundary_threshold
    os.makedirs(output_image_dir, exist_ok=True)
    os.makedirs(output_label_dir, exist_ok=True)
    stride = tile_size - overlap
    assert stride > 0, "overlap must be < tile_size"
    all_tiles: List[dict] = []
    images = {}
    for image_file in _image_files(image_dir, (".jpg", ".png", ".jpeg")):
        image = cv2.imread(os.path.join(image_dir, image_file))
        if image is None:
            print(f"[WARN] cannot read image: {image_file}")
            continue
        images[image_file] = image
        h, w = image.shape[:2]
        labels = read_labels_or_empty(os.path.join(label_dir, os.path.splitext(image_file)[0] + ".txt"), img_w=w, img_h=h)
        int_class = bool(len(labels)) and bool(np.all(labels[:, 0] == np.floor(labels[:, 0])))
        tables = tile_label_tables(labels, h, w, tile_size, overlap, boundary_threshold)
        rows, cols, _ = full_tile_grid(h, w, tile_size, overlap)
        for tile_id in range(rows * cols):
            all_tiles.append({"image_file": image_file, "tile_id": tile_id, "x": (tile_id % cols) * stride,
                              "y": (tile_id // cols) * stride, "is_empty": tile_id not in tables,
                              "tile_labels": tables.get(tile_id), "int_class": int_class})
        print(f"[ENUM] {split_name}:{image_file} -> tiles: {rows * cols}")
    total_tiles = len(all_tiles)
    total_empty = sum(1 for t in all_tiles if t["is_empty"])
    total_nonempty = total_tiles - total_empty
    if keep_empty_fraction is None or keep_empty_fraction == -1:
        keep_empty_fraction = min(1.0, (R_TARGET * total_nonempty) / total_empty) if total_empty > 0 else 0.0
    print(f"\n[{split_name.upper()}] SUMMARY BEFORE EMPTY REMOVAL:")
    print(f"  Total tiles:        {total_tiles:,}")
    print(f"  Non-empty tiles:    {total_nonempty:,}")
    print(f"  Empty tiles:        {total_empty:,}")
    print(f"  -> keep_empty_fraction = {keep_empty_fraction:.3f} (auto={keep_empty_fraction if keep_empty_fraction is not None else 'n/a'})\n")
    empty_idxs = [i for i, t in enumerate(all_tiles) if t["is_empty"]]
    nonempty_idxs = [i for i, t in enumerate(all_tiles) if not t["is_empty"]]
    rng = np.random.RandomState(rng_seed)
    k = int(round(keep_empty_fraction * len(empty_idxs))) if len(empty_idxs) > 0 else 0
    if 0 <= k < len(empty_idxs):
        rng.shuffle(empty_idxs)
        empty_idxs = empty_idxs[:k]
    keep = set(nonempty_idxs + empty_idxs)
    new_paths = []
    for i, t in enumerate(all_tiles):
        if i not in keep:
            continue
        stem = os.path.splitext(t["image_file"])[0]
        crop = images[t["image_file"]][t["y"]:t["y"] + tile_size, t["x"]:t["x"] + tile_size]
        out_img = os.path.join(output_image_dir, f"{stem}_tile_{t['tile_id']}.jpg")
        out_lbl = os.path.join(output_label_dir, f"{stem}_tile_{t['tile_id']}.txt")
        cv2.imwrite(out_img, crop)
        if t["is_empty"]:
            open(out_lbl, "w").close()
        else:
            _write_label_table(out_lbl, t["tile_labels"], t["int_class"])
        new_paths.append(out_img)
    update_txt_file(cropped_txt_file, new_paths)
    print(f"[{split_name}] saved tiles: {len(new_paths)} | non-empty kept: {len(nonempty_idxs)} | empty kept: {len(empty_idxs)} "
          f"(keep_empty_fraction={keep_empty_fraction:.3f})")


def save_tiff_multipage_from_chw(chw: np.ndarray, out_path: str):
    """Train_OBB.py:271-283."""
    import cv2
    if not hasattr(cv2, "imwritemulti"):
        raise RuntimeError("Your OpenCV build lacks 'imwritemulti'. Install opencv-python with TIFF support.")
    assert chw.ndim == 3 and chw.shape[0] in (4, 6), f"Expected (4,H,W) or (6,H,W), got {chw.shape}"
    pages = [np.ascontiguousarray(chw[c].astype(np.uint8, copy=False)) for c in range(chw.shape[0])]
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    if not cv2.imwritemulti(str(out_path), pages):
        raise RuntimeError(f"cv2.imwritemulti failed for: {out_path}")


def convert_folder_to_4ch_tiff_dtedge(src_img_dir: str, dst_img_dir: str, sigmas=(0, 0.8, 1.6, 3.2), **kwargs) -> list:
    """Train_OBB.py:598-613: every image of a folder -> 4-page TIFF [R, G, B, DT-Edge]."""
    import cv2
    os.makedirs(dst_img_dir, exist_ok=True)
    out_paths = []
    for fn in sorted(os.listdir(src_img_dir)):
        if not fn.lower().endswith(IMG_EXTS):
            continue
        ip = os.path.join(src_img_dir, fn)
        bgr = cv2.imread(ip, cv2.IMREAD_COLOR)
        if bgr is None:
            print(f"[WARN] cannot read: {ip}")
            continue
        op = os.path.join(dst_img_dir, os.path.splitext(fn)[0] + ".tiff")
        save_tiff_multipage_from_chw(build_4ch_CHW_from_bgr_dtedge(bgr, sigmas=sigmas, **kwargs), op)
        out_paths.append(os.path.abspath(op))
    return out_paths
