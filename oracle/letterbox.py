"""Ultralytics predictor PRE-processing of one tile, as the reference reaches it through
``model(net_input, conf=...)`` (Detect_OBB.py:76-85).  TEST INFRASTRUCTURE ONLY.

The arithmetic is third-party (ultralytics==8.3.196, requirements.txt:3: ``LetterBox(imgsz, auto=True,
stride=32, scaleup=True)`` + ``BasePredictor.preprocess``), absent from /root/reference and not installable
offline; its published algorithm is restated here on top of the same OpenCV calls it
makes (``cv2.resize(INTER_LINEAR)``, constant border 114), so the only unpinned part is the restated glue:
  r = min(S/h, S/w); new = (round(w r), round(h r)); (dw, dh) = (S - new) mod 32, halved;
  resize iff the size changes; pad top/left = round(d - 0.1), bottom/right = round(d + 0.1) with 114;
  BGR -> RGB only for 3 channels; HWC -> CHW; float32 / 255.
``resize_linear_u8`` restates cv2's 8-bit INTER_LINEAR (11-bit fixed-point coefficients, two passes) in
numpy and is pinned against cv2.resize in tests/test_oracle_letterbox.py.
"""
from __future__ import annotations

import numpy as np


def letterbox_geometry(h: int, w: int, S: int, stride: int = 32, auto: bool = True):
    """(new_h, new_w, top, left, out_h, out_w) of LetterBox for an (h, w) tile."""
    r = min(S / h, S / w)
    new_w, new_h = int(round(w * r)), int(round(h * r))
    dw, dh = S - new_w, S - new_h
    if auto:
        dw, dh = dw % stride, dh % stride
    dw /= 2
    dh /= 2
    top, bottom = int(round(dh - 0.1)), int(round(dh + 0.1))
    left, right = int(round(dw - 0.1)), int(round(dw + 0.1))
    return new_h, new_w, top, left, new_h + top + bottom, new_w + left + right


def _coeffs(src: int, dst: int):
    """cv2 INTER_LINEAR source index and 11-bit coefficient pair per destination index."""
    scale = src / dst
    d = np.arange(dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(np.float32)).astype(np.float32)
    return s, f


def resize_linear_u8(img: np.ndarray, new_h: int, new_w: int) -> np.ndarray:
    """cv2.resize(img, (new_w, new_h), interpolation=cv2.INTER_LINEAR) for uint8 HWC, in integer numpy."""
    h, w = img.shape[:2]
    sx, fx = _coeffs(w, new_w)
    lo = sx < 0
    fx = np.where(lo, np.float32(0), fx); sx = np.where(lo, 0, sx)
    hi = sx >= w - 1
    fx = np.where(hi, np.float32(0), fx); sx = np.where(hi, w - 1, sx)
    a0 = np.rint((np.float32(1) - fx) * np.float32(2048)).astype(np.int64)
    a1 = np.rint(fx * np.float32(2048)).astype(np.int64)
    sx1 = np.minimum(sx + 1, w - 1)
    src = img.astype(np.int64)
    rows = src[:, sx] * a0[None, :, None] + src[:, sx1] * a1[None, :, None]          # [h, new_w, C]
    sy, fy = _coeffs(h, new_h)
    b0 = np.rint((np.float32(1) - fy) * np.float32(2048)).astype(np.int64)
    b1 = np.rint(fy * np.float32(2048)).astype(np.int64)
    y0 = np.clip(sy, 0, h - 1); y1 = np.clip(sy + 1, 0, h - 1)
    out = (((b0[:, None, None] * (rows[y0] >> 4)) >> 16) + ((b1[:, None, None] * (rows[y1] >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def letterbox_u8(tile: np.ndarray, S: int, stride: int = 32, auto: bool = True, use_cv2: bool = True) -> np.ndarray:
    """The uint8 HWC network input before normalisation."""
    h, w = tile.shape[:2]
    new_h, new_w, top, left, out_h, out_w = letterbox_geometry(h, w, S, stride, auto)
    img = tile
    if (h, w) != (new_h, new_w):
        if use_cv2:
            import cv2
            img = cv2.resize(np.ascontiguousarray(tile), (new_w, new_h), interpolation=cv2.INTER_LINEAR)
            if img.ndim == 2:
                img = img[..., None]
        else:
            img = resize_linear_u8(tile, new_h, new_w)
    out = np.full((out_h, out_w, tile.shape[2]), 114, dtype=np.uint8)
    out[top:top + new_h, left:left + new_w] = img
    return out


def preprocess(tile: np.ndarray, S: int, stride: int = 32, auto: bool = True, use_cv2: bool = True) -> np.ndarray:
    """float32 CHW network input of one tile (uint8 HWC, 3 = BGR or 4 = RGB+DT channels)."""
    lb = letterbox_u8(tile, S, stride, auto, use_cv2)
    if lb.shape[2] == 3:
        lb = lb[..., ::-1]
    return np.ascontiguousarray(lb.transpose(2, 0, 1)).astype(np.float32) / np.float32(255)
