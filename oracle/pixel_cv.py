"""OpenCV/numpy port of the reference's DT-Edge builder, for TIMING the CPU path.  TEST INFRASTRUCTURE ONLY.

The reference's CPU implementation of this stage *is* a fixed sequence of OpenCV and numpy
library calls (Detect_OBB.py:95-133); /root/reference cannot travel to the GPU box, so this
module issues the same library calls on the same data there, and bench.py times it as the
"port" CPU baseline (cpu_baseline.kind == "port").  It is checked for equality against the
primitive restatement in oracle/pixel.py (tests/test_oracle_pixel.py), which in turn is pinned
on the lifted reference.
"""
from __future__ import annotations

import numpy as np

import cv2

SIGMAS = (0, 0.6, 1.2, 2.4)


def dt_edge_plane(bgr: np.ndarray, sigmas=SIGMAS, p_hi: float = 90, open_iters: int = 1,
                  bin_method: str = "percentile") -> np.ndarray:
    g = cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY)
    stack = None
    for sg in sigmas:
        src = g if sg <= 0 else cv2.GaussianBlur(g, (0, 0), sg, sg, borderType=cv2.BORDER_REFLECT_101)
        m = cv2.magnitude(cv2.Scharr(src, cv2.CV_32F, 1, 0), cv2.Scharr(src, cv2.CV_32F, 0, 1))
        stack = m if stack is None else np.maximum(stack, m)
    if bin_method == "otsu":                     # Detect_OBB.py:109-111
        stack8 = cv2.normalize(stack, None, 0, 255, cv2.NORM_MINMAX).astype(np.uint8)
        mask = cv2.threshold(stack8, 0, 255, cv2.THRESH_BINARY + cv2.THRESH_OTSU)[1]
    else:
        thr = np.percentile(stack, [p_hi])[0]
        mask = (stack >= thr).astype(np.uint8) * 255
    if open_iters > 0:
        mask = cv2.morphologyEx(mask, cv2.MORPH_OPEN, cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (3, 3)),
                                iterations=open_iters)
    far = cv2.distanceTransform(cv2.bitwise_not(mask), cv2.DIST_L2, 3).astype(np.float32)
    q1, q99 = np.percentile(far, [1, 99])
    far = np.clip((far - q1) / max(1e-6, (q99 - q1)), 0, 1)
    blend = 0.7 * np.exp(-far / 3.0) + 0.3 * cv2.normalize(stack, None, 0, 1, cv2.NORM_MINMAX)
    return (np.clip(blend, 0, 1) * 255).astype(np.uint8)


def build_multich(bgr: np.ndarray, out_channels: int = 3, sigmas=SIGMAS, bin_method: str = "percentile") -> np.ndarray:
    assert out_channels in (3, 4)
    if out_channels == 3:
        return np.ascontiguousarray(bgr)
    return np.ascontiguousarray(np.dstack([cv2.cvtColor(bgr, cv2.COLOR_BGR2RGB),
                                           dt_edge_plane(bgr, sigmas, bin_method=bin_method)]))
