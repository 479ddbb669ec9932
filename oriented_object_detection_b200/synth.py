"""Synthetic inputs for tests and benchmarks (no datasets or weights are available offline).

* maps: integer-only, counter-hash based, so the same (seed, y, x) gives the same pixel on the
  CPU and on any GPU rank without moving data: paper-like low-frequency colour field
  (200..245), +-6 noise, sparse dark line symbols 1-3 px wide and 20-100 px long (about one
  per 128x128 cell).  SURVEY.md section 8(d).
* OBB sets: objects with the size range seen in the reference's Output/*.xlsx (w 12-100,
  h 11-97 px), angle in [-pi/4, 3pi/4), one jittered copy per tile whose safe region contains
  the centre (the seam duplicates the global NMS exists to remove).
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch

_M32 = 0xFFFFFFFF


def _mix(a: torch.Tensor, b: torch.Tensor, c, seed: int) -> torch.Tensor:
    """32-bit avalanche hash of (a, b, c, seed) on int64 tensors (wrap-around products, masked)."""
    h = (a * 0x9E3779B1 + b * 0x85EBCA77 + c * 0xC2B2AE3D + seed * 0x27D4EB2F) & _M32
    h = h ^ (h >> 15)
    h = (h * 0x2C1B3C6D) & _M32
    h = h ^ (h >> 12)
    h = (h * 0x297A2D39) & _M32
    return h ^ (h >> 15)


def synthetic_map(H: int, W: int, seed: int, device="cpu", row0: int = 0, rows: Optional[int] = None,
                  band: int = 1024, col0: int = 0, cols: Optional[int] = None) -> torch.Tensor:
    """uint8 [rows, cols, 3] BGR: rows [row0, row0+rows) x cols [col0, col0+cols) of the H x W map with this seed."""
    rows = H - row0 if rows is None else rows
    cols = W - col0 if cols is None else cols
    Wfull, W = W, cols
    del Wfull
    out = torch.empty((rows, W, 3), dtype=torch.uint8, device=device)
    xs = torch.arange(col0, col0 + W, dtype=torch.int64, device=device).unsqueeze(0)
    for r in range(0, rows, band):
        nb = min(band, rows - r)
        ys = torch.arange(row0 + r, row0 + r + nb, dtype=torch.int64, device=device).unsqueeze(1)
        X = xs.expand(nb, W)
        Y = ys.expand(nb, W)
        # --- background: bilinear blend of hashed node colours on a 64-px lattice
        gx, fx = X // 64, X % 64
        gy, fy = Y // 64, Y % 64
        chans = []
        for ch in range(3):
            def node(ix, iy):
                return 200 + _mix(ix, iy, ch, seed) % 46
            v = ((64 - fx) * (64 - fy) * node(gx, gy) + fx * (64 - fy) * node(gx + 1, gy)
                 + (64 - fx) * fy * node(gx, gy + 1) + fx * fy * node(gx + 1, gy + 1)) >> 12
            v = v + (_mix(X, Y, 7 + ch, seed) % 13) - 6
            chans.append(v)
        # --- symbols: one hashed segment per 128-px cell, looked up in the 3x3 neighbourhood
        cx0, cy0 = X // 128, Y // 128
        dark = torch.zeros((nb, W), dtype=torch.bool, device=device)
        ink = torch.zeros((nb, W), dtype=torch.int64, device=device)
        for oy in (-1, 0, 1):
            for ox in (-1, 0, 1):
                cx, cy = cx0 + ox, cy0 + oy
                h0 = _mix(cx, cy, 100, seed)
                present = (h0 % 10) < 8
                px = cx * 128 + (_mix(cx, cy, 101, seed) % 128)
                py = cy * 128 + (_mix(cx, cy, 102, seed) % 128)
                dx = (_mix(cx, cy, 103, seed) % 141) - 70
                dy = (_mix(cx, cy, 104, seed) % 141) - 70
                small = (dx * dx + dy * dy) < 400
                dx = torch.where(small, dx + 25, dx)
                dy = torch.where(small, dy + 25, dy)
                wd = 1 + (_mix(cx, cy, 105, seed) % 3)
                l2 = dx * dx + dy * dy
                rx, ry = X - px, Y - py
                t = torch.clamp(rx * dx + ry * dy, min=0)
                t = torch.minimum(t, l2)
                ex = rx * l2 - dx * t
                ey = ry * l2 - dy * t
                hit = present & (4 * (ex * ex + ey * ey) <= wd * wd * l2 * l2)
                ink = torch.where(hit & ~dark, 20 + (h0 >> 8) % 40, ink)
                dark |= hit
        for ch in range(3):
            v = torch.where(dark, ink + (_mix(X, Y, 11 + ch, seed) % 9), chans[ch])
            out[r:r + nb, :, ch] = torch.clamp(v, 0, 255).to(torch.uint8)
    return out


def synthetic_map_numpy(H: int, W: int, seed: int) -> np.ndarray:
    return synthetic_map(H, W, seed, device="cpu").numpy()


def _corners(cx, cy, w, h, th):
    """xywhr -> 4 corners, Ultralytics corner order (pt1 = c+v1+v2, pt2 = c+v1-v2, ...)."""
    c, s = np.cos(th), np.sin(th)
    v1x, v1y = w / 2 * c, w / 2 * s
    v2x, v2y = -h / 2 * s, h / 2 * c
    return np.stack([cx + v1x + v2x, cy + v1y + v2y, cx + v1x - v2x, cy + v1y - v2y,
                     cx - v1x - v2x, cy - v1y - v2y, cx - v1x + v2x, cy - v1y + v2y], axis=1)


def synthetic_obbs(n_objects: int, H: int, W: int, n_classes: int = 15, seed: int = 0, dup_prob: float = 0.7,
                   ) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Global-coordinate boxes with jittered near-duplicates: (boxes float64 [n,8], cls int32, conf float32).

    The boxes are fp32 values widened to float64, like the reference's tuples.
    """
    rng = np.random.default_rng(seed)
    cx = rng.uniform(0, W, n_objects)
    cy = rng.uniform(0, H, n_objects)
    w = rng.uniform(12, 100, n_objects)
    h = rng.uniform(11, 97, n_objects)
    th = rng.uniform(-np.pi / 4, 3 * np.pi / 4, n_objects)
    cls = rng.integers(0, n_classes, n_objects)
    conf = rng.uniform(0.25, 1.0, n_objects)
    copies = 1 + (rng.random(n_objects) < dup_prob) + (rng.random(n_objects) < dup_prob * 0.4)
    idx = np.repeat(np.arange(n_objects), copies)
    m = idx.size
    b = _corners(cx[idx] + rng.normal(0, 1.5, m), cy[idx] + rng.normal(0, 1.5, m),
                 w[idx] * rng.uniform(0.97, 1.03, m), h[idx] * rng.uniform(0.97, 1.03, m),
                 th[idx] + np.deg2rad(rng.uniform(-2, 2, m)))
    cf = np.clip(conf[idx] + rng.uniform(-0.05, 0.05, m), 0.01, 0.999)
    perm = rng.permutation(m)
    return (b[perm].astype(np.float32).astype(np.float64), cls[idx][perm].astype(np.int32),
            cf[perm].astype(np.float32))


def synthetic_tile_dets(plan, n_objects: int, n_classes: int = 15, seed: int = 0, margin: int = 20):
    """Per-tile detector output for a tile plan: every object is reported (jittered) by each tile
    whose safe region contains its centre.  Returns tile-local corners float32 [n,8], cls int32,
    conf float32, tile_id int32 (non-decreasing); within a tile the order is confidence-descending
    like the Ultralytics predictor's.
    """
    rng = np.random.default_rng(seed)
    H, W = plan.H, plan.W
    cx = rng.uniform(0, W, n_objects)
    cy = rng.uniform(0, H, n_objects)
    w = rng.uniform(12, 100, n_objects)
    h = rng.uniform(11, 97, n_objects)
    th = rng.uniform(-np.pi / 4, 3 * np.pi / 4, n_objects)
    cls = rng.integers(0, n_classes, n_objects)
    conf = rng.uniform(0.25, 1.0, n_objects)
    ty0 = plan.tiles["y0"].astype(np.float64)
    tx0 = plan.tiles["x0"].astype(np.float64)
    th_ = plan.tiles["h"].astype(np.float64)
    tw_ = plan.tiles["w"].astype(np.float64)
    step = max(1, plan.tile_size - plan.overlap)
    cols = plan.cols
    obj_l, tile_l = [], []
    # candidate tiles of an object: the (at most 2x2) tiles whose origin lies within tile_size of the centre
    r_hi = np.minimum((cy // step).astype(np.int64), plan.rows - 1 + plan.row_begin) - plan.row_begin
    c_hi = np.minimum((cx // step).astype(np.int64), cols - 1)
    span = int(np.ceil(plan.tile_size / step))
    for dr in range(span + 1):
        for dc in range(span + 1):
            r, c = r_hi - dr, c_hi - dc
            ok = (r >= 0) & (c >= 0) & (r < plan.rows)
            t = np.where(ok, r * cols + c, 0)
            inside = ok & (cx - tx0[t] >= margin) & (cx - tx0[t] <= tw_[t] - margin) & \
                (cy - ty0[t] >= margin) & (cy - ty0[t] <= th_[t] - margin)
            o = np.nonzero(inside)[0]
            obj_l.append(o)
            tile_l.append(t[o])
    obj = np.concatenate(obj_l)
    tile = np.concatenate(tile_l)
    m = obj.size
    b = _corners(cx[obj] + rng.normal(0, 1.5, m) - tx0[tile], cy[obj] + rng.normal(0, 1.5, m) - ty0[tile],
                 w[obj] * rng.uniform(0.97, 1.03, m), h[obj] * rng.uniform(0.97, 1.03, m),
                 th[obj] + np.deg2rad(rng.uniform(-2, 2, m)))
    cf = np.clip(conf[obj] + rng.uniform(-0.05, 0.05, m), 0.01, 0.999).astype(np.float32)
    order = np.lexsort((-cf.astype(np.float64), tile))
    return (b[order].astype(np.float32), cls[obj][order].astype(np.int32), cf[order],
            tile[order].astype(np.int32))
