"""numpy (float32) restatement of the Ultralytics 8.3.196 OBB predictor tail.  TEST INFRASTRUCTURE ONLY.

Reference call sites: Detect_OBB.py:81-83 (``model(net_input, conf=...)``) and :228-231
(``results[0].obb[i].xyxyxyxy/.cls/.conf``).  The arithmetic is third-party
(ultralytics==8.3.196, requirements.txt:3) - absent from /root/reference and not installable
offline - so its published algorithm is restated: ``non_max_suppression(
rotated=True)`` (best-class confidence filter, confidence-descending order, class offset,
``batch_probiou`` fast-NMS where suppressed boxes still suppress, ``max_det``),
``regularize_rboxes``, ``scale_boxes(xywh=True)``, ``xywhr2xyxyxyxy``.  PARITY UNPINNED: there is
no reference test, golden vector or runnable upstream for it; the CUDA kernel is checked against
this restatement only.

Letterbox convention (shared with csrc/decode.cu): gain = min(net_h/h, net_w/w), tile centred in the
net_h x net_w input (oracle/letterbox.py gives that shape for Ultralytics' rect letterbox).
"""
from __future__ import annotations

import numpy as np

F = np.float32
EPS = F(1e-7)


def _cov(w, h, th):
    a = w * w / F(12.0)
    b = h * h / F(12.0)
    c, s = np.cos(th), np.sin(th)
    return a * c * c + b * s * s, a * s * s + b * c * c, (a - b) * c * s


def probiou(b1, b2) -> np.float32:
    """b = (cx, cy, w, h, theta) float32."""
    x1, y1, x2, y2 = b1[0], b1[1], b2[0], b2[1]
    a1, bb1, c1 = _cov(b1[2], b1[3], b1[4])
    a2, bb2, c2 = _cov(b2[2], b2[3], b2[4])
    sa, sb, sc = a1 + a2, bb1 + bb2, c1 + c2
    den = sa * sb - sc * sc
    dx, dy = x1 - x2, y1 - y2
    t1 = ((sa * dy * dy + sb * dx * dx) / (den + EPS)) * F(0.25)
    t2 = ((sc * (x2 - x1) * (y1 - y2)) / (den + EPS)) * F(0.5)
    d1 = np.maximum(a1 * bb1 - c1 * c1, F(0))
    d2 = np.maximum(a2 * bb2 - c2 * c2, F(0))
    t3 = np.log(den / (F(4) * np.sqrt(d1 * d2) + EPS) + EPS) * F(0.5)
    bd = np.clip(t1 + t2 + t3, EPS, F(100.0))
    hd = np.sqrt(F(1.0) - np.exp(-bd) + EPS)
    return F(1.0) - hd


def decode_tile(head: np.ndarray, tile_h: int, tile_w: int, net_size, conf_thr: float = 0.25,
                iou_thr: float = 0.7, max_det: int = 300):
    """head float32 [4+nc+1, A] -> (corners float32 [k,8] tile-local, cls int [k], conf float32 [k]).
    ``net_size``: side of a square network input or (net_h, net_w) of a rect-letterboxed one
    (``scale_boxes``: gain = min(net_h / h, net_w / w), pad = round((net - dim * gain) / 2 - 0.1))."""
    net_h, net_w = (net_size, net_size) if np.isscalar(net_size) else net_size
    head = head.astype(F)
    nc = head.shape[0] - 5
    scores = head[4:4 + nc]
    cls = scores.argmax(0)                       # first maximum, like the kernel's strict '>' scan
    conf = scores.max(0)
    cand = np.nonzero(conf > F(conf_thr))[0]
    cand = cand[np.lexsort((cand, -conf[cand].astype(np.float64)))]      # conf desc, anchor asc on ties
    boxes = np.stack([head[0, cand], head[1, cand], head[2, cand], head[3, cand], head[4 + nc, cand]], axis=1)
    ccls, cconf = cls[cand], conf[cand]
    live = []
    for j in range(len(cand)):
        dead = False
        for i in range(j):
            if ccls[i] == ccls[j] and probiou(boxes[i], boxes[j]) >= F(iou_thr):
                dead = True
                break
        if not dead:
            live.append(j)
        if len(live) == max_det:
            break
    out_b, out_c, out_f = [], [], []
    gain64 = min(net_h / tile_h, net_w / tile_w)
    gain = F(gain64)
    padx = F(round((net_w - tile_w * gain64) / 2 - 0.1))
    pady = F(round((net_h - tile_h * gain64) / 2 - 0.1))
    PI = F(np.pi)
    for j in live:
        cx, cy, w, h, th = boxes[j]
        tm = np.fmod(th, PI)
        if tm < 0:
            tm = tm + PI
        swap = tm >= PI / F(2)
        w_, h_ = (h, w) if swap else (w, h)
        tr = np.fmod(tm, PI / F(2))
        cx, cy = (cx - padx) / gain, (cy - pady) / gain
        bw, bh = w_ / gain, h_ / gain
        c, s = np.cos(tr), np.sin(tr)
        v1x, v1y = bw / F(2) * c, bw / F(2) * s
        v2x, v2y = -bh / F(2) * s, bh / F(2) * c
        out_b.append([cx + v1x + v2x, cy + v1y + v2y, cx + v1x - v2x, cy + v1y - v2y,
                      cx - v1x - v2x, cy - v1y - v2y, cx - v1x + v2x, cy - v1y + v2y])
        out_c.append(int(ccls[j]))
        out_f.append(cconf[j])
    return (np.asarray(out_b, dtype=F).reshape(-1, 8), np.asarray(out_c, dtype=np.int32), np.asarray(out_f, dtype=F))


def synthetic_head(n_tiles: int, nc: int, net_size: int, seed: int, density: float = 0.02) -> np.ndarray:
    """Raw head tensors with a controllable number of confident anchors, float32 [n_tiles, 4+nc+1, A]."""
    rng = np.random.default_rng(seed)
    strides = (8, 16, 32)
    axs, ays, sts = [], [], []
    for s in strides:
        g = net_size // s
        yy, xx = np.mgrid[0:g, 0:g]
        axs.append(((xx + 0.5) * s).ravel()); ays.append(((yy + 0.5) * s).ravel()); sts.append(np.full(g * g, s))
    ax, ay, st = np.concatenate(axs), np.concatenate(ays), np.concatenate(sts)
    A = ax.size
    head = np.zeros((n_tiles, 4 + nc + 1, A), dtype=F)
    for t in range(n_tiles):
        head[t, 0] = ax + rng.normal(0, 2, A)
        head[t, 1] = ay + rng.normal(0, 2, A)
        head[t, 2] = rng.uniform(1.5, 6, A) * st
        head[t, 3] = rng.uniform(1.0, 5, A) * st
        head[t, 4:4 + nc] = rng.uniform(0, 0.2, (nc, A))
        hot = rng.random(A) < density
        k = int(hot.sum())
        head[t, 4 + rng.integers(0, nc, k), np.nonzero(hot)[0]] = rng.uniform(0.2, 1.0, k)
        head[t, 4 + nc] = rng.uniform(-np.pi / 4, 3 * np.pi / 4, A)
        # a few exact duplicates of confident anchors: the fast-NMS must drop them
        src = np.nonzero(hot)[0][: max(1, k // 4)]
        dst = (src + 1) % A
        head[t, :, dst] = head[t, :, src]
        head[t, 0, dst] += 0.5
    return head
