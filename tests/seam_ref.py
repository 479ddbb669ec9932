"""Test-side reference pieces for the seam-band exchange (tests only; geometry from oracle/).

  * make_case: per-tile survivors of a small map in list order (tile row-major, confidence-descending inside a tile,
    centres inside the safe region of their tile - what detect_symbols returns, Detect_OBB.py:242-264) plus adversarial
    same-class chains that run across tile seams and across whole bands;
  * deferring_nms / plain_nms: the sequential greedy rule with and without deferral (float64 IoU of the oracle);
  * expected: merge_detections over the whole list (Detect_OBB.py:176-200).
"""
import numpy as np

from oracle import geom_c
from oracle import geometry as G


def _corners(cx, cy, w, h, th):
    c, s = np.cos(th), np.sin(th)
    v1 = np.array([w / 2 * c, w / 2 * s]); v2 = np.array([-h / 2 * s, h / 2 * c]); ctr = np.array([cx, cy])
    return np.concatenate([ctr + v1 + v2, ctr + v1 - v2, ctr - v1 - v2, ctr - v1 + v2])


def make_case(H, W, tile, overlap, margin, n_objects, n_classes, seed, chains=6, chain_len=40):
    """-> boxes float64 [n,8] (fp32 values), cls int32, conf float32, tile_id int32 (non-decreasing)."""
    rng = np.random.default_rng(seed)
    plan = G.tile_plan(H, W, tile, overlap)
    rows = []

    def tiles_of(cx, cy):
        return [t for t, (y0, x0, h, w) in enumerate(plan)
                if margin <= cx - x0 <= w - margin and margin <= cy - y0 <= h - margin]

    def emit(cx, cy, w, h, th, cls, conf, every_tile=True):
        ts = tiles_of(cx, cy)
        if not ts:
            return
        if not every_tile:
            ts = [ts[int(rng.integers(len(ts)))]]
        for t in ts:
            b = _corners(cx + rng.normal(0, 1.0), cy + rng.normal(0, 1.0), w * rng.uniform(0.98, 1.02), h * rng.uniform(0.98, 1.02),
                         th + rng.normal(0, 0.02)).astype(np.float32).astype(np.float64)
            y0, x0, hh, ww = plan[t]
            if G.center_in_safe_region(b, x0, y0, ww, hh, margin):
                rows.append((t, b, cls, np.float32(np.clip(conf + rng.uniform(-0.04, 0.04), 0.01, 0.999))))

    for _ in range(n_objects):
        emit(rng.uniform(0, W), rng.uniform(0, H), rng.uniform(12, 70), rng.uniform(11, 60), rng.uniform(-0.8, 2.3),
             int(rng.integers(n_classes)), rng.uniform(0.25, 1.0))
    # chains: same class, each box overlaps the next (IoU ~ 0.5), running down the map across every seam; confidences
    # ascending, descending, alternating or equal, so verdicts propagate along the chain in both directions
    for k in range(chains):
        cx = rng.uniform(margin + 40, W - margin - 40)
        w, h = rng.uniform(40, 70), rng.uniform(30, 50)
        cls = int(rng.integers(n_classes))
        y = rng.uniform(margin + 5, margin + 60)
        base = rng.uniform(0.3, 0.9)
        for i in range(chain_len):
            mode = k % 4
            conf = (base + 0.002 * i if mode == 0 else base - 0.002 * i if mode == 1 else
                    base + (0.05 if i % 2 else -0.05) if mode == 2 else base)
            emit(cx + rng.normal(0, 0.5), y, w, h, 0.0, cls, conf, every_tile=(i % 3 == 0))
            y += 0.32 * h
            if y > H - margin - 5:
                break
    tid = np.array([r[0] for r in rows]); conf = np.array([r[3] for r in rows], dtype=np.float32)
    order = np.lexsort((-conf.astype(np.float64), tid))           # tile-major, confidence-descending inside a tile (stable)
    boxes = np.array([rows[i][1] for i in order]); cls = np.array([rows[i][2] for i in order], dtype=np.int32)
    return boxes, cls, conf[order], tid[order].astype(np.int32), plan


def _aabb(b):
    return b[:, 0::2].min(1), b[:, 1::2].min(1), b[:, 0::2].max(1), b[:, 1::2].max(1)


def deferring_nms(boxes, cls, conf, cand, thr):
    """Sequential greedy NMS with deferral: -> (stable confidence-descending order, state 1 kept / 2 suppressed / 3 deferred)."""
    boxes, cls, conf, cand = (np.asarray(a) for a in (boxes, cls, conf, cand))
    n = len(conf)
    order = np.argsort(-conf.astype(np.float64), kind="stable")
    state = np.zeros(n, np.uint8)
    x0, y0, x1, y1 = _aabb(boxes)
    done = []
    for i in order:
        if cls[i] < 0:
            state[i] = 2
            continue
        sup, dfr = False, bool(cand[i])
        for j in done:
            if cls[j] != cls[i] or state[j] == 2:
                continue
            if x0[i] > x1[j] or x0[j] > x1[i] or y0[i] > y1[j] or y0[j] > y1[i]:
                continue
            if geom_c.quad_iou(boxes[i], boxes[j]) >= thr:
                if state[j] == 1:
                    sup = True
                    break
                dfr = True                     # state[j] == 3
        state[i] = 2 if sup else (3 if dfr else 1)
        done.append(i)
    return order, state


def plain_nms(boxes, cls, conf, thr):
    """-> (stable confidence-descending order, keep uint8 by row); rows with class < 0 are never kept."""
    boxes, cls, conf = (np.asarray(a) for a in (boxes, cls, conf))
    live = np.nonzero(cls >= 0)[0]
    keep = np.zeros(len(conf), np.uint8)
    if live.size:
        _, k = geom_c.nms(np.ascontiguousarray(boxes[live]), np.ascontiguousarray(cls[live]), np.ascontiguousarray(conf[live]), thr)
        keep[live[k]] = 1
    return np.argsort(-conf.astype(np.float64), kind="stable"), keep


def expected(boxes, cls, conf, thr):
    """Kept input indices of merge_detections over the whole list, in output order."""
    return geom_c.nms(boxes, cls, conf, thr)[1]


def merge_rank_outputs(outs):
    """[(conf, global row index) per rank in the rank's output order] -> global rows in the reference's order:
    stable confidence-descending over the concatenation in rank order."""
    conf = np.concatenate([o[0] for o in outs]); idx = np.concatenate([o[1] for o in outs])
    return idx[np.argsort(-conf.astype(np.float64), kind="stable")]
