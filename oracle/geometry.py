"""Float64 CPU restatement of the geometry half of the hot path.  TEST INFRASTRUCTURE ONLY.

Follows (file:line relative to the reference repository):
  * tile enumeration            Detect_OBB.py:210-223
  * tile->map remap             Detect_OBB.py:233-240
  * border filter               Detect_OBB.py:156-174, 242-249
  * strike angle                Detect_OBB.py:135-142, 251-254
  * rotated IoU                 Detect_OBB.py:144-154   (shapely 2.0.7 / GEOS, restated)
  * greedy class-wise NMS       Detect_OBB.py:176-200
  * dual-scale late fusion      Detect_OBB.py:347-423

The rotated IoU arithmetic lives in shapely==2.0.7 (requirements.txt:4), which is absent
from /root/reference and not installable offline: its published behaviour is restated
(``Polygon.is_valid`` -> simple ring with non-zero area, convex or concave;
``intersection().area`` of two convex quads -> Sutherland-Hodgman clip + shoelace; a concave
simple quad is split at its reflex vertex into two triangles and the clips of the piece
pairs are summed; all float64).  PARITY UNPINNED for the
IoU value itself; the keep-sets are pinned through Output/Test{1,2}.xlsx (tests/).

Pure-Python loops: use on small cases only; ``oracle/geom_c.c`` holds the same
algorithms in C for the 10^5-box cases.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

# ----------------------------------------------------------------------------- tiles


def tile_plan(H: int, W: int, tile_size: int, overlap: int) -> List[Tuple[int, int, int, int]]:
    """Row-major list of (y0, x0, h, w); ragged edge tiles are kept (Detect_OBB.py:210-223)."""
    step = max(1, tile_size - overlap)
    out = []
    for y0 in range(0, H, step):
        for x0 in range(0, W, step):
            h = min(y0 + tile_size, H) - y0
            w = min(x0 + tile_size, W) - x0
            if h == 0 or w == 0:
                continue
            out.append((y0, x0, h, w))
    return out


def margin_for(tile_size: int, margin_small: int = 10, margin_large: int = 20) -> int:
    """Detect_OBB.py:156-157 (MARGIN_128 / MARGIN_416 at :39-40)."""
    return margin_small if tile_size <= 128 else margin_large


def center_in_safe_region(pts8: Sequence[float], x0, y0, w, h, margin) -> bool:
    """Closed-interval test on the mean of the four corners (Detect_OBB.py:159-174)."""
    cx = (pts8[0] + pts8[2] + pts8[4] + pts8[6]) / 4.0
    cy = (pts8[1] + pts8[3] + pts8[5] + pts8[7]) / 4.0
    rx = cx - x0
    ry = cy - y0
    return (margin <= rx <= (w - margin)) and (margin <= ry <= (h - margin))


def strike_angle(pts8: Sequence[float]) -> float:
    """Detect_OBB.py:135-142: direction of edge pt1->pt4, folded to [0, 180]."""
    a = math.atan2(pts8[6] - pts8[0], pts8[7] - pts8[1]) * (180.0 / math.pi)
    return 180.0 - a if a > 0 else abs(a)


# ----------------------------------------------------------------------------- polygons


def _shoelace2(p) -> float:
    """Twice the signed area."""
    s = 0.0
    n = len(p)
    for i in range(n):
        x1, y1 = p[i]
        x2, y2 = p[(i + 1) % n]
        s += x1 * y2 - x2 * y1
    return s


def quad_is_convex_valid(p) -> bool:
    """True iff the 4-point ring is a non-degenerate convex quad (either winding).

    The fast path of every IoU (the detector emits rectangles).  Concave simple quads - valid
    for shapely - are handled by :func:`quad_classify` / :func:`_quad_iou_general`.
    """
    if _shoelace2(p) == 0.0:
        return False
    pos = neg = False
    for i in range(4):
        ax, ay = p[i]
        bx, by = p[(i + 1) % 4]
        cx, cy = p[(i + 2) % 4]
        cr = (bx - ax) * (cy - by) - (by - ay) * (cx - bx)
        if cr > 0:
            pos = True
        elif cr < 0:
            neg = True
    return not (pos and neg)


def _clip_convex(subject, clipper):
    """Sutherland-Hodgman: ``subject`` clipped by CCW convex ``clipper`` (inclusive side test)."""
    out = list(subject)
    n = len(clipper)
    for i in range(n):
        if not out:
            break
        ax, ay = clipper[i]
        bx, by = clipper[(i + 1) % n]
        ex, ey = bx - ax, by - ay
        inp, out = out, []
        m = len(inp)
        for k in range(m):
            px, py = inp[k]
            qx, qy = inp[(k + 1) % m]
            dp = ex * (py - ay) - ey * (px - ax)
            dq = ex * (qy - ay) - ey * (qx - ax)
            if dp >= 0:
                out.append((px, py))
                if dq < 0:
                    t = dp / (dp - dq)
                    out.append((px + t * (qx - px), py + t * (qy - py)))
            elif dq >= 0:
                t = dp / (dp - dq)
                out.append((px + t * (qx - px), py + t * (qy - py)))
    return out


def quad_iou(b1: Sequence[float], b2: Sequence[float]) -> float:
    """Rotated IoU of two quads given as 8 floats (Detect_OBB.py:144-154)."""
    p1 = [(float(b1[i]), float(b1[i + 1])) for i in range(0, 8, 2)]
    p2 = [(float(b2[i]), float(b2[i + 1])) for i in range(0, 8, 2)]
    if not quad_is_convex_valid(p1) or not quad_is_convex_valid(p2):
        return _quad_iou_general(p1, p2)     # concave simple quads are valid for shapely
    s1 = _shoelace2(p1)
    s2 = _shoelace2(p2)
    if s1 < 0:
        p1 = p1[::-1]
    if s2 < 0:
        p2 = p2[::-1]
    # pair-local origin (first corner of box 1): removes the map-scale offset from the
    # cross products; float64 does not need it but the CUDA kernel does and both then clip
    # the same translated polygons.
    ox, oy = p1[0]
    q1 = [(x - ox, y - oy) for x, y in p1]
    q2 = [(x - ox, y - oy) for x, y in p2]
    poly = _clip_convex(q1, q2)
    inter = abs(_shoelace2(poly)) * 0.5 if len(poly) >= 3 else 0.0
    a1 = abs(s1) * 0.5
    a2 = abs(s2) * 0.5
    union = a1 + a2 - inter
    return inter / union if union > 0 else 0.0


def _orient(a, b, c) -> float:
    return (b[0] - a[0]) * (c[1] - a[1]) - (b[1] - a[1]) * (c[0] - a[0])


def _on_seg(a, b, c) -> bool:
    return min(a[0], b[0]) <= c[0] <= max(a[0], b[0]) and min(a[1], b[1]) <= c[1] <= max(a[1], b[1])


def _segs_meet(a, b, c, d) -> bool:
    """Closed segments a-b and c-d share a point."""
    d1, d2, d3, d4 = _orient(c, d, a), _orient(c, d, b), _orient(a, b, c), _orient(a, b, d)
    if ((d1 > 0 and d2 < 0) or (d1 < 0 and d2 > 0)) and ((d3 > 0 and d4 < 0) or (d3 < 0 and d4 > 0)):
        return True
    return ((d1 == 0 and _on_seg(c, d, a)) or (d2 == 0 and _on_seg(c, d, b)) or
            (d3 == 0 and _on_seg(a, b, c)) or (d4 == 0 and _on_seg(a, b, d)))


def quad_classify(p):
    """(kind, ccw ring, reflex index): kind 0 invalid (zero area, bow-tie, spike, self-touching ring), 1 convex,
    2 concave simple - shapely's ``is_valid`` is True for kinds 1 and 2 (Detect_OBB.py:148-151)."""
    s = _shoelace2(p)
    if len(p) != 4 or s == 0.0:
        return 0, list(p), 0
    q = list(p) if s > 0 else [p[0], p[3], p[2], p[1]]
    neg = [i for i in range(4) if _orient(q[i - 1], q[i], q[(i + 1) % 4]) < 0]
    if not neg:
        return 1, q, 0
    if _segs_meet(q[0], q[1], q[2], q[3]) or _segs_meet(q[1], q[2], q[3], q[0]) or len(neg) != 1:
        return 0, q, 0
    return 2, q, neg[0]


def quad_is_valid(p) -> bool:
    return quad_classify(p)[0] != 0


def _pieces(kind, q, r):
    """Counter-clockwise convex pieces: the quad itself, or the two triangles at the reflex vertex."""
    if kind == 1:
        return [q]
    a, b, c, d = (q[(r + i) % 4] for i in range(4))
    return [[a, b, c], [a, c, d]]


def _quad_iou_general(p1, p2) -> float:
    if not quad_is_valid(p1) or not quad_is_valid(p2):
        return 0.0
    inter = _general_inter_area(p1, p2)
    a1, a2 = abs(_shoelace2(p1)) * 0.5, abs(_shoelace2(p2)) * 0.5
    union = a1 + a2 - inter
    return inter / union if union > 0 else 0.0


class Polygon:
    """Minimal stand-in for ``shapely.geometry.Polygon`` (only what Detect_OBB.py uses)."""

    def __init__(self, pts):
        self.pts = [(float(x), float(y)) for x, y in pts]

    @property
    def is_valid(self) -> bool:
        return len(self.pts) == 4 and quad_is_valid(self.pts) or (
            len(self.pts) == 3 and _shoelace2(self.pts) != 0.0)

    @property
    def area(self) -> float:
        return abs(_shoelace2(self.pts)) * 0.5 if len(self.pts) >= 3 else 0.0

    def intersection(self, other: "Polygon") -> "Polygon":
        if len(self.pts) == 4 and len(other.pts) == 4 and not (quad_is_convex_valid(self.pts) and quad_is_convex_valid(other.pts)):
            return _AreaOnly(_general_inter_area(self.pts, other.pts))
        a = self.pts if _shoelace2(self.pts) >= 0 else self.pts[::-1]
        b = other.pts if _shoelace2(other.pts) >= 0 else other.pts[::-1]
        ox, oy = a[0]
        a = [(x - ox, y - oy) for x, y in a]
        b = [(x - ox, y - oy) for x, y in b]
        return Polygon(_clip_convex(a, b))

    def contains(self, pt: "Point") -> bool:
        if len(self.pts) == 4 and not quad_is_convex_valid(self.pts):
            kind, q, r = quad_classify(self.pts)
            if kind != 2:
                return False
            c = (pt.x, pt.y)
            if any(_orient(q[i], q[(i + 1) % 4], c) == 0 and _on_seg(q[i], q[(i + 1) % 4], c) for i in range(4)):
                return False
            return any(all(_orient(t[i], t[(i + 1) % 3], c) >= 0 for i in range(3)) for t in _pieces(kind, q, r))
        p = self.pts if _shoelace2(self.pts) >= 0 else self.pts[::-1]
        n = len(p)
        for i in range(n):
            ax, ay = p[i]
            bx, by = p[(i + 1) % n]
            if (bx - ax) * (pt.y - ay) - (by - ay) * (pt.x - ax) <= 0:
                return False
        return True


class _AreaOnly:
    """Result of an intersection that involved a concave quad: only ``.area`` is defined."""

    def __init__(self, area: float):
        self.area = area


def _general_inter_area(p1, p2) -> float:
    k1, q1, r1 = quad_classify(p1)
    k2, q2, r2 = quad_classify(p2)
    if k1 == 0 or k2 == 0:
        return 0.0
    ox = 0.25 * ((p1[0][0] + p1[1][0]) + (p1[2][0] + p1[3][0]))
    oy = 0.25 * ((p1[0][1] + p1[1][1]) + (p1[2][1] + p1[3][1]))
    inter = 0.0
    for a in _pieces(k1, q1, r1):
        for b in _pieces(k2, q2, r2):
            poly = _clip_convex([(x - ox, y - oy) for x, y in a], [(x - ox, y - oy) for x, y in b])
            inter += abs(_shoelace2(poly)) * 0.5 if len(poly) >= 3 else 0.0
    return inter


class Point:
    def __init__(self, x, y):
        self.x, self.y = float(x), float(y)


# ----------------------------------------------------------------------------- NMS / fusion


def merge_detections(dets: list, iou_threshold: float = 0.5) -> list:
    """Greedy class-wise rotated NMS (Detect_OBB.py:176-200).

    Sorts ``dets`` IN PLACE (stable, confidence descending) like the reference and returns
    the kept members in that order.  Each det is (x1..y4, cls, conf, angle).
    """
    if not dets:
        return []
    dets.sort(key=lambda d: d[9], reverse=True)
    kept: list = []
    for d in dets:
        ok = True
        for k in kept:
            if k[8] == d[8] and quad_iou(d[:8], k[:8]) >= iou_threshold:
                ok = False
                break
        if ok:
            kept.append(d)
    return kept


def nms_keep_indices(boxes, cls, conf, iou_threshold: float) -> List[int]:
    """Index form of :func:`merge_detections`: kept input indices in output order."""
    order = sorted(range(len(conf)), key=lambda i: -float(conf[i]))  # stable
    kept: List[int] = []
    for i in order:
        ok = True
        for k in kept:
            if cls[k] == cls[i] and quad_iou(boxes[i], boxes[k]) >= iou_threshold:
                ok = False
                break
        if ok:
            kept.append(i)
    return kept


CONS_IOU_PARTNER = 0.40  # Detect_OBB.py:349
CONS_LOW = 0.25          # Detect_OBB.py:350
CONS_HIGH = 0.70         # Detect_OBB.py:351


def cross_scale_consensus_filter(dets_by_scale: Dict[int, list]) -> list:
    """Dual-scale late fusion (Detect_OBB.py:347-423), any number of scales.

    Processing order: scales ascending, list order inside a scale.  An unvisited det looks
    for the best unvisited same-class partner in the *other* scales with IoU >= 0.40
    (highest conf, then highest IoU, then first met); with a partner the stronger of the
    two is kept and both are retired; without one it is kept iff conf >= 0.70.
    """
    scales = sorted(dets_by_scale)
    if len(scales) == 1:
        return list(dets_by_scale[scales[0]])
    pools = {s: [d for d in dets_by_scale[s] if d[9] >= CONS_LOW] for s in scales}
    seen = {s: [False] * len(pools[s]) for s in scales}
    kept = []
    for s in scales:
        for i, d in enumerate(pools[s]):
            if seen[s][i]:
                continue
            best = None
            best_conf, best_iou = -1.0, 0.0
            for t in scales:
                if t == s:
                    continue
                for j, p in enumerate(pools[t]):
                    if seen[t][j] or int(p[8]) != int(d[8]):
                        continue
                    v = quad_iou(d[:8], p[:8])
                    if v >= CONS_IOU_PARTNER:
                        cp = float(p[9])
                        if cp > best_conf or (cp == best_conf and v > best_iou):
                            best, best_conf, best_iou = (t, j, p), cp, v
            seen[s][i] = True
            if best is None or best_conf < CONS_LOW:
                if float(d[9]) >= CONS_HIGH:
                    kept.append(d)
                continue
            t, j, p = best
            kept.append(d if float(d[9]) >= best_conf else p)
            seen[t][j] = True
    return kept
