"""Property tests (hypothesis) of the host logic and of the oracle: the reference has no tests to copy, so the
invariants its code implies are stated here.  CPU only."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from oracle import geometry as G

COMMON = dict(deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])


# ---------------------------------------------------------------- a1: tile plan through the C ABI (host code)

@settings(max_examples=150, **COMMON)
@given(H=st.integers(1, 3000), W=st.integers(1, 3000), ts=st.integers(1, 600), ov=st.integers(0, 700))
def test_tile_plan_properties(built_lib, H, W, ts, ov):
    from oriented_object_detection_b200 import ops
    step = max(1, ts - ov)
    if (-(-H // step)) * (-(-W // step)) > 40000:
        return
    plan = ops.make_plan(H, W, ts, ov)
    tiles = [(int(t["y0"]), int(t["x0"]), int(t["h"]), int(t["w"])) for t in plan.tiles]
    assert tiles == G.tile_plan(H, W, ts, ov)                       # Detect_OBB.py:210-223, row-major, ragged tiles kept
    ys = sorted({t[0] for t in tiles}); xs = sorted({t[1] for t in tiles})
    assert ys == list(range(0, H, step)) and xs == list(range(0, W, step))
    assert plan.rows == len(ys) and plan.cols == len(xs) and plan.n == len(ys) * len(xs)
    for y0, x0, h, w in tiles:
        assert 1 <= h <= ts and 1 <= w <= ts and y0 + h <= H and x0 + w <= W
        assert h == min(ts, H - y0) and w == min(ts, W - x0)
    # the tiles cover the map: consecutive origins are never further apart than a tile
    if step <= ts:
        assert max(y0 + h for y0, _, h, _ in tiles) == H and max(x0 + w for _, x0, _, w in tiles) == W
    offs = np.concatenate([[0], np.cumsum([h * w for _, _, h, w in tiles])])
    assert plan.tiles["px_off"].tolist() == offs[:-1].tolist() and plan.total_px == int(offs[-1])


@settings(max_examples=60, **COMMON)
@given(H=st.integers(50, 4000), W=st.integers(50, 2000), world=st.integers(1, 8))
def test_row_bands_partition_any_plan(built_lib, H, W, world):
    from oriented_object_detection_b200 import ops, sharding
    full = ops.make_plan(H, W, 416, 100)
    got, rows = [], 0
    for rank in range(world):
        r0, r1 = sharding.band_rows(full.rows, world, rank)
        assert r0 == rows and r1 >= r0
        rows = r1
        band = ops.make_plan(H, W, 416, 100, r0, r1)
        y0, y1 = sharding.band_pixel_rows(H, 416, 100, r0, r1)
        if band.n:
            assert band.tiles["px_off"][0] == 0
            assert int(band.tiles["y0"].min()) == y0 and int((band.tiles["y0"] + band.tiles["h"]).max()) == y1
        got += [(int(t["y0"]), int(t["x0"]), int(t["h"]), int(t["w"])) for t in band.tiles]
    assert rows == full.rows
    assert got == [(int(t["y0"]), int(t["x0"]), int(t["h"]), int(t["w"])) for t in full.tiles]
    sizes = [sharding.band_rows(full.rows, world, r)[1] - sharding.band_rows(full.rows, world, r)[0] for r in range(world)]
    assert max(sizes) - min(sizes) <= 1


# ---------------------------------------------------------------- oracle geometry invariants

def _rbox(cx, cy, w, h, th):
    c, s = np.cos(th), np.sin(th)
    v1 = np.array([w / 2 * c, w / 2 * s]); v2 = np.array([-h / 2 * s, h / 2 * c]); ctr = np.array([cx, cy])
    return np.concatenate([ctr + v1 + v2, ctr + v1 - v2, ctr - v1 - v2, ctr - v1 + v2])


box_st = st.tuples(st.floats(0, 500), st.floats(0, 500), st.floats(5, 120), st.floats(5, 120), st.floats(-0.8, 2.4))


@settings(max_examples=200, **COMMON)
@given(a=box_st, b=box_st, shift=st.tuples(st.integers(-5000, 5000), st.integers(-5000, 5000)), roll=st.integers(0, 3),
       flip=st.booleans())
def test_iou_invariants(a, b, shift, roll, flip):
    A, B = _rbox(*a), _rbox(*b)
    v = G.quad_iou(A, B)
    assert 0.0 <= v <= 1.0 + 1e-12
    assert abs(G.quad_iou(B, A) - v) < 1e-12                                  # symmetric
    assert abs(G.quad_iou(A, A) - 1.0) < 1e-12
    off = np.tile(np.array(shift, dtype=np.float64), 4)
    assert abs(G.quad_iou(A + off, B + off) - v) < 1e-9                        # translation (tile offset) invariant
    B2 = np.roll(B.reshape(4, 2), roll, axis=0)
    if flip:
        B2 = B2[::-1]
    assert abs(G.quad_iou(A, B2.reshape(-1)) - v) < 1e-12                      # vertex labelling / orientation invariant
    area_a, area_b = a[2] * a[3], b[2] * b[3]
    assert v <= min(area_a, area_b) / max(area_a, area_b) + 1e-9              # IoU <= area ratio


def _dets(draw_boxes, classes, confs):
    return [tuple(float(x) for x in _rbox(*bx)) + (int(c), float(np.float32(cf)), 0.0) for bx, c, cf in zip(draw_boxes, classes, confs)]


dets_st = st.integers(0, 40).flatmap(lambda n: st.tuples(
    st.lists(st.tuples(st.floats(0, 200), st.floats(0, 200), st.floats(10, 80), st.floats(10, 80), st.floats(-0.8, 2.4)), min_size=n, max_size=n),
    st.lists(st.integers(0, 2), min_size=n, max_size=n),
    st.lists(st.floats(0.05, 0.999), min_size=n, max_size=n)))


@settings(max_examples=80, **COMMON)
@given(d=dets_st, thr=st.sampled_from([0.2, 0.4, 0.5]))
def test_merge_detections_invariants(d, thr):
    dets = _dets(*d)
    work = list(dets)
    kept = G.merge_detections(work, thr)
    assert all(work[i][9] >= work[i + 1][9] for i in range(len(work) - 1))     # sorts the caller's list in place (Detect_OBB.py:183)
    assert sorted(map(id, work)) == sorted(map(id, dets))
    ids = {id(x) for x in dets}
    assert all(id(k) in ids for k in kept)                                      # members of the input, identity preserved
    assert all(kept[i][9] >= kept[i + 1][9] for i in range(len(kept) - 1))
    for i, a in enumerate(kept):                                                # no two kept boxes of a class reach the threshold
        for b in kept[i + 1:]:
            assert a[8] != b[8] or G.quad_iou(a[:8], b[:8]) < thr
    kept_ids = {id(k) for k in kept}
    for x in work:                                                              # every dropped box has a kept, at-least-as-confident suppressor
        if id(x) not in kept_ids:
            assert any(k[8] == x[8] and k[9] >= x[9] and G.quad_iou(k[:8], x[:8]) >= thr for k in kept)
    again = G.merge_detections(list(kept), thr)                                 # fixed point
    assert [id(x) for x in again] == [id(x) for x in kept]
    per_class = []                                                              # classes are independent (Detect_OBB.py:193)
    for c in range(3):
        per_class += G.merge_detections([x for x in dets if x[8] == c], thr)
    assert sorted(map(id, per_class)) == sorted(map(id, kept))
    assert G.merge_detections([], thr) == []


@settings(max_examples=60, **COMMON)
@given(d1=dets_st, d2=dets_st)
def test_fusion_invariants(d1, d2):
    s128, s416 = _dets(*d1), _dets(*d2)
    single = G.cross_scale_consensus_filter({128: s128})
    assert single == s128 and single is not s128                                # one scale: a new list, no confidence filter
    out = G.cross_scale_consensus_filter({128: list(s128), 416: list(s416)})
    ids = {id(x) for x in s128} | {id(x) for x in s416}
    assert all(id(x) in ids for x in out) and len({id(x) for x in out}) == len(out)
    assert all(x[9] >= 0.25 for x in out)                                       # CONS_LOW
    out_ids = {id(x) for x in out}
    for x in s128 + s416:                                                       # a confident box is only ever dropped in favour of a partner
        if x[9] >= 0.70 and id(x) not in out_ids:
            other = s416 if any(x is y for y in s128) else s128
            assert any(o[8] == x[8] and G.quad_iou(o[:8], x[:8]) >= 0.40 for o in other)
    for x in s416:                                                              # leftover large-scale boxes survive only when confident
        if id(x) in out_ids and x[9] < 0.70:
            assert any(o[8] == x[8] and o[9] >= 0.25 and G.quad_iou(o[:8], x[:8]) >= 0.40 for o in s128)
