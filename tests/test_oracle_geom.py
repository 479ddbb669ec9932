"""The float64 geometry restatement against the reference's xlsx outputs and against the lifted
reference's merge / fusion / detect_symbols runs (tests/golden/merge_golden.json)."""
import numpy as np
import pytest

from oracle import geometry as G


def _rows(xlsx_rows, name):
    return [r for r in xlsx_rows[name]["rows"]]


def test_strike_angles_reproduce_xlsx(xlsx_rows):
    n = 0
    for name in ("Test1", "Test2"):
        for r in _rows(xlsx_rows, name):
            pts = r[1:9]
            if r[0] == "Strike":
                assert abs(G.strike_angle(pts) - r[10]) < 1e-9
                n += 1
            else:
                assert r[10] == 0
    assert n == 35


def test_xlsx_is_an_nms_fixed_point_and_sorted(xlsx_rows):
    names = sorted({r[0] for nm in ("Test1", "Test2") for r in _rows(xlsx_rows, nm)})
    for name in ("Test1", "Test2"):
        dets = [tuple(r[1:9]) + (names.index(r[0]), r[9], r[10]) for r in _rows(xlsx_rows, name)]
        confs = [d[9] for d in dets]
        assert confs == sorted(confs, reverse=True)
        work = list(dets)
        kept = G.merge_detections(work, 0.4)
        assert kept == dets and work == dets
        worst = max((G.quad_iou(a[:8], b[:8]) for i, a in enumerate(dets) for b in dets[:i] if a[8] == b[8]),
                    default=0.0)
        assert worst < 0.4
    assert len(_rows(xlsx_rows, "Test1")) == 34 and len(_rows(xlsx_rows, "Test2")) == 10


@pytest.mark.parametrize("case", ["sparse", "dense", "mapscale"])
def test_merge_matches_lifted_reference(merge_golden, case):
    g = merge_golden["merge"][case]
    dets = [tuple(d) for d in g["dets"]]
    work = list(dets)
    kept = G.merge_detections(work, 0.4)
    pos = {id(d): i for i, d in enumerate(dets)}
    assert [pos[id(d)] for d in work] == g["sorted"]
    assert [pos[id(d)] for d in kept] == g["kept"]
    boxes = np.array([d[:8] for d in dets]); cls = np.array([d[8] for d in dets]); conf = np.array([d[9] for d in dets])
    assert G.nms_keep_indices(boxes, cls, conf, 0.4) == g["kept"]


@pytest.mark.parametrize("case", ["two", "three"])
def test_fusion_matches_lifted_reference(merge_golden, case):
    g = merge_golden["fusion"][case]
    by_scale = {int(s): [tuple(d) for d in v] for s, v in g["by_scale"].items()}
    flat = [d for s in sorted(by_scale) for d in by_scale[s]]
    pos = {id(d): i for i, d in enumerate(flat)}
    kept = G.cross_scale_consensus_filter(by_scale)
    assert [pos[id(d)] for d in kept] == g["kept"]
    one = {416: flat[:10]}
    assert G.cross_scale_consensus_filter(one) == flat[:10]


def test_detect_symbols_golden_replays_through_oracle(merge_golden):
    g = merge_golden["detect_symbols"]
    plan = G.tile_plan(g["H"], g["W"], g["tile"], g["overlap"])
    assert [list((h, w, 3)) for (_, _, h, w) in plan] == g["calls"]
    out = []
    margin = G.margin_for(g["tile"])
    for (y0, x0, h, w), rec in zip(plan, g["per_tile"]):
        dets = []
        for pts, c, f in zip(rec["corners"], rec["cls"], rec["conf"]):
            gp = [pts[k] + (x0 if k % 2 == 0 else y0) for k in range(8)]
            if not G.center_in_safe_region(gp, x0, y0, w, h, margin):
                continue
            ang = G.strike_angle(pts) if c == 1 else 0.0
            dets.append(tuple(gp) + (c, f, ang))
        out.extend(G.merge_detections(dets, 0.4))
    want = [tuple(d) for d in g["out"]]
    assert len(out) == len(want)
    for a, b in zip(out, want):
        assert a[8] == b[8] and a[9] == b[9]
        assert np.allclose(a[:8], b[:8], rtol=0, atol=0) and abs(a[10] - b[10]) < 1e-9


def test_iou_edge_cases():
    A = [0, 0, 10, 0, 10, 10, 0, 10]
    assert G.quad_iou(A, A) == 1.0
    assert G.quad_iou(A, A[::-1][1:] + A[::-1][:1]) in (0.0, 1.0)    # reversed list is still a valid ring
    assert G.quad_iou(A, [10, 0, 20, 0, 20, 10, 10, 10]) == 0.0         # touching along an edge
    assert G.quad_iou(A, [0, 0, 20, 0, 20, 10, 0, 10]) == 0.5
    assert G.quad_iou(A, [0, 0, 10, 10, 10, 0, 0, 10]) == 0.0           # bow-tie: invalid
    assert G.quad_iou(A, [0] * 8) == 0.0                                # collapsed: invalid
    assert abs(G.quad_iou(A, [5, 5, 15, 5, 15, 15, 5, 15]) - 25 / 175) < 1e-15


def test_grid_indexed_oracle_equals_plain_loops():
    """oracle/geom_c.c's grid variants (used for the 10^6-box GPU parity cases) are the plain sequential
    loops with a different candidate lookup: same order, same kept sets, ties included."""
    from oracle import geom_c
    from oriented_object_detection_b200 import synth
    for seed, (n, ext, nc) in enumerate([(3000, 1500, 5), (2000, 400, 2), (5000, 6000, 15), (50, 100, 1)]):
        boxes, cls, conf = synth.synthetic_obbs(n, ext, ext, n_classes=nc, seed=seed)
        conf[::9] = conf[4]
        o1, k1 = geom_c.nms(boxes, cls, conf, 0.4)
        o2, k2 = geom_c.nms(boxes, cls, conf, 0.4, grid=True)
        assert np.array_equal(o1, o2) and np.array_equal(k1, k2)
        rng = np.random.default_rng(seed)
        ns = 2 + seed % 2
        sid = np.sort(rng.integers(0, ns, len(conf))).astype(np.int32)
        assert np.array_equal(geom_c.fuse(boxes, cls, conf, sid, ns), geom_c.fuse(boxes, cls, conf, sid, ns, grid=True))
