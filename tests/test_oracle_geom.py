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


def test_iou_oracle_against_two_independent_libraries():
    """shapely / GEOS (what the reference calls, Detect_OBB.py:148-154) cannot be installed here, so the value of the
    float64 IoU restatement is cross-checked against two independent third-party polygon-intersection implementations
    that ARE in the image: Qhull (scipy HalfspaceIntersection + ConvexHull area, float64) and OpenCV's
    intersectConvexConvex (fp32).  Neither shares code with oracle/geometry.py."""
    import cv2
    from scipy.optimize import linprog
    from scipy.spatial import ConvexHull, HalfspaceIntersection

    def rb(cx, cy, w, h, th):
        c, s = np.cos(th), np.sin(th)
        v1 = np.array([w / 2 * c, w / 2 * s]); v2 = np.array([-h / 2 * s, h / 2 * c]); ctr = np.array([cx, cy])
        return np.concatenate([ctr + v1 + v2, ctr + v1 - v2, ctr - v1 - v2, ctr - v1 + v2])

    def halfspaces(q):
        P = q.reshape(4, 2)
        a = 0.5 * sum(P[i, 0] * P[(i + 1) % 4, 1] - P[(i + 1) % 4, 0] * P[i, 1] for i in range(4))
        if a < 0:
            P = P[::-1]
        hs = []
        for i in range(4):
            p, e = P[i], P[(i + 1) % 4] - P[i]
            n = np.array([e[1], -e[0]])                      # outward normal of a CCW edge: n.x + b <= 0 inside
            hs.append([n[0], n[1], -(n @ p)])
        return np.array(hs), abs(a)

    def qhull_iou(A, B):
        ha, aa = halfspaces(A)
        hb, ab = halfspaces(B)
        H = np.vstack([ha, hb])
        norm = np.linalg.norm(H[:, :2], axis=1)
        res = linprog([0, 0, -1], A_ub=np.hstack([H[:, :2], norm[:, None]]), b_ub=-H[:, 2],
                      bounds=[(None, None), (None, None), (0, None)])           # Chebyshev centre = an interior point
        if res.status != 0 or res.x[2] <= 1e-9:
            return 0.0, aa, ab
        inter = ConvexHull(HalfspaceIntersection(H, res.x[:2]).intersections).volume
        return inter / (aa + ab - inter), aa, ab

    rng = np.random.default_rng(0)
    overlapping = 0
    for k in range(400):
        cx, cy = rng.uniform(0, 16000, 2)
        w, h = rng.uniform(12, 100, 2)
        th = rng.uniform(-1, 2)
        A = rb(cx, cy, w, h, th)
        if k % 4 == 3:          # general convex quads (GT labels are parallelograms, evaluation path)
            ang = np.sort(rng.uniform(0, 2 * np.pi, 4)); r = rng.uniform(10, 60, 4)
            B = (np.array([cx, cy]) + rng.normal(0, 10, 2) + np.stack([r * np.cos(ang), r * np.sin(ang)], 1)).ravel()
            if not G.quad_is_convex_valid([tuple(v) for v in B.reshape(4, 2)]):   # concave quads are reported invalid by the oracle
                continue
        else:
            B = rb(cx + rng.normal(0, 15), cy + rng.normal(0, 15), w * rng.uniform(.6, 1.4), h * rng.uniform(.6, 1.4),
                   th + rng.normal(0, .4))
        v = G.quad_iou(A, B)
        vq, aa, ab = qhull_iou(A, B)
        assert abs(v - vq) < 1e-9, (k, v, vq)
        # OpenCV works in fp32: compare in a pair-local frame, as the CUDA path does
        o = A.reshape(4, 2).mean(0)
        ic, _ = cv2.intersectConvexConvex((A.reshape(4, 2) - o).astype(np.float32), (B.reshape(4, 2) - o).astype(np.float32))
        assert abs(v - ic / (aa + ab - ic)) < 2e-5, (k, v)
        overlapping += v > 0.05
    assert overlapping > 200
