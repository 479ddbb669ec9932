"""GPU parity of rotated IoU, exact greedy NMS (global and per tile), tile post-processing and the
dual-scale fusion against the float64 oracle and the lifted-reference golden runs."""
import numpy as np
import pytest
import torch

from oracle import geometry as G

pytestmark = pytest.mark.gpu


def _t(a, dev, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    return t.to(dtype) if dtype is not None else t


def test_iou_pairs_matrix_and_checksum(cuda_dev):
    from oriented_object_detection_b200 import ops, synth
    boxes, cls, conf = synth.synthetic_obbs(700, 3000, 3000, seed=2)
    rng = np.random.default_rng(0)
    n = boxes.shape[0]
    ia = rng.integers(0, n, 4000); ib = np.clip(ia + rng.integers(-3, 4, 4000), 0, n - 1)
    # neighbours in a shuffled list rarely overlap: add explicit near-duplicates
    b2 = boxes.copy(); b2[:, 0::2] += rng.normal(0, 4, (n, 1)); b2[:, 1::2] += rng.normal(0, 4, (n, 1))
    got = ops.rotated_iou_pairs(_t(boxes, cuda_dev), _t(b2, cuda_dev)).cpu().numpy()
    ref = np.array([G.quad_iou(boxes[i], b2[i]) for i in range(n)])
    assert (ref > 0.3).sum() > n // 2
    err = np.abs(got - ref)
    assert err.max() < 5e-6
    big = ref >= 0.05
    assert (err[big] / ref[big]).max() < 1e-5          # IoU tolerance of the north star (fp32, relative)
    got_idx = ops.rotated_iou_pairs(_t(boxes, cuda_dev), _t(boxes, cuda_dev), _t(ia, cuda_dev), _t(ib, cuda_dev)).cpu().numpy()
    ref_idx = np.array([G.quad_iou(boxes[a], boxes[b]) for a, b in zip(ia, ib)])
    assert np.abs(got_idx - ref_idx).max() < 5e-6
    A, B = boxes[:130], b2[:97]
    mat = ops.rotated_iou_matrix(_t(A, cuda_dev), _t(B, cuda_dev)).cpu().numpy()
    refm = np.array([[G.quad_iou(a, b) for b in B] for a in A])
    assert np.abs(mat - refm).max() < 5e-6
    rs = ops.rotated_iou_matrix_sum(_t(A, cuda_dev), _t(B, cuda_dev)).cpu().numpy()
    assert rs.shape == (len(B),) and np.abs(rs - mat.astype(np.float64).sum(0)).max() < 1e-4


def test_iou_pairs_f64_is_the_reference_arithmetic(cuda_dev):
    """gm_rotated_iou_pairs_f64: the float64 value the reference computes, for a list of pairs (with and without index
    lists), concave simple quads included; and the fp32 dense kernel against it per IoU bucket (the numbers bench.py prints)."""
    from oriented_object_detection_b200 import ops, synth
    boxes, cls, conf = synth.synthetic_obbs(900, 16000, 16000, seed=5)
    rng = np.random.default_rng(1)
    n = boxes.shape[0]
    b2 = boxes.copy(); b2[:, 0::2] += rng.normal(0, 6, (n, 1)); b2[:, 1::2] += rng.normal(0, 6, (n, 1))
    b2[0] = [0, 0, 10, 0, 3, 3, 0, 10]                   # concave simple quad: valid for shapely
    boxes[0] = [0, 0, 10, 0, 10, 10, 0, 10]
    got = ops.rotated_iou_pairs_f64(_t(boxes, cuda_dev), _t(b2, cuda_dev)).cpu().numpy()
    ref = np.array([G.quad_iou(boxes[i], b2[i]) for i in range(n)])
    # device float64 (FMA-contracted products) against numpy float64 at 16 k map coordinates: 4e-11 measured
    assert got.dtype == np.float64 and np.abs(got - ref).max() < 1e-9
    ia = rng.integers(0, n, 3000); ib = np.clip(ia + rng.integers(-2, 3, 3000), 0, n - 1)
    got_idx = ops.rotated_iou_pairs_f64(_t(boxes, cuda_dev), _t(b2, cuda_dev), _t(ia, cuda_dev), _t(ib, cuda_dev)).cpu().numpy()
    assert np.abs(got_idx - np.array([G.quad_iou(boxes[a], b2[b]) for a, b in zip(ia, ib)])).max() < 1e-9
    # fp32 dense kernel against the float64 entry point, bucketed by IoU: absolute error everywhere, relative error
    # where the north star's 1e-5 applies meaningfully
    A, B = boxes[1:400], b2[1:400]
    mat = ops.rotated_iou_matrix(_t(A, cuda_dev), _t(B, cuda_dev)).cpu().numpy().astype(np.float64)
    ii, jj = np.nonzero(mat > 0)
    ref64 = ops.rotated_iou_pairs_f64(_t(A, cuda_dev), _t(B, cuda_dev), _t(ii, cuda_dev), _t(jj, cuda_dev)).cpu().numpy()
    err = np.abs(mat[ii, jj] - ref64)
    assert err.max() < 5e-6
    for lo, hi, rel in ((0.05, 0.3, 1e-5), (0.3, 1.01, 1e-5)):
        m = (ref64 >= lo) & (ref64 < hi)
        assert m.sum() > 0 and (err[m] / ref64[m]).max() < rel


def test_iou_matrix_with_invalid_and_ragged_shapes(cuda_dev):
    """Dense kernels on prepared records: invalid boxes (zero area, bow-tie, NaN) and concave quads read 0 against everything
    (the header's contract for the matrix entry points), valid pairs equal the oracle, for shapes that are not multiples of
    the CTA shape (256 rows x 128 columns) and for every GM_IOU_VARIANT-independent path (store and checksum)."""
    from oriented_object_detection_b200 import ops, synth
    boxes, _, _ = synth.synthetic_obbs(400, 900, 900, seed=11)
    boxes = boxes[:301].copy()
    bad = {3: [5, 5, 5, 5, 5, 5, 5, 5], 17: [0, 0, 10, 10, 10, 0, 0, 10], 40: [np.nan] * 8, 77: [0, 0, 10, 0, 3, 3, 0, 10],
           300: [1, 1, 9, 1, 9, 1, 1, 1]}
    for i, b in bad.items():
        boxes[i] = b
    A, B = boxes, boxes[::-1][:131].copy()
    mat = ops.rotated_iou_matrix(_t(A, cuda_dev), _t(B, cuda_dev)).cpu().numpy()
    assert mat.shape == (301, 131) and np.isfinite(mat).all()
    bad_b = {len(boxes) - 1 - i for i in bad if len(boxes) - 1 - i < 131}
    for i in bad:
        assert (mat[i] == 0).all()
    for j in bad_b:
        assert (mat[:, j] == 0).all()
    rng = np.random.default_rng(3)
    for _ in range(600):
        i, j = int(rng.integers(0, 301)), int(rng.integers(0, 131))
        if i in bad or j in bad_b:
            continue
        assert abs(mat[i, j] - G.quad_iou(A[i], B[j])) < 5e-6
    rs = ops.rotated_iou_matrix_sum(_t(A, cuda_dev), _t(B, cuda_dev)).cpu().numpy()
    assert np.abs(rs - mat.astype(np.float64).sum(0)).max() < 1e-4
    # one row against many columns and the reverse (grid edges on both axes)
    one = ops.rotated_iou_matrix(_t(A[:1], cuda_dev), _t(A, cuda_dev)).cpu().numpy()
    assert one.shape == (1, 301) and abs(one[0, 0] - 1.0) < 5e-6
    col = ops.rotated_iou_matrix(_t(A, cuda_dev), _t(A[:1], cuda_dev)).cpu().numpy()
    assert np.abs(col[:, 0] - one[0]).max() < 5e-6


def test_iou_degenerate_cases_and_host_api(cuda_dev):
    from oriented_object_detection_b200 import detect
    A = [0, 0, 10, 0, 10, 10, 0, 10]
    cases = [(A, A), (A, [10, 0, 20, 0, 20, 10, 10, 10]), (A, [0, 0, 20, 0, 20, 10, 0, 10]),
             (A, [5, 5, 15, 5, 15, 15, 5, 15]), (A, [0, 0, 10, 10, 10, 0, 0, 10]), (A, [0] * 8),
             (A, [0, 0, 0, 10, 10, 10, 10, 0]), ([x + 16000.25 for x in A], [x + 16000.25 for x in A])]
    for a, b in cases:
        assert abs(detect.compute_polygon_iou(a, b) - G.quad_iou(a, b)) < 1e-12


@pytest.mark.parametrize("case", ["sparse", "dense", "mapscale"])
def test_nms_matches_lifted_reference(cuda_dev, merge_golden, case):
    from oriented_object_detection_b200 import detect
    g = merge_golden["merge"][case]
    dets = [tuple(d) for d in g["dets"]]
    work = list(dets)
    kept = detect.merge_detections(work, 0.4)
    pos = {id(d): i for i, d in enumerate(dets)}
    assert [pos[id(d)] for d in work] == g["sorted"]       # caller's list sorted in place, stable
    assert [pos[id(d)] for d in kept] == g["kept"]         # identical members, identical order
    assert detect.merge_detections([], 0.4) == []


def test_nms_xlsx_fixed_point(cuda_dev, xlsx_rows):
    from oriented_object_detection_b200 import detect
    names = sorted({r[0] for nm in ("Test1", "Test2") for r in xlsx_rows[nm]["rows"]})
    for name in ("Test1", "Test2"):
        dets = [tuple(r[1:9]) + (names.index(r[0]), r[9], r[10]) for r in xlsx_rows[name]["rows"]]
        work = list(dets)
        assert detect.merge_detections(work, 0.4) == dets and work == dets


@pytest.mark.parametrize("n_obj,n_cls,extent,seed", [(900, 15, 4000, 0), (700, 2, 900, 1), (400, 1, 300, 2)])
def test_nms_random_sets_against_oracle(cuda_dev, n_obj, n_cls, extent, seed):
    from oriented_object_detection_b200 import ops, synth
    boxes, cls, conf = synth.synthetic_obbs(n_obj, extent, extent, n_classes=n_cls, seed=seed)
    conf[::7] = conf[3]                                            # confidence ties: stable order decides
    order, keep, kept = ops.nms_global(_t(boxes, cuda_dev), _t(cls, cuda_dev), _t(conf, cuda_dev), 0.4, max_class=n_cls - 1)
    want_order = sorted(range(len(conf)), key=lambda i: -float(conf[i]))
    assert order.cpu().tolist() == want_order
    want = G.nms_keep_indices(boxes, cls, conf, 0.4)
    assert kept.cpu().tolist() == want
    flags = np.zeros(len(conf), np.uint8); flags[want] = 1
    assert np.array_equal(keep.cpu().numpy(), flags)


def test_nms_edge_buffer_overflow_retries(cuda_dev):
    from oriented_object_detection_b200 import ops
    base = np.array([100, 100, 140, 100, 140, 120, 100, 120], dtype=np.float64)
    boxes = np.stack([base + 0.01 * k for k in range(300)])            # 300 near-identical boxes: 44850 pairs
    cls = np.zeros(300, np.int32); conf = np.linspace(0.9, 0.3, 300).astype(np.float32)
    order, keep, kept = ops.nms_global(_t(boxes, cuda_dev), _t(cls, cuda_dev), _t(conf, cuda_dev), 0.4, max_class=0,
                                       edge_capacity=100)
    assert kept.cpu().tolist() == [0]


@pytest.mark.parametrize("case", ["two", "three"])
def test_fusion_matches_lifted_reference(cuda_dev, merge_golden, case):
    from oriented_object_detection_b200 import detect
    g = merge_golden["fusion"][case]
    by_scale = {int(s): [tuple(d) for d in v] for s, v in g["by_scale"].items()}
    flat = [d for s in sorted(by_scale) for d in by_scale[s]]
    pos = {id(d): i for i, d in enumerate(flat)}
    kept = detect.cross_scale_consensus_filter(by_scale)
    assert [pos[id(d)] for d in kept] == g["kept"]
    single = {416: flat[:25]}
    out = detect.cross_scale_consensus_filter(single)
    assert out == flat[:25] and out is not single[416]


@pytest.mark.parametrize("seed,n_scales", [(0, 2), (1, 2), (2, 3)])
def test_fusion_random_against_oracle(cuda_dev, seed, n_scales):
    from oriented_object_detection_b200 import ops, synth
    boxes, cls, conf = synth.synthetic_obbs(500, 1200, 1200, n_classes=5, seed=10 + seed)
    conf = (conf * 1.1 - 0.1).clip(0.05, 0.999).astype(np.float32)      # some below CONS_LOW
    rng = np.random.default_rng(seed)
    sid = np.sort(rng.integers(0, n_scales, len(conf))).astype(np.int32)
    kept = ops.fuse_scales(_t(boxes, cuda_dev), _t(cls, cuda_dev), _t(conf, cuda_dev), _t(sid, cuda_dev), n_scales, max_class=4)
    dets = [tuple(boxes[i]) + (int(cls[i]), float(conf[i]), 0.0) for i in range(len(conf))]
    by_scale = {s: [d for d, k in zip(dets, sid) if k == s] for s in range(n_scales)}
    flat = [d for s in range(n_scales) for d in by_scale[s]]
    pos = {id(d): i for i, d in enumerate(flat)}
    want = [pos[id(d)] for d in G.cross_scale_consensus_filter(by_scale)]
    assert kept.cpu().tolist() == want


def test_detect_symbols_matches_lifted_reference(cuda_dev, merge_golden):
    """The reference's detect_symbols, replayed: same tile calls, same 11-tuples in the same order."""
    from oracle.lift_reference import FakeModel
    from oriented_object_detection_b200 import detect
    g = merge_golden["detect_symbols"]
    per_tile = iter(g["per_tile"])

    def fn(crop, conf):
        rec = next(per_tile)
        return (np.asarray(rec["corners"], dtype=np.float32).reshape(-1, 4, 2), rec["cls"], rec["conf"])

    model = FakeModel(fn)
    img = np.zeros((g["H"], g["W"], 3), np.uint8)
    detect.channels = 3
    out = detect.detect_symbols(img, model, g["tile"], g["overlap"])
    assert [list(c[0]) for c in model.calls] == g["calls"]
    assert all(c[1] == "uint8" and c[2] and c[3] == 0.25 for c in model.calls)
    want = [tuple(d) for d in g["out"]]
    assert len(out) == len(want)
    for a, b in zip(out, want):
        assert a[:10] == b[:10]                     # corners, class, confidence: exact
        assert abs(a[10] - b[10]) < 1e-9            # strike angle (float64 atan2)
        assert isinstance(a[8], int) and isinstance(a[9], float)


@pytest.mark.parametrize("case", ["map", "dense_tile", "ties_and_shuffle"])
def test_bounded_per_tile_nms_equals_the_engine(cuda_dev, monkeypatch, case):
    """gm_tile_postprocess_bounded (max_per_tile: one CTA per tile, score-sorted bit-mask sweep) against the general
    engine (max_per_tile = 0) and the oracle: a map-like set; a tile with more boxes than the bit matrix holds (the
    in-CTA sequential path); equal confidences (stable order by input index), shuffled tile ids, boxes the border filter
    drops, a concave quad.  GM_TILE_NMS_FAST=0 must route the bounded call to the engine."""
    from oriented_object_detection_b200 import ops, synth
    H, W = 1700, 1500
    plan = ops.make_plan(H, W, 416, 100, device=cuda_dev)
    rng = np.random.default_rng(21)
    if case == "map":
        local, cls, conf, tid = synth.synthetic_tile_dets(plan, 2500, n_classes=5, seed=8, margin=14)
    elif case == "dense_tile":
        local, cls, conf, tid = synth.synthetic_tile_dets(plan, 900, n_classes=3, seed=9, margin=14)
        # 700 extra boxes in tile 3: chains of heavily overlapping boxes of two classes
        m = 700
        c = rng.uniform(60, 350, (m, 2)).astype(np.float32)
        wh = rng.uniform(15, 60, (m, 2)).astype(np.float32)
        extra = np.stack([c[:, 0] - wh[:, 0], c[:, 1] - wh[:, 1], c[:, 0] + wh[:, 0], c[:, 1] - wh[:, 1],
                          c[:, 0] + wh[:, 0], c[:, 1] + wh[:, 1], c[:, 0] - wh[:, 0], c[:, 1] + wh[:, 1]], 1).astype(np.float32)
        local = np.concatenate([local, extra]); cls = np.concatenate([cls, rng.integers(0, 2, m).astype(cls.dtype)])
        conf = np.concatenate([conf, rng.uniform(0.25, 1, m).astype(np.float32)])
        tid = np.concatenate([tid, np.full(m, 3, tid.dtype)])
    else:
        local, cls, conf, tid = synth.synthetic_tile_dets(plan, 1800, n_classes=4, seed=10, margin=14)
        conf = (np.round(conf * 8) / 8).astype(np.float32)                  # many exactly equal confidences
        perm = rng.permutation(len(conf))
        local, cls, conf, tid = local[perm], cls[perm], conf[perm], tid[perm]
        local[5] = np.array([100, 100, 140, 100, 110, 110, 100, 140], np.float32)   # concave simple quad
        local[6] = np.array([100, 100, 140, 100, 140, 140, 100, 140], np.float32); cls[5] = cls[6]; tid[5] = tid[6]
    args = [_t(a, cuda_dev) for a in (local, cls, conf, tid)]
    ref = ops.tile_postprocess(*args, plan, 20, 1, 0.4, max_class=5)
    fast = ops.tile_postprocess(*args, plan, 20, 1, 0.4, max_class=5, max_per_tile=300)
    assert fast["src"].cpu().tolist() == ref["src"].cpu().tolist()
    for k in ("boxes", "cls", "conf", "angle"):
        assert torch.equal(fast[k], ref[k]), k
    assert 0 < len(ref["src"]) < len(conf)
    raw = ops.tile_postprocess(*args, plan, 20, 1, 0.4, max_class=5, max_per_tile=300, sync=False)
    k = int(raw["count"].item())
    assert k == len(ref["src"]) and raw["src"][:k].cpu().tolist() == ref["src"].cpu().tolist()
    monkeypatch.setenv("GM_TILE_NMS_FAST", "0")
    off = ops.tile_postprocess(*args, plan, 20, 1, 0.4, max_class=5, max_per_tile=300)
    assert off["src"].cpu().tolist() == ref["src"].cpu().tolist()
    if case == "dense_tile":
        assert np.bincount(tid).max() > 320


def test_tile_postprocess_synthetic_against_oracle(cuda_dev):
    from oriented_object_detection_b200 import ops, synth
    H, W = 2100, 2300
    plan = ops.make_plan(H, W, 416, 100, device=cuda_dev)
    local, cls, conf, tid = synth.synthetic_tile_dets(plan, 1500, n_classes=6, seed=4, margin=14)
    out = ops.tile_postprocess(_t(local, cuda_dev), _t(cls, cuda_dev), _t(conf, cuda_dev), _t(tid, cuda_dev), plan,
                               margin_px=20, angle_class=1, iou_merge=0.4, max_class=5)
    want = []
    for ti, t in enumerate(plan.tiles):
        sel = np.nonzero(tid == ti)[0]
        dets = []
        for i in sel:
            gp = [float(local[i, k]) + (int(t["x0"]) if k % 2 == 0 else int(t["y0"])) for k in range(8)]
            if not G.center_in_safe_region(gp, int(t["x0"]), int(t["y0"]), int(t["w"]), int(t["h"]), 20):
                continue
            ang = G.strike_angle([float(v) for v in local[i]]) if cls[i] == 1 else 0.0
            dets.append(tuple(gp) + (int(cls[i]), float(conf[i]), ang, int(i)))
        want.extend(G.merge_detections(dets, 0.4))
    assert out["src"].cpu().tolist() == [d[11] for d in want]
    assert np.array_equal(out["boxes"].cpu().numpy(), np.array([d[:8] for d in want]))
    assert np.abs(out["angle"].cpu().numpy() - np.array([d[10] for d in want])).max() < 1e-9
    assert len(want) < len(conf) and len(want) > 100


def test_window_iou_coincident_edges_and_general_quads_gpu(cuda_dev):
    """Device arithmetic (approximate reciprocals) of the boundary-integral IoU on the cases where a
    boundary piece could be counted twice or not at all, and on general convex quadrilaterals."""
    from oriented_object_detection_b200 import ops
    rng = np.random.default_rng(11)

    def rb(cx, cy, w, h, th):
        c, s = np.cos(th), np.sin(th)
        v1 = np.array([w / 2 * c, w / 2 * s]); v2 = np.array([-h / 2 * s, h / 2 * c]); ctr = np.array([cx, cy])
        return np.concatenate([ctr + v1 + v2, ctr + v1 - v2, ctr - v1 - v2, ctr - v1 + v2])

    A, B, kinds = [], [], []
    for _ in range(4000):
        cx, cy = rng.uniform(0, 16000, 2)
        w, h = rng.uniform(12, 100, 2)
        th = [0.0, np.pi / 2, np.pi / 4, rng.uniform(-1, 2)][rng.integers(4)]
        a = rb(cx, cy, w, h, th)
        if rng.integers(2):
            a = a.astype(np.float32).astype(np.float64)
        kind = rng.integers(8)
        if kind == 0: b = a.copy()
        elif kind == 1: b = a[[2, 3, 4, 5, 6, 7, 0, 1]]
        elif kind == 2: b = a[[6, 7, 4, 5, 2, 3, 0, 1]]
        elif kind == 3: b = rb(cx + w * np.cos(th), cy + w * np.sin(th), w, h, th)
        elif kind == 4: b = rb(cx, cy, w * 0.5, h * 0.5, th)
        elif kind == 5: b = rb(cx + w * 0.25 * np.cos(th), cy + w * 0.25 * np.sin(th), w * 0.5, h, th)
        elif kind == 6: b = a + rng.normal(0, 1e-4, 8)
        else: b = rb(cx, cy, h, w, th + np.pi / 2)
        A.append(a); B.append(b); kinds.append(int(kind))
    for _ in range(4000):
        c = rng.uniform(100, 5000, 2)
        for ctr, dst in ((c, A), (c + rng.normal(0, 20, 2), B)):
            ang = np.sort(rng.uniform(0, 2 * np.pi, 4)); r = rng.uniform(10, 60, 4)
            dst.append((ctr + np.stack([r * np.cos(ang), r * np.sin(ang)], 1)).ravel())
    A, B = np.array(A), np.array(B)
    got = ops.rotated_iou_pairs(_t(A, cuda_dev), _t(B, cuda_dev)).cpu().numpy()
    ref = np.array([G.quad_iou(a, b) for a, b in zip(A, B)])
    err = np.abs(got - ref)
    jitter = np.array(kinds + [-1] * 4000) == 6         # an edge of A inside the sliver of a near-parallelogram window (geom.cuh)
    assert err[~jitter].max() < 3e-6
    assert err[jitter].max() < 6e-6                     # slab form, first-order sliver term; IoU ~ 1 for these pairs
    assert (ref[4000:] > 0).mean() > 0.1


def test_padded_one_sync_merge_equals_synchronous_path(cuda_dev):
    """tile_postprocess(sync=False) + sharding.merge_bands_padded (single rank, capacity > count) give the
    members and order of the synchronous tile_postprocess + nms_global path."""
    from oriented_object_detection_b200 import ops, sharding, synth
    H, W = 3000, 3300
    plan = ops.make_plan(H, W, 416, 100, device=cuda_dev)
    local, cls, conf, tid = synth.synthetic_tile_dets(plan, 6000, n_classes=7, seed=12, margin=20)
    args = (_t(local, cuda_dev), _t(cls, cuda_dev), _t(conf, cuda_dev), _t(tid, cuda_dev), plan, 20, 1, 0.4)
    pp = ops.tile_postprocess(*args, max_class=6)
    kept = ops.nms_global(pp["boxes"], pp["cls"], pp["conf"], 0.4, max_class=6)[2].to(torch.int64)
    raw = ops.tile_postprocess(*args, max_class=6, sync=False)
    assert raw["boxes"].shape[0] == len(conf) and int(raw["count"].item()) == pp["conf"].shape[0]
    got = sharding.merge_bands_padded(raw, raw["count"], len(conf) + 100, 0.4, 6)
    assert got["index"].cpu().tolist() == kept.cpu().tolist()
    assert torch.equal(got["boxes"], pp["boxes"][kept]) and torch.equal(got["angle"], pp["angle"][kept])
    assert got["n_valid"] == pp["conf"].shape[0]


def test_captured_detection_path_replays_like_the_eager_calls(cuda_dev):
    """sharding.CapturedCall: tile_postprocess(sync=False) + merge_bands_device recorded into one CUDA graph; replays
    on refreshed input tensors give the members and order of the synchronous path, twice with different inputs."""
    from oriented_object_detection_b200 import ops, sharding, synth
    H, W = 3000, 3300
    plan = ops.make_plan(H, W, 416, 100, device=cuda_dev)
    sets = [synth.synthetic_tile_dets(plan, 6000, n_classes=7, seed=s, margin=20) for s in (12, 13)]
    n = min(len(s[2]) for s in sets)
    sets = [tuple(a[:n] for a in s) for s in sets]                      # one static input size for both replays
    d_in = [_t(a, cuda_dev).clone() for a in sets[0]]

    def device_part():
        pp = ops.tile_postprocess(d_in[0], d_in[1], d_in[2], d_in[3], plan, 20, 1, 0.4, max_class=6, sync=False)
        return sharding.merge_bands_device(pp, pp["count"], n + 64, 0.4, 6)

    call = sharding.CapturedCall(device_part)
    assert call.captured, call.error
    assert call.launches > 20
    for s in (sets[1], sets[0]):
        for d, a in zip(d_in, s):
            d.copy_(_t(a, cuda_dev))
        got = sharding.merge_bands_finish(call())
        pp = ops.tile_postprocess(*[_t(a, cuda_dev) for a in s], plan, 20, 1, 0.4, max_class=6)
        kept = ops.nms_global(pp["boxes"], pp["cls"], pp["conf"], 0.4, max_class=6)[2].to(torch.int64)
        assert got["index"].cpu().tolist() == kept.cpu().tolist()
        assert torch.equal(got["boxes"], pp["boxes"][kept]) and torch.equal(got["conf"], pp["conf"][kept])
        assert got["n_valid"] == pp["conf"].shape[0]


def test_band_record_kernels_emulated_two_ranks(cuda_dev):
    """gm_band_pack / unpack / mask_keep / extract with the collectives emulated on one GPU: two bands are packed with a
    common capacity, concatenated like an all_gather, each 'rank' resolves its classes, the keep flags are OR-ed like
    the all_reduce, and the compacted result must equal the single-rank NMS of the concatenated survivors."""
    from oriented_object_detection_b200 import ops, synth
    H, W = 3000, 3300
    plan = ops.make_plan(H, W, 416, 100, device=cuda_dev)
    local, cls, conf, tid = synth.synthetic_tile_dets(plan, 7000, n_classes=7, seed=21, margin=20)
    pp = ops.tile_postprocess(_t(local, cuda_dev), _t(cls, cuda_dev), _t(conf, cuda_dev), _t(tid, cuda_dev), plan, 20, 1, 0.4,
                              max_class=6)
    n = pp["conf"].shape[0]
    c0 = n // 2 + 37
    counts = [c0, n - c0]
    cap = max(counts) + 50
    recs = []
    for b, (lo, hi) in enumerate(((0, c0), (c0, n))):
        pad = 64 + 13 * b                                                   # input arrays longer than the count, garbage behind it
        part = {k: torch.cat([pp[k][lo:hi], pp[k][:pad]]) for k in ("boxes", "cls", "conf", "angle")}
        recs.append(ops.band_pack(part, torch.tensor([hi - lo], dtype=torch.int64, device=cuda_dev), cap))
        assert recs[-1].shape == (cap, ops.BAND_RECORD_BYTES)
    recv = torch.cat(recs).contiguous()
    keep_all, order0, u0 = None, None, None
    for rank in (0, 1):
        u = ops.band_unpack(recv, 2, rank)
        assert int(u["n_valid"].item()) == n
        assert torch.equal(u["cls"][:c0], pp["cls"][:c0]) and bool((u["cls"][c0:cap] == -1).all())
        assert torch.equal(u["boxes"][cap:cap + counts[1]], pp["boxes"][c0:]) and bool(torch.isnan(u["boxes"][c0:cap]).all())
        assert bool(((u["cls_owned"] >= 0) == ((u["cls"] >= 0) & (u["cls"] % 2 == rank))).all())
        order, keep, _, _ = ops.nms_global(u["boxes"], u["cls_owned"], u["conf"], 0.4, max_class=6, sync=False)
        ops.band_mask_keep(keep, u["cls_owned"])
        keep_all = keep.clone() if keep_all is None else torch.maximum(keep_all, keep)
        if rank == 0:
            order0, u0 = order, u
        else:
            assert torch.equal(order, order0)                              # the confidence order does not depend on the class mask
    x = ops.band_extract(order0, keep_all, u0)
    m = int(x["n_out"].item())
    kept = ops.nms_global(pp["boxes"], pp["cls"], pp["conf"], 0.4, max_class=6)[2].to(torch.int64)
    want_pos = torch.where(kept < c0, kept, kept - c0 + cap)
    assert m == kept.numel()
    assert x["index"][:m].cpu().tolist() == want_pos.cpu().tolist()
    assert torch.equal(x["boxes"][:m], pp["boxes"][kept]) and torch.equal(x["conf"][:m], pp["conf"][kept])
    assert torch.equal(x["cls"][:m], pp["cls"][kept]) and torch.equal(x["angle"][:m], pp["angle"][kept])


def test_float64_confidences_decide_like_the_reference(cuda_dev):
    """Callers may pass genuine float64 confidences to the list API: 0.7 (float64) must pass the CONS_HIGH >= 0.70 test
    although float32(0.7) = 0.69999999 would not, and confidences that differ only beyond float32 precision must sort as
    the reference sorts them (Detect_OBB.py:183, :401).  The mirror maps them to dense ranks before the kernels."""
    from oracle import geometry as G
    from oriented_object_detection_b200 import detect
    sq = lambda x, y, s: (x, y, x + s, y, x + s, y + s, x, y + s)
    # fusion: solo detections at exactly the two thresholds, and a pair decided by a sub-float32 confidence difference
    small = [sq(0, 0, 20) + (0, 0.7, 0.0), sq(100, 0, 20) + (0, 0.25, 0.0), sq(200, 0, 20) + (1, 0.5000000001, 0.0)]
    large = [sq(300, 0, 20) + (0, 0.6999999999, 0.0), sq(201, 0, 20) + (1, 0.5000000002, 0.0)]
    want = G.cross_scale_consensus_filter({128: list(small), 416: list(large)})
    got = detect.cross_scale_consensus_filter({128: list(small), 416: list(large)})
    assert got == want and small[0] in got and large[0] not in got and large[1] in got and small[2] not in got
    # NMS: the winner of an overlapping pair is the one whose float64 confidence is larger by 1e-10
    dets = [sq(0, 0, 20) + (0, 0.3000000001, 0.0), sq(1, 0, 20) + (0, 0.3000000002, 0.0), sq(50, 0, 20) + (0, 0.9, 0.0)]
    ref = list(dets)
    assert detect.merge_detections(dets, 0.4) == G.merge_detections(ref, 0.4) and dets == ref and len(dets) == 3
