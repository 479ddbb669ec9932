"""csrc/geom.cuh compiled for the HOST (its functions are __host__ __device__): the edge-integration
IoU formulation against the float64 Sutherland-Hodgman oracle, without a GPU."""
import os
import struct
import subprocess

import numpy as np
import pytest

from oracle import geometry as G

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("geom") / "geom_host")
    subprocess.run(["nvcc", "-O2", "-Wno-deprecated-gpu-targets", "-o", exe,
                    os.path.join(ROOT, "tests", "host_harness", "geom_host.cu")], check=True)
    return exe


def _run(exe, pairs):
    arr = np.array([np.concatenate(p) for p in pairs], dtype=np.float64)
    out = subprocess.run([exe], input=struct.pack("q", len(arr)) + arr.tobytes(), capture_output=True, check=True).stdout
    # fp32 clip, fp64 clip, fp32 window, fp32 general-quad window, slab path taken, packed form lane 0 / lane 1, scalar form of lane 1
    return np.frombuffer(out, dtype=np.float64).reshape(-1, 8)


def _rbox(cx, cy, w, h, th):
    c, s = np.cos(th), np.sin(th)
    v1 = np.array([w / 2 * c, w / 2 * s]); v2 = np.array([-h / 2 * s, h / 2 * c]); ctr = np.array([cx, cy])
    return np.concatenate([ctr + v1 + v2, ctr + v1 - v2, ctr - v1 - v2, ctr - v1 + v2]).astype(np.float32).astype(np.float64)


def test_random_pairs_at_map_scale(harness):
    rng = np.random.default_rng(0)
    pairs = []
    for scale in (1000.0, 16384.0):
        for _ in range(1500):
            cx, cy = rng.uniform(0, scale, 2)
            w, h, th = rng.uniform(12, 100), rng.uniform(11, 97), rng.uniform(-np.pi / 4, 3 * np.pi / 4)
            a = _rbox(cx, cy, w, h, th)
            b = _rbox(cx + rng.normal(0, 15), cy + rng.normal(0, 15), w * rng.uniform(0.7, 1.3),
                      h * rng.uniform(0.7, 1.3), th + rng.normal(0, 0.3))
            if rng.random() < 0.2:
                b = b[[6, 7, 4, 5, 2, 3, 0, 1]]
            pairs.append((a, b))
    got = _run(harness, pairs)
    ref = np.array([G.quad_iou(a, b) for a, b in pairs])
    assert np.abs(got[:, 1] - ref).max() < 1e-12                   # float64 path
    big = ref >= 0.05
    for col in (0, 2):                                             # fp32 pair-local paths: clip, window
        err = np.abs(got[:, col] - ref)
        assert err.max() < 5e-6                                    # absolute
        assert (err[big] / ref[big]).max() < 1e-5                  # north-star tolerance where IoU is not tiny


def test_degenerate_and_exact_cases(harness):
    q = lambda *p: np.array(p, dtype=np.float64)
    A = q(0, 0, 10, 0, 10, 10, 0, 10)
    cases = [(A, A.copy(), 1.0), (A, q(10, 0, 20, 0, 20, 10, 10, 10), 0.0), (A, q(0, 0, 20, 0, 20, 10, 0, 10), 0.5),
             (A, q(5, 5, 15, 5, 15, 15, 5, 15), 25 / 175), (A, q(2, 2, 8, 2, 8, 8, 2, 8), 0.36),
             (A, q(10, 10, 20, 10, 20, 20, 10, 20), 0.0), (A, q(0, 0, 10, 10, 10, 0, 0, 10), 0.0),
             (A, q(0, 0, 0, 0, 0, 0, 0, 0), 0.0), (A, q(0, 0, 10, 0, 20, 0, 30, 0), 0.0),
             (A, q(0, 10, 10, 10, 10, 20, 0, 20), 0.0), (A, q(0, 0, 0, 10, 10, 10, 10, 0), 1.0),
             (A + 16000, A + 16000, 1.0), (q(3, 0, 10, 0, 10, 10, 0, 10), q(0, 0, 10, 0, 10, 7, 0, 10), None)]
    got = _run(harness, [(a, b) for a, b, _ in cases])
    for (a, b, want), g in zip(cases, got):
        ref = G.quad_iou(a, b)
        if want is not None:
            assert abs(ref - want) < 1e-12
        assert abs(g[1] - ref) < 1e-12 and abs(g[0] - ref) < 1e-6 and abs(g[2] - ref) < 1e-6


def test_window_iou_on_coincident_edges_and_general_quads(harness):
    """The boundary-integral (window) formulation must count coincident boundary pieces exactly once:
    identical boxes under every vertex labelling, boxes sharing an edge, flush inner boxes, jitter at
    the rounding level, and general convex quadrilaterals (no parallelogram assumption)."""
    rng = np.random.default_rng(7)

    def rb(cx, cy, w, h, th):
        c, s = np.cos(th), np.sin(th)
        v1 = np.array([w / 2 * c, w / 2 * s]); v2 = np.array([-h / 2 * s, h / 2 * c]); ctr = np.array([cx, cy])
        return np.concatenate([ctr + v1 + v2, ctr + v1 - v2, ctr - v1 - v2, ctr - v1 + v2])

    pairs, kinds = [], []
    for _ in range(1600):
        cx, cy = rng.uniform(0, 8000, 2)
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
        pairs.append((a, b)); kinds.append(int(kind))
    for _ in range(1600):
        c = rng.uniform(100, 5000, 2)
        quads = []
        for ctr in (c, c + rng.normal(0, 20, 2)):
            ang = np.sort(rng.uniform(0, 2 * np.pi, 4)); r = rng.uniform(10, 60, 4)
            quads.append((ctr + np.stack([r * np.cos(ang), r * np.sin(ang)], 1)).ravel())
        pairs.append(tuple(quads))
    got = _run(harness, pairs)
    # sorted angles with free radii also produce CONCAVE simple quads: valid for shapely (Detect_OBB.py:148-151), so
    # the float64 path intersects them piecewise (two triangles at the reflex vertex) and must equal the oracle; the
    # fp32 window forms are defined for convex quads only and return 0 (their callers take the float64 path).
    concave = np.array([any(G.quad_classify([tuple(v) for v in q.reshape(4, 2)])[0] == 2 for q in pr) for pr in pairs])
    assert 100 < concave.sum() < 1600 and not concave[:1600].any()
    ref = np.array([G.quad_iou(a, b) for a, b in np.asarray(pairs, dtype=object)[concave]])
    assert np.abs(got[concave, 1] - ref).max() < 1e-12 and (ref > 0.05).sum() > 20
    assert (got[concave, 2] == 0).all() and (got[concave, 3] == 0).all()
    err = np.abs(got[:, 2] - got[:, 1])
    jitter = np.array(kinds + [-1] * 1600) == 6         # an edge of A inside the sliver between B and its parallelogram
    assert err[~jitter & ~concave].max() < 2e-6
    assert err[jitter].max() < 5e-6                     # slab form: first-order sliver term (geom.cuh), IoU ~ 1 there
    assert np.abs(got[:, 3] - got[:, 1])[~concave].max() < 2e-6   # the general form on every convex pair
    assert got[:1600, 4].mean() > 0.8 and got[1600:, 4].max() == 0     # rectangles take the slab form, general quads do not
    assert (got[1600:, 1] > 0).mean() > 0.1            # the general quads do overlap


def test_slab_form_stress_families(harness):
    """The dispatched fp32 window IoU against the float64 clip of the same harness (itself equal to the oracle to 1e-12,
    first test) on 9 families x 20 k pairs with pipeline-faithful rounding (fp32 tile-local corners + integer tile
    offset): random neighbours, axis-aligned boxes on integer shifts, integer grids full of exact coincidences, boxes
    sharing a corner, overlaps of <= 1e-3 px, nested boxes, 16 k map coordinates, 0.5-3 px thin boxes, and the same
    rectangle described with w/h swapped and theta + 90 degrees."""
    rng = np.random.default_rng(4)
    N = 20000

    def rb(cx, cy, w, h, th):
        c, s = np.cos(th), np.sin(th)
        v1 = np.stack([w / 2 * c, w / 2 * s], 1); v2 = np.stack([-h / 2 * s, h / 2 * c], 1); ctr = np.stack([cx, cy], 1)
        return np.concatenate([ctr + v1 + v2, ctr + v1 - v2, ctr - v1 - v2, ctr - v1 + v2], 1)

    limits = {"random": 2e-6, "axis": 2e-6, "grid": 5e-7, "shared_corner": 2e-6, "tiny_overlap": 1e-6, "nested": 2e-6,
              "scale16k": 2e-6, "thin": 8e-6, "rot90": 6e-6}
    for name, limit in limits.items():
        cx, cy = rng.uniform(0, 4000, N), rng.uniform(0, 4000, N)
        w, h = rng.uniform(12, 100, N), rng.uniform(11, 97, N)
        th = rng.uniform(-np.pi / 4, 3 * np.pi / 4, N)
        if name == "random":
            A = rb(cx, cy, w, h, th)
            B = rb(cx + rng.normal(0, 20, N), cy + rng.normal(0, 20, N), w * rng.uniform(.5, 1.5, N), h * rng.uniform(.5, 1.5, N),
                   th + rng.normal(0, .5, N))
        elif name == "axis":
            th0 = rng.choice([0, np.pi / 2, np.pi / 4, -np.pi / 4], N)
            A = rb(cx, cy, w, h, th0)
            B = rb(cx + rng.integers(-30, 30, N), cy + rng.integers(-30, 30, N), w, h, th0)
        elif name == "grid":
            cx, cy, w, h = np.round(cx), np.round(cy), 2 * rng.integers(6, 50, N), 2 * rng.integers(6, 50, N)
            z = np.zeros(N)
            A = rb(cx, cy, w, h, z)
            B = rb(cx + rng.integers(-3, 4, N) * (w // 2), cy + rng.integers(-3, 4, N) * (h // 2), w * rng.choice([.5, 1, 2], N),
                   h * rng.choice([.5, 1, 2], N), z)
        elif name == "shared_corner":
            A = rb(cx, cy, w, h, th)
            s, a = rng.uniform(.3, 1.5, N)[:, None], rng.normal(0, .2, N)
            P = A.reshape(N, 4, 2); O = P[:, :1]
            R = np.stack([np.stack([np.cos(a), -np.sin(a)], 1), np.stack([np.sin(a), np.cos(a)], 1)], 1)
            B = (O + np.einsum('nij,nkj->nki', R, (P - O)) * s[:, :, None]).reshape(N, 8)
        elif name == "tiny_overlap":
            A = rb(cx, cy, w, h, th)
            d = w - rng.uniform(0, 1e-3, N)
            B = rb(cx + d * np.cos(th), cy + d * np.sin(th), w, h, th + rng.normal(0, 1e-3, N))
        elif name == "nested":
            A = rb(cx, cy, w, h, th)
            B = rb(cx + rng.normal(0, 1, N), cy + rng.normal(0, 1, N), w * rng.uniform(.1, .9, N), h * rng.uniform(.1, .9, N),
                   th + rng.normal(0, .3, N))
        elif name == "scale16k":
            cx, cy = cx * 4, cy * 4
            A = rb(cx, cy, w, h, th)
            B = rb(cx + rng.normal(0, 10, N), cy + rng.normal(0, 10, N), w, h, th + rng.normal(0, .2, N))
        elif name == "thin":
            h = rng.uniform(0.5, 3, N)
            A = rb(cx, cy, w, h, th)
            B = rb(cx + rng.normal(0, 2, N), cy + rng.normal(0, 2, N), w, h * rng.uniform(.5, 2, N), th + rng.normal(0, .05, N))
        else:
            A = rb(cx, cy, w, h, th)
            B = rb(cx, cy, h, w, th + np.pi / 2) + rng.choice([0, 1e-5, 1e-3], N)[:, None]
        off = np.tile(np.stack([np.floor(cx / 316) * 316, np.floor(cy / 316) * 316], 1), (1, 4))
        A = (A - off).astype(np.float32).astype(np.float64) + off
        B = (B - off).astype(np.float32).astype(np.float64) + off
        got = _run(harness, list(zip(A, B)))
        err = np.abs(got[:, 2] - got[:, 1])
        assert err.max() < limit, (name, float(err.max()))
        assert got[:, 4].mean() > 0.7, name                     # the slab form is what is being exercised


def test_float64_path_is_contraction_proof(tmp_path):
    """nvcc contracts a*b - c*d into an FMA in device code; for a collapsed label (four equal corners) that turns an
    exactly-zero area into a rounding residue and an invalid polygon into a "valid" one.  The host build with
    -mfma -ffp-contract=fast shows the same behaviour as the device build: the float64 IoU of every (detection, label)
    pair of both evaluation goldens - bow-ties, collapsed and concave labels included - must equal the oracle."""
    import json
    exe = str(tmp_path / "geom_host_fma")
    subprocess.run(["nvcc", "-O2", "-Wno-deprecated-gpu-targets", "-Xcompiler", "-mfma,-ffp-contract=fast", "-o", exe,
                    os.path.join(ROOT, "tests", "host_harness", "geom_host.cu")], check=True)
    pairs = []
    for name in ("eval_golden.json", "eval_concave_golden.json"):
        with open(os.path.join(ROOT, "tests", "golden", name)) as fh:
            g = json.load(fh)
        for rec in g["images"].values():
            for d in rec["dets"][:120]:
                pairs += [(np.array(d[:8]), np.array(gt["pts"]).ravel()) for gt in rec["gts"] if gt["cls"] == d[8]]
    got = _run(exe, pairs)
    ref = np.array([G.quad_iou(a, b) for a, b in pairs])
    assert len(pairs) > 3000 and np.abs(got[:, 1] - ref).max() < 1e-12


def test_packed_two_polygon_form_equals_the_scalar_slab_form(harness):
    """qbox_iou_rect2 (the f32x2 form of the dense kernel; here through its struct emulation) against qbox_iou_rect on the
    same polygons: lane 0 = the pair itself, lane 1 = a DIFFERENT polygon against the same window.  The two forms differ
    only in where the negations of X and Y happen, so they agree to rounding; both stay within the fp32 tolerance of float64."""
    rng = np.random.default_rng(11)
    pairs = []
    for k in range(6000):
        cx, cy = rng.uniform(0, 16000, 2)
        w, h, th = rng.uniform(12, 100), rng.uniform(11, 97), [0.0, np.pi / 2, rng.uniform(-1, 2)][k % 3]
        a = _rbox(cx, cy, w, h, th)
        kind = k % 5
        if kind == 0: b = a.copy()
        elif kind == 1: b = _rbox(cx + rng.normal(0, 15), cy + rng.normal(0, 15), w * rng.uniform(.7, 1.3), h * rng.uniform(.7, 1.3), th + rng.normal(0, .3))
        elif kind == 2: b = _rbox(cx + w * np.cos(th), cy + w * np.sin(th), w, h, th)            # shared edge
        elif kind == 3: b = _rbox(cx, cy, w * 0.5, h * 0.5, th)                                  # nested
        else: b = a + rng.normal(0, 1e-4, 8)
        pairs.append((a, b))
    got = _run(harness, pairs)
    rect = got[:, 4] == 1
    assert rect.mean() > 0.6                                          # fp32 GLOBAL corners at 16 k px: not every box is a parallelogram to 5e-6
    assert np.abs(got[rect, 5] - got[rect, 2]).max() < 2e-6          # lane 0 against the scalar form of the same pair
    assert np.abs(got[rect, 6] - got[rect, 7]).max() < 2e-6          # lane 1 against the scalar form of (next polygon, this window)
    jitter = (np.arange(len(pairs)) % 5 == 4)
    assert np.abs(got[rect & ~jitter, 5] - got[rect & ~jitter, 1]).max() < 3e-6     # and against float64
    assert np.abs(got[rect & jitter, 5] - got[rect & jitter, 1]).max() < 6e-6
