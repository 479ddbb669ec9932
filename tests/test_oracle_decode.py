"""Sanity of the numpy restatement of the Ultralytics OBB predictor tail (no upstream to pin against)."""
import numpy as np

from oracle import decode as D


def test_probiou_identity_and_separation():
    b = np.array([50, 60, 30, 12, 0.4], dtype=np.float32)
    assert D.probiou(b, b) > 0.99
    far = b.copy(); far[0] += 500
    assert D.probiou(b, far) < 1e-3
    rot = b.copy(); rot[4] += np.float32(np.pi)        # same rectangle
    assert D.probiou(b, rot) > 0.99


def test_decode_tile_shapes_order_and_unletterbox():
    head = D.synthetic_head(1, 5, 128, seed=3, density=0.05)[0]
    b, c, f = D.decode_tile(head, 23, 13, 128)
    assert b.shape[1] == 8 and len(b) == len(c) == len(f) and len(f) > 0
    assert (np.diff(f) <= 0).all() and (f > 0.25).all()
    # a 23x13 tile is upscaled by 128/23 and centred: network x in [0, 128] maps to [-pad, 128-pad]/gain
    gain = 128 / 23
    pad = round((128 - 13 * gain) / 2 - 0.1)
    cx = b[:, 0::2].mean(1)
    assert cx.min() > -pad / gain - 1 and cx.max() < (128 - pad) / gain + 1
    # exact duplicates injected by synthetic_head never both survive
    full, _, ff = D.decode_tile(head, 128, 128, 128, conf_thr=0.25, iou_thr=0.7)
    none_removed, _, fn = D.decode_tile(head, 128, 128, 128, conf_thr=0.25, iou_thr=1.01)
    assert len(fn) > len(ff)


def test_corner_order_matches_strike_angle_convention():
    # pt4 - pt1 = -2*v1: the edge compute_angle_from_bbox reads is the width axis
    head = np.zeros((4 + 2 + 1, 1), dtype=np.float32)
    head[:4, 0] = [64, 64, 40, 10]
    head[4, 0] = 0.9
    head[6, 0] = 0.3
    b, c, f = D.decode_tile(head, 128, 128, 128)
    v = b[0, 6:8] - b[0, 0:2]
    assert np.allclose(v, [-40 * np.cos(0.3), -40 * np.sin(0.3)], atol=1e-4)


def test_probiou_equals_the_published_gaussian_definition():
    """ProbIoU (Llerena et al.): a box is the Gaussian N(centre, R diag(w^2/12, h^2/12) R^T); the similarity is
    1 - sqrt(1 - exp(-BD)) with BD the Bhattacharyya distance.  The restated scalar formula (Ultralytics' batch_probiou,
    SURVEY.md Appendix B) against that definition evaluated with matrix algebra in float64."""
    import numpy as np
    from oracle import decode as D
    rng = np.random.default_rng(0)

    def gauss(b):
        cx, cy, w, h, th = (float(v) for v in b)
        R = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
        return np.array([cx, cy]), R @ np.diag([w * w / 12.0, h * h / 12.0]) @ R.T

    worst = 0.0
    for _ in range(500):
        b1 = np.array([rng.uniform(0, 400), rng.uniform(0, 400), rng.uniform(8, 120), rng.uniform(8, 120), rng.uniform(-0.78, 2.35)], np.float32)
        b2 = b1 + np.array([rng.normal(0, 12), rng.normal(0, 12), rng.normal(0, 8), rng.normal(0, 8), rng.normal(0, 0.4)], np.float32)
        b2[2:4] = np.maximum(b2[2:4], 4)
        m1, S1 = gauss(b1)
        m2, S2 = gauss(b2)
        S = 0.5 * (S1 + S2)
        d = m1 - m2
        bd = 0.125 * d @ np.linalg.solve(S, d) + 0.5 * np.log(np.linalg.det(S) / np.sqrt(np.linalg.det(S1) * np.linalg.det(S2)))
        want = 1.0 - np.sqrt(max(1.0 - np.exp(-bd), 0.0))
        worst = max(worst, abs(float(D.probiou(b1, b2)) - want))
    assert worst < 2e-3                    # fp32 + the eps terms of the restated formula (1e-7 under a square root)


def test_corners_equal_opencv_box_points():
    """xywhr -> 4 corners (Ultralytics' xywhr2xyxyxyxy as restated in decode_tile) against cv2.boxPoints, as point sets."""
    import cv2
    import numpy as np
    from oracle import decode as D
    nc, A = 3, 6
    head = np.zeros((4 + nc + 1, A), np.float32)
    rng = np.random.default_rng(1)
    head[0], head[1] = rng.uniform(40, 90, A), rng.uniform(40, 90, A)
    head[2], head[3] = rng.uniform(10, 40, A), rng.uniform(10, 40, A)
    head[4 + nc] = rng.uniform(-0.78, 2.35, A)
    head[4] = np.linspace(0.9, 0.4, A)                   # one class, distinct confidences
    head[0] += np.arange(A) * 200                        # far apart: nothing is suppressed
    boxes, cls, conf = D.decode_tile(head, 2000, 2000, 2000, 0.25, 0.7, 300)
    assert len(conf) == A
    for k in range(A):
        want = cv2.boxPoints(((float(head[0, k]), float(head[1, k])), (float(head[2, k]), float(head[3, k])),
                              float(np.degrees(head[4 + nc, k]))))
        got = boxes[k].reshape(4, 2)
        for p in want:
            assert np.abs(got - p).sum(1).min() < 2e-3
