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
