"""oracle/letterbox.py: the integer restatement of cv2's 8-bit INTER_LINEAR against cv2.resize itself
(IPP off), and the LetterBox geometry on the tile shapes of the reference's plans."""
import numpy as np
import pytest

from oracle import letterbox as LB


@pytest.mark.parametrize("shape,new", [((175, 263, 3), (277, 416)), ((23, 13, 3), (128, 72)), ((292, 292, 4), (416, 416)),
                                       ((48, 76, 3), (81, 128)), ((416, 416, 3), (200, 311)), ((1, 40, 3), (3, 128)),
                                       ((33, 1, 4), (128, 4)), ((58, 58, 3), (128, 128)), ((100, 37, 1), (345, 128))])
def test_resize_restatement_equals_cv2(shape, new):
    import cv2
    rng = np.random.default_rng(sum(shape) + sum(new))
    for kind in range(3):
        img = rng.integers(0, 256, shape, dtype=np.uint8)
        if kind == 1:
            img = np.where(rng.random(shape) < 0.5, 0, 255).astype(np.uint8)
        if kind == 2:
            img = np.repeat(np.repeat(img[::4, ::4], 4, 0), 4, 1)[:shape[0], :shape[1]]
        want = cv2.resize(np.ascontiguousarray(img), (new[1], new[0]), interpolation=cv2.INTER_LINEAR)
        if want.ndim == 2:
            want = want[..., None]
        assert np.array_equal(LB.resize_linear_u8(img, new[0], new[1]), want)


def test_letterbox_geometry_of_the_reference_tile_shapes():
    # full tiles are untouched; ragged tiles are padded to a multiple of 32 (rect) and/or upscaled
    assert LB.letterbox_geometry(416, 416, 416) == (416, 416, 0, 0, 416, 416)
    assert LB.letterbox_geometry(175, 263, 416) == (277, 416, 5, 0, 288, 416)
    assert LB.letterbox_geometry(416, 292, 416) == (416, 292, 0, 14, 416, 320)
    assert LB.letterbox_geometry(292, 292, 416) == (416, 416, 0, 0, 416, 416)
    assert LB.letterbox_geometry(23, 13, 128) == (128, 72, 0, 12, 128, 96)
    for (h, w, S) in [(128, 58, 128), (58, 128, 128), (80, 108, 416), (396, 416, 416), (48, 76, 128), (1, 40, 128)]:
        nh, nw, top, left, oh, ow = LB.letterbox_geometry(h, w, S)
        assert oh % 32 == 0 and ow % 32 == 0 and oh <= S and ow <= S and nh <= oh and nw <= ow
        x = np.random.default_rng(h * w).integers(0, 256, (h, w, 3), dtype=np.uint8)
        assert np.array_equal(LB.letterbox_u8(x, S, use_cv2=True), LB.letterbox_u8(x, S, use_cv2=False))
        p = LB.preprocess(x, S)
        assert p.shape == (3, oh, ow) and p.dtype == np.float32 and p.max() <= 1.0
