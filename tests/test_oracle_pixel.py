"""The numpy restatement of the pixel path against the reference's own outputs (golden vectors,
and the lifted reference itself wherever /root/reference exists)."""
import os

import numpy as np
import pytest

from oracle import geometry as G
from oracle import lift_reference as LR
from oracle import pixel as P

KEYS = ["t1_416_ragged", "t1_128_full", "t1_128_noedge", "t1_128_sliver", "t2_416_crop", "t2_128_ragged",
        "const_5x7", "row_1x40", "col_33x1"]


@pytest.mark.parametrize("key", KEYS)
def test_build_multich_matches_reference_vectors(pixel_golden, key):
    crop = pixel_golden["in_" + key]
    want = pixel_golden["out_" + key]
    got = P.build_multich(crop, 4)
    assert got.dtype == np.uint8 and got.flags["C_CONTIGUOUS"]
    assert np.array_equal(got, want)


def test_three_channel_is_a_contiguous_bgr_copy(pixel_golden):
    crop = pixel_golden["in_t1_128_full"][:, ::-1]           # non-contiguous view
    got = P.build_multich(crop, 3)
    assert got.flags["C_CONTIGUOUS"] and np.array_equal(got, crop)
    with pytest.raises(AssertionError):
        P.build_multich(crop, 5)


def test_train_twin_chw(pixel_golden):
    assert np.array_equal(P.build_4ch_chw(pixel_golden["in_t1_128_full"]), pixel_golden["chw_t1_128_full"])
    assert np.array_equal(P.build_4ch_chw(pixel_golden["in_t2_128_ragged"], (0, 0.8, 1.6, 3.2)),
                          pixel_golden["chw_default_sigmas_t2_128_ragged"])


def test_gaussian_taps_known_values():
    assert P.gaussian_kernel_q8(0.6) == [1, 42, 170, 42, 1]
    assert P.gaussian_kernel_q8(1.2) == [0, 4, 21, 60, 86, 60, 21, 4, 0]
    assert P.gaussian_kernel_q8(2.4) == [1, 1, 5, 11, 19, 31, 39, 42, 39, 31, 19, 11, 5, 1, 1]
    for s in (0.5, 0.8, 1.6, 3.2):
        assert sum(P.gaussian_kernel_q8(s)) == 256


def test_constant_tile_is_178():
    out = P.build_multich(np.full((9, 11, 3), 200, np.uint8), 4)
    assert (out[..., 3] == 178).all()


def test_chamfer_closed_form():
    rng = np.random.default_rng(5)
    z = rng.random((23, 31)) < 0.02
    z[4, 7] = True
    t = P.chamfer_fixed(z)
    ys, xs = np.nonzero(z)
    yy, xx = np.mgrid[0:23, 0:31]
    dx = np.abs(xx[..., None] - xs)
    dy = np.abs(yy[..., None] - ys)
    M, m = np.maximum(dx, dy), np.minimum(dx, dy)
    brute = (P.HV * (M - m) + P.DG * m).min(axis=2)
    assert np.array_equal(t.astype(np.int64), brute)
    assert (P.chamfer_fixed(np.zeros((5, 6), bool)) == P.DIST_MAX).all()


def test_tile_plan_matches_reference_shapes():
    # Test1 807x895
    p = G.tile_plan(807, 895, 416, 100)
    assert len(p) == 9 and p[-1] == (632, 632, 175, 263)
    p = G.tile_plan(807, 895, 128, 30)
    assert len(p) == 90 and p[-1] == (784, 882, 23, 13)
    assert len(G.tile_plan(8192, 8192, 416, 100)) == 676
    assert sum(h * w for _, _, h, w in G.tile_plan(8192, 8192, 416, 100)) == 114_318_864
    assert sum(h * w for _, _, h, w in G.tile_plan(16384, 16384, 416, 100)) == 461_562_256


@pytest.mark.skipif(not LR.fixtures_available(), reason="/root/reference not present (GPU box)")
def test_against_lifted_reference_all_tiles_of_test1():
    import cv2
    cv2.ipp.setUseIPP(False)
    ref = LR.load_detect(4)
    img = cv2.imread(os.path.join(LR.REFERENCE_ROOT, "Input", "Test1.png"))
    bad = 0
    for ts, ov in ((416, 100), (128, 30)):
        for (y, x, h, w) in G.tile_plan(img.shape[0], img.shape[1], ts, ov):
            crop = img[y:y + h, x:x + w]
            bad += int((ref.build_multich(crop, 4) != P.build_multich(crop, 4)).sum())
    assert bad == 0


@pytest.mark.skipif(not LR.fixtures_available(), reason="/root/reference not present (GPU box)")
def test_lifted_detect_symbols_tile_order_and_shapes():
    ref = LR.load_detect(3)
    calls = []
    model = LR.FakeModel(lambda crop, conf: (np.zeros((0, 4, 2)), [], []))
    img = np.zeros((807, 895, 3), np.uint8)
    assert ref.detect_symbols(img, model, 416, 100) == []
    assert [c[0] for c in model.calls] == [(h, w, 3) for (_, _, h, w) in G.tile_plan(807, 895, 416, 100)]
    assert all(c[1] == "uint8" and c[2] and c[3] == 0.25 for c in model.calls)
    del calls


def test_cv_port_equals_primitive_restatement(pixel_golden):
    """The OpenCV port bench.py times as the CPU baseline computes the same bytes (IPP off)."""
    import cv2
    from oracle import pixel_cv
    cv2.ipp.setUseIPP(False)
    for key in KEYS:
        crop = pixel_golden["in_" + key]
        assert np.array_equal(pixel_cv.build_multich(crop, 4), pixel_golden["out_" + key]), key


# ---------------------------------------------------------------- DT_BIN_METHOD = "otsu" (Detect_OBB.py:109-111)

OTSU_KEYS = ["t1_416_ragged", "t1_128_full", "t1_128_noedge", "t2_416_crop", "t2_128_ragged", "const_5x7", "row_1x40"]


@pytest.mark.parametrize("key", OTSU_KEYS)
def test_otsu_build_multich_matches_reference_vectors(pixel_golden, otsu_golden, key):
    """Vectors written by the lifted reference with DT_BIN_METHOD = "otsu" (make_otsu_golden.py)."""
    got = P.build_multich(pixel_golden["in_" + key], 4, bin_method="otsu")
    assert np.array_equal(got, otsu_golden["out_" + key])
    assert int(P.dt_edge_stages(pixel_golden["in_" + key], bin_method="otsu")["hi"]) == int(otsu_golden["thr_" + key])


def test_otsu_threshold_and_normalize_equal_cv2(pixel_golden):
    """The two library calls of the Otsu branch, restated, against cv2 itself (IPP off) on every golden crop
    and on synthetic histograms (bimodal, single-valued, two-valued, ramp)."""
    import cv2
    cv2.ipp.setUseIPP(False)
    rng = np.random.default_rng(5)
    imgs = [rng.integers(0, 256, (64, 80), dtype=np.uint8), np.full((9, 9), 200, np.uint8),
            np.where(rng.random((50, 50)) < 0.3, 10, 240).astype(np.uint8),
            np.tile(np.arange(256, dtype=np.uint8), (4, 1)),
            np.clip(np.concatenate([rng.normal(60, 10, 3000), rng.normal(180, 25, 1000)]), 0, 255).astype(np.uint8).reshape(50, 80)]
    for key in KEYS:
        crop = pixel_golden["in_" + key]
        acc = P.acc_from_S(P.max_scharr_sq(P.gray_u8(crop)))
        n255 = P.normalize_minmax(acc, 0.0, 255.0)
        assert np.array_equal(n255.view(np.uint32), cv2.normalize(acc, None, 0, 255, cv2.NORM_MINMAX).view(np.uint32)), key
        assert np.array_equal(P.acc8_from_acc(acc), n255.astype(np.uint8))
        imgs.append(P.acc8_from_acc(acc))
    for im in imgs:
        thr, mask = cv2.threshold(im, 0, 255, cv2.THRESH_BINARY + cv2.THRESH_OTSU)
        assert P.otsu_threshold_u8(im) == int(thr)
        assert np.array_equal(mask > 0, im > int(thr))


@pytest.mark.skipif(not LR.fixtures_available(), reason="/root/reference not present (GPU box)")
def test_otsu_against_lifted_reference_all_tiles_of_test2():
    import cv2
    cv2.ipp.setUseIPP(False)
    ref = LR.load_detect(4)
    ref.DT_BIN_METHOD = "otsu"
    img = cv2.imread(os.path.join(LR.REFERENCE_ROOT, "Input", "Test2.png"))
    bad = 0
    for ts, ov in ((416, 100), (128, 30)):
        for (y, x, h, w) in G.tile_plan(img.shape[0], img.shape[1], ts, ov):
            crop = img[y:y + h, x:x + w]
            want = ref.build_multich(crop, 4)
            bad += int((want != P.build_multich(crop, 4, bin_method="otsu")).sum())
            from oracle import pixel_cv
            bad += int((want != pixel_cv.build_multich(crop, 4, bin_method="otsu")).sum())
    assert bad == 0


@pytest.mark.skipif(not LR.fixtures_available(), reason="/root/reference not present (GPU box)")
def test_ipp_on_reference_differs_by_at_most_one_level():
    """Parity is defined against OpenCV with IPP off (cv2.magnitude correctly rounded, integer chamfer).  The pip
    wheel's default is IPP ON: the same reference code then differs from the restatement - and so from the CUDA path -
    in a few final-channel pixels by exactly one level.  Reported separately: the bound is
    asserted here, the measured figure on Test1 at both scales is 97 of 2,296,585 px."""
    import cv2
    if not hasattr(cv2, "ipp"):
        pytest.skip("OpenCV without IPP")
    ref = LR.load_detect(4)
    img = cv2.imread(os.path.join(LR.REFERENCE_ROOT, "Input", "Test1.png"))
    total = differ = worst = 0
    try:
        cv2.ipp.setUseIPP(True)
        if not cv2.ipp.useIPP():
            pytest.skip("OpenCV built without IPP")
        for ts, ov in ((416, 100), (128, 30)):
            for (y, x, h, w) in G.tile_plan(img.shape[0], img.shape[1], ts, ov):
                crop = img[y:y + h, x:x + w]
                a = ref.build_multich(crop, 4).astype(np.int16)
                b = P.build_multich(crop, 4).astype(np.int16)
                assert np.array_equal(a[..., :3], b[..., :3])
                d = np.abs(a[..., 3] - b[..., 3])
                total += d.size
                differ += int((d != 0).sum())
                worst = max(worst, int(d.max()))
    finally:
        cv2.ipp.setUseIPP(False)
    assert worst <= 1 and differ < 1e-3 * total


def test_iterated_open_is_n_erosions_then_n_dilations_like_cv2():
    """DT_MORPH_OPEN = n (Detect_OBB.py:116-118): cv2 runs n erosions and then n dilations with the 3x3 cross; applying
    the opening n times would be the opening itself (idempotent).  The oracle follows cv2, borders included."""
    import cv2
    rng = np.random.default_rng(0)
    k = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (3, 3))
    differs_from_single = 0
    for _ in range(150):
        h, w = rng.integers(1, 48, 2)
        m = rng.random((h, w)) < rng.uniform(0.3, 0.97)
        for n in (1, 2, 3):
            want = cv2.morphologyEx(m.astype(np.uint8) * 255, cv2.MORPH_OPEN, k, iterations=n) > 0
            got = P.cross_open(m, n)
            assert np.array_equal(got, want), (h, w, n)
            if n > 1:
                differs_from_single += int(not np.array_equal(got, P.cross_open(m, 1)))
    assert differs_from_single > 50            # the iteration count matters on these masks
