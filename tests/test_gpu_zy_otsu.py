"""GPU parity of the Otsu binarisation branch of the DT-Edge builder (DT_BIN_METHOD = "otsu", Detect_OBB.py:109-111;
Train_OBB.py:633-635) against the reference's own vectors (tests/golden/otsu_golden.npz, written by the lifted
reference) and the oracle, bit-exact per stage.  The file sorts last on purpose: the branch was added after the
percentile path (the reference default) and must never mask a regression there under `pytest -x`."""
import numpy as np
import pytest
import torch

from oracle import pixel as P

pytestmark = pytest.mark.gpu

KEYS = ["t1_416_ragged", "t1_128_full", "t1_128_noedge", "t1_128_sliver", "t2_416_crop", "t2_128_ragged",
        "const_5x7", "row_1x40", "col_33x1"]


def _tiles(plan):
    for t in plan.tiles:
        yield tuple(int(t[k]) for k in ("y0", "x0", "h", "w", "px_off"))


@pytest.mark.parametrize("key", KEYS)
def test_otsu_build_multich_matches_reference_vectors(cuda_dev, pixel_golden, otsu_golden, monkeypatch, key):
    from oriented_object_detection_b200 import detect
    monkeypatch.setattr(detect, "DT_BIN_METHOD", "otsu")
    got = detect.build_multich(pixel_golden["in_" + key], 4)
    assert np.array_equal(got, otsu_golden["out_" + key])
    monkeypatch.setattr(detect, "DT_BIN_METHOD", "anything else")        # the reference's `else:` branch is the percentile one
    assert np.array_equal(detect.build_multich(pixel_golden["in_" + key], 4), pixel_golden["out_" + key])


@pytest.mark.parametrize("H,W,ts,ov", [(700, 820, 128, 30), (807, 895, 416, 100), (300, 300, 256, 64)])
def test_otsu_stage_by_stage(cuda_dev, H, W, ts, ov):
    from oriented_object_detection_b200 import _lib, ops, synth
    img = synth.synthetic_map_numpy(H, W, seed=H + W + 1)
    m = torch.from_numpy(img).to(cuda_dev)
    plan = ops.make_plan(H, W, ts, ov, device=cuda_dev)
    t4 = ops.dtedge_build(m, plan, _lib.make_params(flags=_lib.GM_DTEDGE_OTSU)).cpu().numpy()
    dbg = ops.dtedge_debug_views(plan, cuda_dev)
    for ti, (y0, x0, h, w, off) in enumerate(_tiles(plan)):
        crop = img[y0:y0 + h, x0:x0 + w]
        st = P.dt_edge_stages(crop, bin_method="otsu")
        assert np.array_equal(dbg["zero"][ti], st["opened"]), f"opened edge map tile {ti}"
        assert np.array_equal(dbg["t"][ti], st["t"]), f"chamfer tile {ti}"
        got = t4[4 * off:4 * (off + h * w)].reshape(h, w, 4)
        assert np.array_equal(got[..., :3], crop[..., ::-1]), f"RGB planes tile {ti}"
        assert np.array_equal(got[..., 3], st["dt_edge"]), f"DT-Edge plane tile {ti}"


def test_otsu_degenerate_tiles_and_train_twin(cuda_dev, pixel_golden, otsu_golden):
    from oriented_object_detection_b200 import detect
    rng = np.random.default_rng(1)
    cases = [np.full((1, 1, 3), 7, np.uint8), np.full((40, 40, 3), 13, np.uint8)]
    for shp in [(1, 9), (9, 1), (3, 7), (2, 30), (1, 416), (416, 1), (23, 13), (7, 7)]:
        cases.append(rng.integers(0, 256, (shp[0], shp[1], 3), dtype=np.uint8))
    two = np.full((64, 64, 3), 230, np.uint8)
    two[:, 32:] = 20
    cases += [two, rng.integers(0, 256, (128, 128, 3), dtype=np.uint8)]
    for c in cases:
        chw = detect.build_4ch_CHW_from_bgr_dtedge(c, sigmas=(0, 0.6, 1.2, 2.4), bin_method="otsu")
        assert np.array_equal(chw.transpose(1, 2, 0), P.build_multich(c, 4, bin_method="otsu")), c.shape
    plane = detect.dt_edge_channel_from_bgr(pixel_golden["in_t1_128_full"], bin_method="otsu")
    assert np.array_equal(plane, otsu_golden["out_t1_128_full"][..., 3])
