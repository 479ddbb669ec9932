"""GPU parity of the pixel path (tile plan, 3-ch gather, DT-Edge builder) against the oracle and
the reference's golden vectors.  Integer stages are compared bit-exactly, stage by stage."""
import numpy as np
import pytest
import torch

from oracle import geometry as G
from oracle import pixel as P

pytestmark = pytest.mark.gpu

KEYS = ["t1_416_ragged", "t1_128_full", "t1_128_noedge", "t1_128_sliver", "t2_416_crop", "t2_128_ragged",
        "const_5x7", "row_1x40", "col_33x1"]


def _tiles(plan):
    for t in plan.tiles:
        yield tuple(int(t[k]) for k in ("y0", "x0", "h", "w", "px_off"))


@pytest.mark.parametrize("key", KEYS)
def test_build_multich_host_api_matches_reference_vectors(cuda_dev, pixel_golden, key):
    from oriented_object_detection_b200 import detect
    crop = pixel_golden["in_" + key]
    got4 = detect.build_multich(crop, 4)
    assert got4.dtype == np.uint8 and got4.flags["C_CONTIGUOUS"] and got4.shape == crop.shape[:2] + (4,)
    assert np.array_equal(got4, pixel_golden["out_" + key])
    got3 = detect.build_multich(crop, 3)
    assert np.array_equal(got3, crop) and got3.flags["C_CONTIGUOUS"]
    with pytest.raises(AssertionError):
        detect.build_multich(crop, 5)


def test_train_twin_chw(cuda_dev, pixel_golden):
    from oriented_object_detection_b200 import detect
    a = detect.build_4ch_CHW_from_bgr_dtedge(pixel_golden["in_t1_128_full"], sigmas=(0, 0.6, 1.2, 2.4))
    assert np.array_equal(a, pixel_golden["chw_t1_128_full"])
    b = detect.build_4ch_CHW_from_bgr_dtedge(pixel_golden["in_t2_128_ragged"])
    assert np.array_equal(b, pixel_golden["chw_default_sigmas_t2_128_ragged"])
    c = detect.dt_edge_channel_from_bgr(pixel_golden["in_t1_128_full"])
    assert np.array_equal(c, pixel_golden["chw_t1_128_full"][3])


@pytest.mark.parametrize("H,W,ts,ov", [(700, 820, 128, 30), (807, 895, 416, 100), (430, 417, 416, 100),
                                       (257, 129, 128, 30), (64, 1000, 416, 100), (300, 300, 256, 64)])
def test_tiler_and_dtedge_stage_by_stage(cuda_dev, H, W, ts, ov):
    from oriented_object_detection_b200 import ops, synth
    img = synth.synthetic_map_numpy(H, W, seed=H + W)
    m = torch.from_numpy(img).to(cuda_dev)
    plan = ops.make_plan(H, W, ts, ov, device=cuda_dev)
    assert [(y, x, h, w) for y, x, h, w, _ in _tiles(plan)] == G.tile_plan(H, W, ts, ov)
    t3 = ops.tile_gather(m, plan).cpu().numpy()
    t4 = ops.dtedge_build(m, plan).cpu().numpy()
    dbg = ops.dtedge_debug_views(plan, cuda_dev)
    for ti, (y0, x0, h, w, off) in enumerate(_tiles(plan)):
        crop = img[y0:y0 + h, x0:x0 + w]
        assert np.array_equal(t3[3 * off:3 * (off + h * w)].reshape(h, w, 3), crop), f"3-ch tile {ti}"
        st = P.dt_edge_stages(crop)
        assert np.array_equal(dbg["S"][ti].astype(np.int64), st["S"]), f"S tile {ti}"
        assert np.array_equal(dbg["zero"][ti], st["opened"]), f"opened edge map tile {ti}"
        assert np.array_equal(dbg["t"][ti], st["t"]), f"chamfer tile {ti}"
        got = t4[4 * off:4 * (off + h * w)].reshape(h, w, 4)
        assert np.array_equal(got[..., :3], crop[..., ::-1]), f"RGB planes tile {ti}"
        assert np.array_equal(got[..., 3], st["dt_edge"]), f"DT-Edge plane tile {ti}"


def test_degenerate_tiles(cuda_dev):
    from oriented_object_detection_b200 import detect
    rng = np.random.default_rng(0)
    cases = [np.full((1, 1, 3), 7, np.uint8), np.full((2, 2, 3), 200, np.uint8), np.full((40, 40, 3), 13, np.uint8)]
    for shp in [(1, 9), (9, 1), (3, 7), (2, 30), (30, 2), (5, 5), (1, 416), (416, 1), (23, 13), (4, 4), (16, 3), (7, 7), (13, 2)]:
        cases.append(rng.integers(0, 256, (shp[0], shp[1], 3), dtype=np.uint8))
    cb = (np.indices((32, 32)).sum(0) % 2 * 255).astype(np.uint8)
    cases.append(np.dstack([cb] * 3))
    line = np.full((64, 64, 3), 220, np.uint8); line[30, :] = 0
    cases.append(line)
    cases.append(rng.integers(0, 256, (128, 128, 3), dtype=np.uint8))
    for c in cases:
        assert np.array_equal(detect.build_multich(c, 4), P.build_multich(c, 4)), c.shape


def test_other_sigmas_and_no_open(cuda_dev):
    from oriented_object_detection_b200 import _lib, ops, synth
    img = synth.synthetic_map_numpy(200, 260, seed=3)
    m = torch.from_numpy(img).to(cuda_dev)
    plan = ops.plan_from_tiles(200, 260, [(0, 0, 200, 260), (10, 20, 128, 128)], device=cuda_dev)
    for sigmas, p_hi, mo in (((0, 0.8, 1.6, 3.2), 90, 1), ((1.0,), 80, 0), ((0, 2.0), 97.5, 1)):
        out = ops.dtedge_build(m, plan, _lib.make_params(sigmas, p_hi, mo)).cpu().numpy()
        for (y0, x0, h, w, off) in _tiles(plan):
            crop = img[y0:y0 + h, x0:x0 + w]
            want = P.dt_edge_channel(crop, sigmas, p_hi, mo)
            assert np.array_equal(out[4 * off:4 * (off + h * w)].reshape(h, w, 4)[..., 3], want), (sigmas, p_hi, mo)


@pytest.mark.parametrize("mo", [2, 3, 8])
def test_iterated_open_like_cv2(cuda_dev, mo):
    """DT_MORPH_OPEN = n >= 2: n erosions then n dilations (cv2.morphologyEx iterations, Detect_OBB.py:116-118), on large,
    small (<= 128: the fused small-tile kernels step aside), ragged and 1-px tiles; the opened mask and the final channel
    against the oracle (pinned on cv2 in tests/test_oracle_pixel.py)."""
    from oriented_object_detection_b200 import _lib, ops, synth
    H, W = 500, 620
    img = synth.synthetic_map_numpy(H, W, seed=12)
    rng = np.random.default_rng(mo)
    img[0:200, 0:200] = np.where(rng.random((200, 200, 1)) < 0.5, 20, 230)      # blobs that survive several erosions
    img[0:200, 0:200] = np.repeat(np.repeat(img[0:200:8, 0:200:8], 8, 0), 8, 1)
    tiles = [(0, 0, 416, 416), (0, 0, 200, 200), (30, 40, 128, 128), (100, 300, 97, 130), (250, 100, 33, 260),
             (0, 600, 300, 1), (499, 0, 1, 400), (60, 60, 5, 5)]
    plan = ops.plan_from_tiles(H, W, tiles, device=cuda_dev)
    m = torch.from_numpy(img).to(cuda_dev)
    out = ops.dtedge_build(m, plan, _lib.make_params((0, 0.6, 1.2, 2.4), 90, mo)).cpu().numpy()
    dbg = ops.dtedge_debug_views(plan, cuda_dev)
    kept_any = False
    for ti, (y0, x0, h, w, off) in enumerate(_tiles(plan)):
        st = P.dt_edge_stages(img[y0:y0 + h, x0:x0 + w], (0, 0.6, 1.2, 2.4), 90, mo)
        assert np.array_equal(dbg["zero"][ti], st["opened"]), f"opened edge map tile {ti}"
        assert np.array_equal(out[4 * off:4 * (off + h * w)].reshape(h, w, 4)[..., 3], st["dt_edge"]), f"tile {ti}"
        kept_any |= bool(st["opened"].any())
    assert kept_any or mo == 8


def test_generic_and_specialised_gradient_kernels_agree(cuda_dev):
    """The configured (0, 0.6, 1.2, 2.4) stack runs the IDP kernel; the flag forces the generic one.
    Both must give the oracle's S on map-like data, uniform noise (every byte lane saturates) and
    extreme black/white stripes, for full, ragged and sliver tiles at unaligned map offsets."""
    from oriented_object_detection_b200 import _lib, ops, synth
    rng = np.random.default_rng(5)
    H, W = 333, 471
    imgs = [synth.synthetic_map_numpy(H, W, seed=11), rng.integers(0, 256, (H, W, 3), dtype=np.uint8),
            np.where((np.indices((H, W)).sum(0) // 3 % 2)[..., None] > 0, 255, 0).astype(np.uint8).repeat(3, 2)]
    tiles = [(0, 0, 128, 128), (5, 7, 130, 97), (200, 301, 133, 170), (1, 466, 332, 5), (300, 0, 33, 471),
             (17, 33, 64, 3), (100, 100, 31, 33), (0, 0, 333, 471)]
    plan = ops.plan_from_tiles(H, W, tiles, device=cuda_dev)
    for img in imgs:
        m = torch.from_numpy(img).to(cuda_dev)
        res = []
        for flags in (0, _lib.GM_DTEDGE_GENERIC_GRAD):
            out = ops.dtedge_build(m, plan, _lib.make_params(flags=flags)).cpu().numpy()
            res.append((out, [s.copy() for s in ops.dtedge_debug_views(plan, cuda_dev)["S"]]))
        assert np.array_equal(res[0][0], res[1][0])
        for ti, (y0, x0, h, w, off) in enumerate(_tiles(plan)):
            want = P.dt_edge_stages(img[y0:y0 + h, x0:x0 + w])["S"]
            assert np.array_equal(res[0][1][ti].astype(np.int64), want), f"IDP kernel, tile {ti}"
            assert np.array_equal(res[1][1][ti].astype(np.int64), want), f"generic kernel, tile {ti}"


def test_full_size_map_sampled_tiles_and_properties(cuda_dev):
    """BASELINE config 3 size: 8192^2, 416/100 -> 676 tiles; sampled tiles against the oracle plus
    size-independent properties (RGB planes == gathered BGR reversed; every tile written)."""
    from oriented_object_detection_b200 import ops, synth
    H = W = 8192
    m = synth.synthetic_map(H, W, seed=1000, device=cuda_dev)
    plan = ops.make_plan(H, W, 416, 100, device=cuda_dev)
    assert plan.n == 676
    t3 = ops.tile_gather(m, plan)
    out = torch.full((4 * plan.total_px,), 0xAB, dtype=torch.uint8, device=cuda_dev)
    ops.dtedge_build(m, plan, out=out)
    o4 = out.view(-1, 4)
    assert torch.equal(o4[:, :3].flip(1), t3.view(-1, 3))
    rng = np.random.default_rng(1)
    picks = [0, 25, 675, 650] + rng.integers(0, 676, 4).tolist()
    for ti in picks:
        y0, x0, h, w, off = (int(plan.tiles[ti][k]) for k in ("y0", "x0", "h", "w", "px_off"))
        crop = m[y0:y0 + h, x0:x0 + w].cpu().numpy()
        got = o4[off:off + h * w].cpu().numpy().reshape(h, w, 4)
        assert np.array_equal(got, P.build_multich(crop, 4)), f"tile {ti}"
    # CPU and GPU synthetic generators agree (integer-only hash)
    assert np.array_equal(m[4000:4100, 5000:5200].cpu().numpy(),
                          synth.synthetic_map(H, W, 1000, "cpu", row0=4000, rows=100).numpy()[:, 5000:5200])


@pytest.mark.parametrize("channels,chunks", [(4, 5), (3, 3), (4, 1), (4, 64)])
def test_streamed_host_build_equals_resident_build(cuda_dev, channels, chunks):
    """Chunked upload overlapped with the build (ops.build_tiles_from_host) gives the bytes of the
    one-shot build on a resident map, for both channel counts and any chunk count."""
    from oriented_object_detection_b200 import ops, synth
    H, W = 1500, 1100
    img = synth.synthetic_map_numpy(H, W, seed=21)
    h_map = torch.from_numpy(img).pin_memory()
    plan = ops.make_plan(H, W, 416, 100, device=cuda_dev)
    m = torch.from_numpy(img).to(cuda_dev)
    want = ops.dtedge_build(m, plan) if channels == 4 else ops.tile_gather(m, plan)
    want = want.clone()
    got, m2 = ops.build_tiles_from_host(h_map, plan, channels, n_chunks=chunks)
    torch.cuda.synchronize()
    assert torch.equal(m2, m)
    assert torch.equal(got, want)


def test_sampled_selection_equals_two_pass_selection(cuda_dev, monkeypatch):
    """k_select_* first try a sampled bracket + one counting pass and fall back to the two-pass selection when
    the counts do not prove the bracket (GM_SELECT_SAMPLED=0 forces the two-pass path).  Same thresholds and
    bytes either way, on map-like tiles, noise, flat tiles (all keys equal), two-level tiles (heavy ties: the
    bracket overflows the list) and a gradient ramp; the oracle pins a few of them."""
    from oriented_object_detection_b200 import ops, synth
    rng = np.random.default_rng(9)
    H, W = 1300, 1700
    img = synth.synthetic_map_numpy(H, W, seed=77)
    img[0:416, 0:416] = 131                                              # flat
    img[0:416, 416:832] = np.where(rng.random((416, 416, 1)) < 0.03, 0, 255)   # sparse dots on white
    img[416:832, 0:416] = rng.integers(0, 256, (416, 416, 3), dtype=np.uint8)  # noise
    img[416:832, 416:832] = (np.arange(416) * 255 // 415).astype(np.uint8)[None, :, None]   # ramp
    img[832:1248, 0:416] = np.where((np.indices((416, 416))[1] // 52 % 2)[..., None] > 0, 250, 5)  # bars
    tiles = [(0, 0, 416, 416), (0, 416, 416, 416), (416, 0, 416, 416), (416, 416, 416, 416), (832, 0, 416, 416),
             (300, 900, 416, 416), (800, 1200, 416, 416), (10, 1000, 200, 300), (700, 700, 181, 182), (0, 0, 1024, 1024)]
    plan = ops.plan_from_tiles(H, W, tiles, device=cuda_dev)
    m = torch.from_numpy(img).to(cuda_dev)
    res = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("GM_SELECT_SAMPLED", mode)
        out = ops.dtedge_build(m, plan).cpu().numpy()
        dbg = ops.dtedge_debug_views(plan, cuda_dev)
        res[mode] = (out, [z.copy() for z in dbg["zero"]], [t.copy() for t in dbg["t"]])
    assert np.array_equal(res["0"][0], res["1"][0])
    for a, b in zip(res["0"][1], res["1"][1]):
        assert np.array_equal(a, b)
    for ti in (0, 1, 3, 5, 8):
        y0, x0, h, w, off = (int(plan.tiles[ti][k]) for k in ("y0", "x0", "h", "w", "px_off"))
        want = P.build_multich(img[y0:y0 + h, x0:x0 + w], 4)
        assert np.array_equal(res["1"][0][4 * off:4 * (off + h * w)].reshape(h, w, 4), want), f"tile {ti}"


@pytest.mark.parametrize("layout", [0, 1])
def test_small_tile_fused_kernels_equal_the_large_tile_kernels(cuda_dev, monkeypatch, layout):
    """Plans whose tiles are at most 128 px on a side take k_select_edge_small / k_select_tail_small (selection fused
    with its consumer, keys in shared memory; GM_SMALL_FUSED=0 keeps the large-tile kernels).  Same opened edge map,
    same chamfer field, same bytes - on map-like tiles, flat / two-level / noise / ramp tiles, ragged and 1-px tiles,
    a tile without any edge pixel after the open - and equal to the oracle on a sample."""
    from oriented_object_detection_b200 import detect, ops, synth
    rng = np.random.default_rng(4)
    H, W = 560, 700
    img = synth.synthetic_map_numpy(H, W, seed=31)
    img[0:128, 0:128] = 77                                                     # flat: every pixel an edge, dist 0
    img[0:128, 128:256] = np.where(rng.random((128, 128, 1)) < 0.03, 0, 255)   # sparse dots: the open deletes them
    img[128:256, 0:128] = rng.integers(0, 256, (128, 128, 3), dtype=np.uint8)  # noise
    img[128:256, 128:256] = (np.arange(128) * 2).astype(np.uint8)[None, :, None]   # ramp
    img[256:384, 0:128] = np.where((np.indices((128, 128))[1] // 16 % 2)[..., None] > 0, 250, 5)   # bars: heavy ties
    tiles = [(0, 0, 128, 128), (0, 128, 128, 128), (128, 0, 128, 128), (128, 128, 128, 128), (256, 0, 128, 128),
             (300, 300, 128, 128), (400, 500, 128, 97), (431, 571, 23, 13), (100, 650, 1, 50), (200, 699, 128, 1),
             (10, 10, 1, 1), (5, 600, 3, 7), (250, 250, 127, 126), (50, 400, 64, 128)]
    plan = ops.plan_from_tiles(H, W, tiles, device=cuda_dev)
    assert plan.max_tile <= 128
    m = torch.from_numpy(img).to(cuda_dev)
    params = detect._params(layout=layout)
    res = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("GM_SMALL_FUSED", mode)
        out = ops.dtedge_build(m, plan, params).cpu().numpy()
        dbg = ops.dtedge_debug_views(plan, cuda_dev)
        res[mode] = (out, [z.copy() for z in dbg["zero"]], [t.copy() for t in dbg["t"]])
    assert np.array_equal(res["0"][0], res["1"][0])
    for a, b in zip(res["0"][1], res["1"][1]):
        assert np.array_equal(a, b)
    for a, b in zip(res["0"][2], res["1"][2]):
        assert np.array_equal(a, b)
    if layout == 0:
        for ti in (0, 1, 3, 4, 5, 6, 7, 8, 9, 10, 11):
            y0, x0, h, w, off = (int(plan.tiles[ti][k]) for k in ("y0", "x0", "h", "w", "px_off"))
            want = P.build_multich(img[y0:y0 + h, x0:x0 + w], 4)
            assert np.array_equal(res["1"][0][4 * off:4 * (off + h * w)].reshape(h, w, 4), want), f"tile {ti}"
        # DT_MORPH_OPEN = 0: the fused kernel writes the raw edge bits (its own branch)
        from oriented_object_detection_b200 import _lib
        p0 = _lib.make_params((0, 0.6, 1.2, 2.4), 90, 0)
        got = {}
        for mode in ("0", "1"):
            monkeypatch.setenv("GM_SMALL_FUSED", mode)
            got[mode] = ops.dtedge_build(m, plan, p0).cpu().numpy()
        assert np.array_equal(got["0"], got["1"])
        for ti in (1, 5, 6, 7):
            y0, x0, h, w, off = (int(plan.tiles[ti][k]) for k in ("y0", "x0", "h", "w", "px_off"))
            want = P.dt_edge_channel(img[y0:y0 + h, x0:x0 + w], (0, 0.6, 1.2, 2.4), 90, 0)
            assert np.array_equal(got["1"][4 * off:4 * (off + h * w)].reshape(h, w, 4)[..., 3], want), f"no-open tile {ti}"


@pytest.mark.parametrize("chunks,streams", [(1, 1), (3, 2), (4, 4), (7, 8)])
def test_forked_build_equals_single_stream_build(cuda_dev, monkeypatch, chunks, streams):
    """gm_dtedge_build_u8 forks tile ranges onto internal side streams (GM_DTEDGE_CHUNKS x GM_DTEDGE_STREAMS) and
    joins the caller's stream again: same bytes, and work queued behind it on the caller's stream sees them."""
    from oriented_object_detection_b200 import ops, synth
    H, W = 3000, 3200
    m = synth.synthetic_map(H, W, seed=5, device=cuda_dev)
    plan = ops.make_plan(H, W, 416, 100, device=cuda_dev)
    assert plan.n >= 64
    monkeypatch.setenv("GM_DTEDGE_CHUNKS", "1")
    monkeypatch.setenv("GM_DTEDGE_STREAMS", "1")
    want = ops.dtedge_build(m, plan).clone()
    monkeypatch.setenv("GM_DTEDGE_CHUNKS", str(chunks))
    monkeypatch.setenv("GM_DTEDGE_STREAMS", str(streams))
    out = torch.zeros_like(want)
    s = torch.cuda.Stream(device=cuda_dev)
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        ops.dtedge_build(m, plan, out=out)
        got = out.clone()                      # stream-ordered after the join
    s.synchronize()
    assert torch.equal(got, want)
