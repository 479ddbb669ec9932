"""Seam-band exchange through the C ABI (gm_band_merge_local / gm_band_merge_finish), N ranks emulated on one GPU: each
emulated rank runs the local phase on its tile range in its own workspace, the records are concatenated the way the
all_gather would, every rank finishes; the union of the kept lists must equal the single-rank merge - members and order
(merge_detections over the whole list, Detect_OBB.py:291, :176-200)."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def _emulate(boxes, cls, conf, tid, H, W, tile, ov, margin, world, max_class, thr, bound, seam_cap, split="tiles", pad=9,
             angle=None, reach=None, stats=None):
    """-> (global rows in the reference's order, per-rank metas).  reach: resolve only the seam blocks of ranks within
    `reach` of each rank (sharding.seam_scope); a rank whose chains leave that range repeats the seam phase on all blocks."""
    import torch
    from oriented_object_detection_b200 import ops, sharding
    import seam_ref as R
    dev = torch.device("cuda:0")
    step = tile - ov
    rows_t, cols_t = -(-H // step), -(-W // step)
    recs, sends, wss, mines, ranges = [], [], [], [], []
    for r in range(world):
        if split == "tiles":
            t0, t1 = sharding.tile_range(rows_t * cols_t, world, r)
        else:
            r0, r1 = sharding.band_rows(rows_t, world, r)
            t0, t1 = r0 * cols_t, r1 * cols_t
        ranges.append((t0, t1))
        mine = np.nonzero((tid >= t0) & (tid < t1))[0]
        rects = sharding.foreign_center_rects(H, W, tile, ov, t0, t1, margin)

        def padded(a, fill):
            return torch.from_numpy(np.concatenate([a[mine], np.full((pad,) + a.shape[1:], fill, dtype=a.dtype)])).to(dev)
        rec = {"boxes": padded(boxes, 5.0), "cls": padded(cls, 0), "conf": padded(conf, 0.999)}
        if angle is not None:
            rec["angle"] = padded(angle, 0.0)
        with ops.workspace_scope(f"emul{r}"):
            send, ws = ops.band_merge_local(rec, torch.tensor([len(mine)], device=dev), max_class, thr, rects, bound, world, seam_cap)
        recs.append(rec); sends.append(send); wss.append(ws); mines.append(mine)
    gathered = torch.cat(sends, 0).contiguous()
    outs, metas = [], []
    for r in range(world):
        if reach is None:
            out = ops.band_merge_finish(gathered, world, r, seam_cap, recs[r], max_class, thr, wss[r])
        else:
            sc = sharding.seam_scope(H, W, tile, ov, margin, world, r, ranges=ranges, reach=reach)
            blocks = sc["blocks"] if tuple(sc["blocks"]) != (0, world) else None
            out = ops.band_merge_finish(gathered, world, r, seam_cap, recs[r], max_class, thr, wss[r], blocks=blocks,
                                        outside_rects=sc["rects"] if blocks else None, extent_bound=bound)
            if int(out["meta"][1].item()) == ops.SEAM_CHAIN_ESCAPES:
                if stats is not None:
                    stats["fallbacks"] = stats.get("fallbacks", 0) + 1
                out = ops.band_merge_finish(gathered, world, r, seam_cap, recs[r], max_class, thr, wss[r], out=out)
        meta = [int(v) for v in out["meta"].tolist()]
        metas.append(meta)
        m = meta[0]
        rows = mines[r][out["src"][:m].cpu().numpy()]
        assert np.array_equal(out["boxes"][:m].cpu().numpy(), boxes[rows]) and np.array_equal(out["cls"][:m].cpu().numpy(), cls[rows])
        if angle is not None:
            assert np.array_equal(out["angle"][:m].cpu().numpy(), angle[rows])
        outs.append((out["conf"][:m].cpu().numpy(), rows))
    return R.merge_rank_outputs(outs), metas


@pytest.mark.parametrize("world,split", [(1, "tiles"), (2, "rows"), (3, "tiles"), (4, "tiles"), (8, "tiles")])
def test_small_map_with_chains_across_every_seam(cuda_dev, world, split):
    import seam_ref as R
    H, W, tile, ov, margin, thr = 1500, 900, 416, 100, 20, 0.4
    boxes, cls, conf, tid, plan = R.make_case(H, W, tile, ov, margin, 260, 3, seed=21)
    want = R.expected(boxes, cls, conf, thr)
    angle = np.arange(len(conf), dtype=np.float64) * 0.25
    got, metas = _emulate(boxes, cls, conf, tid, H, W, tile, ov, margin, world, 2, thr, 55.0, 700, split, angle=angle)
    assert got.tolist() == want.tolist()
    # neighbour-restricted seam phase: this map's chains run through every band, so ranks with absent neighbours must
    # notice (GM_SEAM_CHAIN_ESCAPES) and fall back; the result is the same
    stats = {}
    got_r, metas_r = _emulate(boxes, cls, conf, tid, H, W, tile, ov, margin, world, 2, thr, 55.0, 700, split, angle=angle,
                              reach=1, stats=stats)
    assert got_r.tolist() == want.tolist() and all(m[1] == 0 for m in metas_r)
    fb = stats.get("fallbacks", 0)
    assert (fb > 0 or world < 8) and (fb == 0 or world > 3)
    assert all(m[1] == 0 and m[3] == len(conf) for m in metas) and len({m[2] for m in metas}) == 1
    assert (metas[0][2] == 0) == (world == 1)                          # one rank: nothing is deferred
    # bounds that do not hold come back as status bits on EVERY rank
    if world > 1:
        from oriented_object_detection_b200 import ops
        _, metas = _emulate(boxes, cls, conf, tid, H, W, tile, ov, margin, world, 2, thr, 55.0, 5, split)
        assert all(m[1] & ops.SEAM_CAPACITY_OVERFLOW for m in metas)
        _, metas = _emulate(boxes, cls, conf, tid, H, W, tile, ov, margin, world, 2, thr, 12.0, 700, split)
        assert all(m[1] & ops.SEAM_EXTENT_EXCEEDED for m in metas)


def test_config2_survivors_eight_ranks_equal_single_rank(cuda_dev):
    """BASELINE config 2/5 shape: the per-tile survivors of ~100 k OBBs on the 8192^2 tiling, 8 emulated ranks (tile-range
    bands) and 2 maps batched (class key = map * n_classes + class) against the single-rank gm_nms_global."""
    import torch
    from oriented_object_detection_b200 import ops, synth
    H = W = 8192
    tile, ov, margin, nc, thr = 416, 100, 20, 15, 0.4
    plan = ops.make_plan(H, W, tile, ov, device=cuda_dev)
    per_map = []
    for m in range(2):
        local, cls, conf, tid = synth.synthetic_tile_dets(plan, 59000, nc, seed=m, margin=margin)
        pp = ops.tile_postprocess(torch.from_numpy(local).to(cuda_dev), torch.from_numpy(cls).to(cuda_dev),
                                  torch.from_numpy(conf).to(cuda_dev), torch.from_numpy(tid).to(cuda_dev), plan, margin, 1, thr,
                                  max_class=nc - 1)
        src = pp["src"].cpu().numpy()
        per_map.append((pp["boxes"].cpu().numpy(), pp["cls"].cpu().numpy() + m * nc, pp["conf"].cpu().numpy(), tid[src]))
    # a rank holds its tile range of EVERY map of the batch: list order = (tile range, map, tile, confidence)
    world = 8
    from oriented_object_detection_b200 import sharding
    parts = []
    for r in range(world):
        t0, t1 = sharding.tile_range(plan.n, world, r)
        for (b, c, f, t) in per_map:
            sel = (t >= t0) & (t < t1)
            parts.append((b[sel], c[sel], f[sel], t[sel]))
    boxes, cls, conf, tid = (np.concatenate([p[k] for p in parts]) for k in range(4))
    order, _, kept = ops.nms_global(torch.from_numpy(boxes).to(cuda_dev), torch.from_numpy(cls).to(cuda_dev),
                                    torch.from_numpy(conf).to(cuda_dev), thr, max_class=2 * nc - 1)
    want = kept.cpu().numpy()
    got, metas = _emulate(boxes, cls, conf, tid, H, W, tile, ov, margin, world, 2 * nc - 1, thr, 80.0, 65536)
    assert got.tolist() == want.tolist() and len(want) > 100000
    stats = {}
    got_r, metas_r = _emulate(boxes, cls, conf, tid, H, W, tile, ov, margin, world, 2 * nc - 1, thr, 80.0, 65536, reach=1, stats=stats)
    assert got_r.tolist() == want.tolist() and all(m[1] == 0 for m in metas_r)
    assert stats.get("fallbacks", 0) == 0                     # random objects: no chain of overlaps spans a whole band
    n_seam = metas[0][2]
    assert all(m[1] == 0 for m in metas) and 0 < n_seam < 0.6 * len(conf), (n_seam, len(conf))


def test_agree_seam_bounds_single_rank(cuda_dev):
    """The set-up helper: bound >= the largest box reach, capacity >= the deferred boxes (none on one rank without foreign
    rectangles; all boxes near the rectangle otherwise), and a merge with the agreed values reports status 0."""
    import torch
    import seam_ref as R
    from oriented_object_detection_b200 import ops, sharding
    H, W, tile, ov, margin, thr = 1500, 900, 416, 100, 20, 0.4
    boxes, cls, conf, tid, plan = R.make_case(H, W, tile, ov, margin, 260, 3, seed=5)
    n = len(conf)
    rec = {"boxes": torch.from_numpy(boxes).to(cuda_dev), "cls": torch.from_numpy(cls).to(cuda_dev), "conf": torch.from_numpy(conf).to(cuda_dev)}
    count = torch.tensor([n], device=cuda_dev)
    bound, cap = sharding.agree_seam_bounds(rec, count, thr, 2, [])
    assert float(sharding.box_reach(rec["boxes"]).max()) <= bound < 80 and cap == 1024
    rects = sharding.foreign_center_rects(H, W, tile, ov, 0, 6, margin)          # as if the last 6 tiles belonged to another rank
    bound2, cap2 = sharding.agree_seam_bounds(rec, count, thr, 2, rects)
    assert bound2 == bound and 1024 < cap2 <= int(1.25 * n) + 1024
    out = sharding.merge_bands_seam_finish(sharding.merge_bands_seam_device({k: v.clone() for k, v in rec.items()}, count, cap2, thr, 2, [], bound))
    assert out["boxes"].shape[0] == len(R.expected(boxes, cls, conf, thr))
