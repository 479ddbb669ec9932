"""Full-size parity of the geometry half (BASELINE configs 2 and 4) against the C restatement of the
reference loops: the per-tile detections of a whole map go through remap / border filter / per-tile NMS,
dual-scale fusion and the global merge on the GPU, and every kept set must equal the oracle's - members
and order."""
import numpy as np
import pytest
import torch

from oracle import geom_c

pytestmark = pytest.mark.gpu


def _dev(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def _tile_stage(ops, synth, H, W, ts, ov, margin, n_obj, n_cls, seed, dev):
    plan = ops.make_plan(H, W, ts, ov, device=dev)
    local, cls, conf, tid = synth.synthetic_tile_dets(plan, n_obj, n_cls, seed=seed, margin=margin)
    out = ops.tile_postprocess(_dev(local, dev), _dev(cls, dev), _dev(conf, dev), _dev(tid, dev), plan, margin, 1, 0.4,
                               max_class=n_cls - 1)
    # oracle: the same per-tile steps on the host (vectorised remap + filter, C NMS per tile group)
    x0 = plan.tiles["x0"][tid].astype(np.float64); y0 = plan.tiles["y0"][tid].astype(np.float64)
    tw = plan.tiles["w"][tid].astype(np.float64); th = plan.tiles["h"][tid].astype(np.float64)
    g = local.astype(np.float64)
    g[:, 0::2] += x0[:, None]; g[:, 1::2] += y0[:, None]
    cx = (g[:, 0] + g[:, 2] + g[:, 4] + g[:, 6]) / 4.0 - x0
    cy = (g[:, 1] + g[:, 3] + g[:, 5] + g[:, 7]) / 4.0 - y0
    ok = (cx >= margin) & (cx <= tw - margin) & (cy >= margin) & (cy <= th - margin)
    # per-tile NMS == one NMS whose class key is (tile, class): tiles never interact
    key = (tid.astype(np.int64) * n_cls + cls).astype(np.int32)
    idx = np.nonzero(ok)[0]
    _, kept = geom_c.nms(g[idx], key[idx], conf[idx], 0.4, grid=True)
    kept = idx[kept]
    # reference list order: tiles row-major, inside a tile confidence-descending (stable)
    order = np.lexsort((np.arange(len(kept)), -conf[kept].astype(np.float64), tid[kept]))
    want = kept[order]
    assert out["src"].cpu().numpy().tolist() == want.tolist()
    assert np.array_equal(out["boxes"].cpu().numpy(), g[want])
    return out, g[want], cls[want], conf[want]


def test_config2_100k_boxes_tile_nms_and_global_merge(cuda_dev):
    from oriented_object_detection_b200 import ops, synth
    out, boxes, cls, conf = _tile_stage(ops, synth, 8192, 8192, 416, 100, 20, 59000, 15, 0, cuda_dev)
    assert 75000 < len(conf) < 110000
    order, keep, kept = ops.nms_global(out["boxes"], out["cls"], out["conf"], 0.4, max_class=14)
    o, k = geom_c.nms(boxes, cls, conf, 0.4, grid=True)
    assert order.cpu().numpy().tolist() == o.tolist()
    assert kept.cpu().numpy().tolist() == k.tolist()
    assert 0.5 * len(conf) < len(k) < 0.9 * len(conf)          # the seam duplicates are what gets removed


def test_config4_dual_scale_fusion_one_million_boxes(cuda_dev):
    """16384^2 map, 128/30 (28,224 tiles) + 416/100 (2,704 tiles), ~1 M candidate boxes."""
    from oriented_object_detection_b200 import ops, synth
    H = W = 16384
    n_obj = 400000
    a_out, a_boxes, a_cls, a_conf = _tile_stage(ops, synth, H, W, 128, 30, 10, n_obj, 15, 1, cuda_dev)
    b_out, b_boxes, b_cls, b_conf = _tile_stage(ops, synth, H, W, 416, 100, 20, n_obj, 15, 1, cuda_dev)
    n_a, n_b = len(a_conf), len(b_conf)
    assert n_a + n_b > 950000
    boxes = torch.cat([a_out["boxes"], b_out["boxes"]]); cls = torch.cat([a_out["cls"], b_out["cls"]])
    conf = torch.cat([a_out["conf"], b_out["conf"]])
    sid = torch.cat([torch.zeros(n_a, dtype=torch.int32), torch.ones(n_b, dtype=torch.int32)]).to(cuda_dev)
    fused = ops.fuse_scales(boxes, cls, conf, sid, 2, max_class=14)
    hb = np.concatenate([a_boxes, b_boxes]); hc = np.concatenate([a_cls, b_cls]); hf = np.concatenate([a_conf, b_conf])
    hs = np.concatenate([np.zeros(n_a, np.int32), np.ones(n_b, np.int32)])
    want_f = geom_c.fuse(hb, hc, hf, hs, 2, grid=True)
    assert fused.cpu().numpy().tolist() == want_f.tolist()
    fi = fused.to(torch.int64)
    order, keep, kept = ops.nms_global(boxes[fi], cls[fi], conf[fi], 0.4, max_class=14)
    o, k = geom_c.nms(hb[want_f], hc[want_f], hf[want_f], 0.4, grid=True)
    assert kept.cpu().numpy().tolist() == k.tolist()
    assert len(k) < len(want_f) < n_a + n_b
