"""Threshold-adjacent pairs are counted and reported (gm_threshold_adjacent_stats): the keep-sets stay the float64
oracle's, the counters say how many decisions needed float64 and how many sit within 1e-5 of the threshold."""
import numpy as np
import pytest
import torch

from oracle import geometry as G

pytestmark = pytest.mark.gpu


def _sq(x, y, s=10.0):
    return [x, y, x + s, y, x + s, y + s, x, y + s]


def test_threshold_adjacent_pairs_are_counted(cuda_dev):
    from oriented_object_detection_b200 import ops
    # two equal squares shifted by d have IoU (100 - 10 d) / (100 + 10 d): exactly 0.4 at d = 30 / 7
    d = 30.0 / 7.0
    boxes = np.array([_sq(0, 0), _sq(d - 3e-5, 0),              # IoU = 0.4 + 2.9e-6: float64 decides, within 1e-5
                      _sq(100, 0), _sq(100 + d + 2e-4, 0),      # IoU = 0.4 - 2.8e-5: float64 path, not within 1e-5
                      _sq(200, 0), _sq(201, 0),                 # IoU 0.82: fp32 decides
                      _sq(300, 0), _sq(309, 0)], np.float64)    # IoU 0.05: fp32 decides
    cls = np.zeros(8, np.int32)
    conf = np.linspace(0.9, 0.2, 8).astype(np.float32)
    ops.threshold_adjacent_stats(reset=True)
    assert ops.threshold_adjacent_stats() == {"float64_decided": 0, "within_1e-5": 0}
    order, keep, kept = ops.nms_global(torch.from_numpy(boxes).to(cuda_dev), torch.from_numpy(cls).to(cuda_dev),
                                       torch.from_numpy(conf).to(cuda_dev), 0.4, max_class=0)
    assert kept.cpu().tolist() == G.nms_keep_indices(boxes, cls, conf, 0.4)
    st = ops.threshold_adjacent_stats(reset=True)
    assert st == {"float64_decided": 2, "within_1e-5": 1}
    assert ops.threshold_adjacent_stats() == {"float64_decided": 0, "within_1e-5": 0}


def test_counts_on_a_synthetic_detection_set(cuda_dev):
    from oriented_object_detection_b200 import ops, synth
    boxes, cls, conf = synth.synthetic_obbs(6000, 1200, 1200, n_classes=3, seed=4)      # dense: 14 k pairs reach 0.4, 15 within 1e-4
    ops.threshold_adjacent_stats(reset=True)
    ops.nms_global(torch.from_numpy(boxes).to(cuda_dev), torch.from_numpy(cls).to(cuda_dev),
                   torch.from_numpy(conf).to(cuda_dev), 0.4, max_class=2)
    st = ops.threshold_adjacent_stats()
    # oracle count of same-class pairs within 1e-5 of the threshold (O(n^2) over AABB-overlapping pairs only)
    n = len(conf)
    lo = boxes.reshape(n, 4, 2).min(1); hi = boxes.reshape(n, 4, 2).max(1)
    near = 0
    wide = 0
    for c in range(3):
        idx = np.nonzero(cls == c)[0]
        for a in range(len(idx)):
            i = idx[a]
            cand = idx[a + 1:]
            ov = (lo[cand, 0] <= hi[i, 0]) & (lo[i, 0] <= hi[cand, 0]) & (lo[cand, 1] <= hi[i, 1]) & (lo[i, 1] <= hi[cand, 1])
            for j in cand[ov]:
                v = G.quad_iou(boxes[i], boxes[j])
                near += abs(v - 0.4) < 1e-5
                wide += abs(v - 0.4) < 1e-4 - 6e-6          # certainly inside the fp32 window
    assert near >= 1 and wide >= 10
    assert st["within_1e-5"] == near
    assert st["float64_decided"] >= wide and st["float64_decided"] >= st["within_1e-5"]
