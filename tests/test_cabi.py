"""CPU-side checks of the C ABI: the library builds for sm_100a, loads, exports every symbol the
header declares, and its host-only entry points (tile plan, status strings, argument validation)
behave.  No kernel is launched here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle import geometry as G

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "geomap_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gm_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound(built_lib):
    from oriented_object_detection_b200 import _lib
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(_lib.lib, n), f"{n} declared in the header but not exported"
        assert n in _lib.PROTOTYPES, f"{n} has no ctypes prototype"
    assert sorted(_lib.PROTOTYPES) == names
    assert _lib.lib.gm_version() == 100
    assert _lib.lib.gm_status_string(0) == b"ok"
    assert b"invalid" in _lib.lib.gm_status_string(-1)


def test_struct_layouts_match_header(built_lib):
    from oriented_object_detection_b200 import _lib, ops
    assert C.sizeof(_lib.gm_tile) == 24 and ops.TILE_DTYPE.itemsize == 24
    assert C.sizeof(_lib.gm_dtedge_params) == 8 * 8 + 8 + 16


@pytest.mark.parametrize("H,W,ts,ov", [(807, 895, 416, 100), (807, 895, 128, 30), (1028, 1056, 416, 100),
                                       (8192, 8192, 416, 100), (16384, 16384, 128, 30), (5, 7, 416, 100),
                                       (417, 317, 416, 100), (300, 300, 128, 128), (129, 1, 128, 30)])
def test_tile_plan_matches_oracle(built_lib, H, W, ts, ov):
    from oriented_object_detection_b200 import ops
    plan = ops.make_plan(H, W, ts, ov)
    want = G.tile_plan(H, W, ts, ov)
    got = [(int(t["y0"]), int(t["x0"]), int(t["h"]), int(t["w"])) for t in plan.tiles]
    assert got == want
    offs = np.concatenate([[0], np.cumsum([h * w for _, _, h, w in want])])
    assert plan.tiles["px_off"].tolist() == offs[:-1].tolist()
    assert plan.total_px == offs[-1]
    assert plan.rows * plan.cols == len(want)


def test_tile_plan_known_counts(built_lib):
    from oriented_object_detection_b200 import ops
    # SURVEY.md Appendix D
    assert ops.make_plan(8192, 8192, 416, 100).n == 676
    assert ops.make_plan(8192, 8192, 416, 100).total_px == 114_318_864
    assert ops.make_plan(16384, 16384, 128, 30).n == 28224
    assert ops.make_plan(16384, 16384, 416, 100).n == 2704


def test_row_band_plans_partition_the_full_plan(built_lib):
    from oriented_object_detection_b200 import ops
    full = ops.make_plan(3000, 2100, 416, 100)
    got = []
    for r0, r1 in ((0, 3), (3, 7), (7, full.rows)):
        band = ops.make_plan(3000, 2100, 416, 100, r0, r1)
        assert band.tiles["px_off"][0] == 0
        got += [(int(t["y0"]), int(t["x0"]), int(t["h"]), int(t["w"])) for t in band.tiles]
    assert got == [(int(t["y0"]), int(t["x0"]), int(t["h"]), int(t["w"])) for t in full.tiles]


def test_argument_validation_without_gpu(built_lib):
    from oriented_object_detection_b200 import _lib
    L = _lib.lib
    assert L.gm_tile_plan_count(0, 10, 416, 100, None, None, None) == -1
    assert L.gm_tile_gather_u8(None, 10, 10, None, 1, 416, None, None) == -1
    assert L.gm_dtedge_workspace_bytes(1000, 3) > 9000
    assert L.gm_nms_workspace_bytes(1000, 0) > 0 and L.gm_fuse_workspace_bytes(1000, 0) > L.gm_nms_workspace_bytes(1000, 0)
    cnt = C.c_int64(7)
    assert L.gm_nms_global(None, None, None, -1, 0, 0.4, 0, None, None, None, C.byref(cnt), None, 0, None) == -1


def test_package_refuses_to_run_without_cuda(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from oriented_object_detection_b200 import detect, ops
    with pytest.raises(RuntimeError):
        ops.tile_gather(torch.zeros((4, 4, 3), dtype=torch.uint8), ops.make_plan(4, 4, 416, 100))
    with pytest.raises(RuntimeError):
        detect.merge_detections([(0, 0, 1, 0, 1, 1, 0, 1, 0, 0.9, 0.0)], 0.4)
