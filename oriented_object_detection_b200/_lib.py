"""ctypes binding of ``libgeomap_b200.so`` (the C ABI declared in ``include/geomap_b200.h``).

There is no CPU fallback: importing this module without the built library raises, and
every compute call on a machine without an sm_100 GPU fails with the library's status.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# GM_LIB_PATH: a differently tuned build of the same sources (scripts/build_variant.sh); tuning only
LIB_PATH = os.environ.get("GM_LIB_PATH") or os.path.join(_HERE, "lib", "libgeomap_b200.so")

GM_OK = 0
GM_MAX_SCALES = 8
GM_MAX_TILE = 1024


class GmError(RuntimeError):
    def __init__(self, status: int, where: str):
        self.status = status
        msg = lib.gm_status_string(status).decode() if lib is not None else "?"
        super().__init__(f"{where}: status {status} ({msg})")


class gm_tile(C.Structure):
    _fields_ = [("y0", C.c_int32), ("x0", C.c_int32), ("h", C.c_int32), ("w", C.c_int32),
                ("px_off", C.c_int64)]


class gm_dtedge_params(C.Structure):
    _fields_ = [("sigmas", C.c_double * GM_MAX_SCALES), ("p_hi", C.c_double),
                ("n_sigmas", C.c_int32), ("morph_open", C.c_int32), ("layout", C.c_int32),
                ("flags", C.c_int32)]


def _load() -> C.CDLL:
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  This package has no CPU fallback.")
    return C.CDLL(LIB_PATH)


lib = None
lib = _load()

_vp, _i32, _i64, _f32, _f64, _sz = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_double, C.c_size_t
_p = C.POINTER

# name -> (restype, argtypes); must list every symbol of include/geomap_b200.h
PROTOTYPES = {
    "gm_version": (C.c_int, []),
    "gm_status_string": (C.c_char_p, [C.c_int]),
    "gm_device_check": (C.c_int, []),
    "gm_launch_count": (_i64, []),
    "gm_tile_plan_count": (_i64, [_i32, _i32, _i32, _i32, _p(_i32), _p(_i32), _p(_i64)]),
    "gm_tile_plan_fill": (_i64, [_i32, _i32, _i32, _i32, _i32, _i32, _p(gm_tile), _i64, _p(_i64)]),
    "gm_tile_gather_u8": (C.c_int, [_vp, _i32, _i32, _vp, _i32, _i32, _vp, _vp]),
    "gm_dtedge_workspace_bytes": (_sz, [_i64, _i32]),
    "gm_dtedge_select_stats": (C.c_int, [_p(C.c_uint32), _i32]),
    "gm_band_pack": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _vp, _i64, _vp, _vp]),
    "gm_band_unpack": (C.c_int, [_vp, _i64, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "gm_band_mask_keep": (C.c_int, [_vp, _vp, _i64, _vp]),
    "gm_band_extract_workspace_bytes": (_sz, [_i64]),
    "gm_band_extract": (C.c_int, [_vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "gm_band_merge_workspace_bytes": (_sz, [_i64, _i32, _i64, _i64]),
    "gm_band_merge_local": (C.c_int, [_vp, _vp, _vp, _i64, _vp, _i32, _f64, _i64, _p(_f32), _i32, _f32, _i32, _i64, _vp, _vp, _sz, _vp]),
    "gm_band_merge_finish": (C.c_int, [_vp, _i32, _i32, _i64, _i32, _i32, _p(_f32), _i32, _f32, _vp, _vp, _vp, _vp, _i64, _i32, _f64,
                                       _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "gm_dtedge_build_u8": (C.c_int, [_vp, _i32, _i32, _vp, _i32, _i32, _i64, _p(gm_dtedge_params), _vp, _vp, _sz, _vp]),
    "gm_dtedge_build_range_u8": (C.c_int, [_vp, _i32, _i32, _vp, _i32, _i32, _i64, _i32, _i32, _p(gm_dtedge_params), _vp, _vp,
                                           _sz, _vp]),
    "gm_dtedge_build_timed": (C.c_int, [_vp, _i32, _i32, _vp, _i32, _i32, _i64, _p(gm_dtedge_params), _vp, _vp, _sz, _vp,
                                        _p(_f32)]),
    "gm_dtedge_workspace_views": (C.c_int, [_vp, _i64, _i32, _p(_vp), _p(_vp), _p(_vp)]),
    "gm_train_tile_grid": (_i64, [_i32, _i32, _i32, _i32, _p(_i32), _p(_i32), _p(_i32)]),
    "gm_train_label_tiles": (C.c_int, [_vp, _i64, _i32, _i32, _i32, _i32, _f64, _vp, _vp, _vp, _vp]),
    "gm_letterbox_shape": (C.c_int, [_i32, _i32, _i32, _i32, _i32, _p(_i32), _p(_i32), _p(_i32), _p(_i32), _p(_i32), _p(_i32)]),
    "gm_letterbox_tiles": (C.c_int, [_vp, _i32, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp]),
    "gm_eval_iou_segments": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i64, _vp, _vp]),
    "gm_eval_match_greedy": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i64, _i64, _vp, _i32, _vp, _vp, _vp]),
    "gm_eval_center_hit": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp]),
    "gm_rotated_iou_pairs": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _vp, _vp]),
    "gm_rotated_iou_pairs_f64": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _vp, _vp]),
    "gm_rotated_iou_matrix": (C.c_int, [_vp, _i32, _vp, _i32, _vp, _vp]),
    "gm_rotated_iou_matrix_sum": (C.c_int, [_vp, _i32, _vp, _i32, _vp, _vp]),
    "gm_decode_workspace_bytes": (_sz, [_i32, _i32]),
    "gm_decode_tiles": (C.c_int, [_vp, _i32, _i32, _i32, _vp, _i32, _i32, _f32, _f32, _i32, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "gm_tile_postprocess_workspace_bytes": (_sz, [_i64, _i64]),
    "gm_tile_postprocess": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _vp, _i32, _i32, _i32, _i32, _f64, _i64,
                                      _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "gm_tile_postprocess_bounded": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _vp, _i32, _i32, _i32, _i32, _f64, _i64, _i32,
                                              _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "gm_nms_workspace_bytes": (_sz, [_i64, _i64]),
    "gm_threshold_adjacent_stats": (C.c_int, [_p(C.c_uint64), _i32]),
    "gm_nms_global": (C.c_int, [_vp, _vp, _vp, _i64, _i32, _f64, _i64, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "gm_fuse_workspace_bytes": (_sz, [_i64, _i64]),
    "gm_fuse_scales": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, _i32, _f64, _f64, _f64, _i64, _vp, _vp, _vp, _sz, _vp]),
    "gm_build_multich_host": (C.c_int, [_vp, _i32, _i32, _i32, _p(gm_dtedge_params), _vp]),
    "gm_polygon_iou_host": (C.c_int, [_p(_f64), _p(_f64), _p(_f64)]),
    "gm_ffma_peak": (C.c_int, [_i32, _p(_f64), _vp]),
}

for _name, (_res, _args) in PROTOTYPES.items():
    _fn = getattr(lib, _name)          # AttributeError here == symbol missing from the build
    _fn.restype = _res
    _fn.argtypes = _args


def check(status: int, where: str) -> None:
    if status != GM_OK:
        raise GmError(int(status), where)


GM_DTEDGE_GENERIC_GRAD = 1
GM_DTEDGE_OTSU = 2


def bin_method_flags(bin_method: str) -> int:
    """DT_BIN_METHOD -> flags, with the reference's own dispatch (Detect_OBB.py:109-114): "otsu" selects the Otsu
    branch, every other value falls through to the percentile branch."""
    return GM_DTEDGE_OTSU if bin_method == "otsu" else 0


def make_params(sigmas=(0, 0.6, 1.2, 2.4), p_hi=90.0, morph_open=1, layout=0, flags=0) -> gm_dtedge_params:
    sigmas = tuple(float(s) for s in sigmas)
    if not 1 <= len(sigmas) <= GM_MAX_SCALES:
        raise ValueError(f"1..{GM_MAX_SCALES} sigmas supported, got {len(sigmas)}")
    p = gm_dtedge_params()
    for i, s in enumerate(sigmas):
        p.sigmas[i] = s
    p.n_sigmas = len(sigmas)
    p.p_hi = float(p_hi)
    p.morph_open = int(morph_open)
    p.layout = int(layout)
    p.flags = int(flags)
    return p
