"""ctypes front end of oracle/geom_c.c (the C restatement of the geometry oracle).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle_geom.so")


def _load():
    if not os.path.isfile(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, "geom_c.c")):
        subprocess.run(["make", "-s", "-C", _HERE], check=True)
    lib = C.CDLL(_SO)
    dp, ip, fp = C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_float)
    lib.orc_quad_iou.restype = C.c_double
    lib.orc_quad_iou.argtypes = [dp, dp]
    lib.orc_nms.restype = C.c_int
    lib.orc_nms.argtypes = [dp, ip, fp, C.c_int, C.c_double, ip, ip]
    lib.orc_fuse.restype = C.c_int
    lib.orc_fuse.argtypes = [dp, ip, fp, ip, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, ip]
    lib.orc_nms_grid.restype = C.c_int
    lib.orc_nms_grid.argtypes = lib.orc_nms.argtypes
    lib.orc_fuse_grid.restype = C.c_int
    lib.orc_fuse_grid.argtypes = lib.orc_fuse.argtypes
    lib.orc_iou_pairs.restype = None
    lib.orc_iou_pairs.argtypes = [dp, dp, C.c_longlong, dp]
    return lib


_lib = _load()
_dp, _ip, _fp = C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_float)


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def quad_iou(a, b) -> float:
    a, b = _d(a), _d(b)
    return float(_lib.orc_quad_iou(a.ctypes.data_as(_dp), b.ctypes.data_as(_dp)))


def iou_pairs(a, b) -> np.ndarray:
    a, b = _d(a), _d(b)
    out = np.empty(a.shape[0], dtype=np.float64)
    _lib.orc_iou_pairs(a.ctypes.data_as(_dp), b.ctypes.data_as(_dp), a.shape[0], out.ctypes.data_as(_dp))
    return out


def nms(boxes, cls, conf, thr, grid: bool = False):
    """(order, kept): stable conf-desc permutation and kept input indices in output order.
    grid=True: the same sequential algorithm with a uniform-grid candidate lookup (10^6-box cases)."""
    boxes = _d(boxes)
    cls = np.ascontiguousarray(cls, dtype=np.int32)
    conf = np.ascontiguousarray(conf, dtype=np.float32)
    n = boxes.shape[0]
    order = np.empty(n, dtype=np.int32)
    kept = np.empty(n, dtype=np.int32)
    k = (_lib.orc_nms_grid if grid else _lib.orc_nms)(boxes.ctypes.data_as(_dp), cls.ctypes.data_as(_ip), conf.ctypes.data_as(_fp), n, float(thr),
                     order.ctypes.data_as(_ip), kept.ctypes.data_as(_ip))
    return order, kept[:k]


def fuse(boxes, cls, conf, scale, n_scales, iou_partner=0.40, conf_low=0.25, conf_high=0.70, grid: bool = False):
    boxes = _d(boxes)
    cls = np.ascontiguousarray(cls, dtype=np.int32)
    conf = np.ascontiguousarray(conf, dtype=np.float32)
    scale = np.ascontiguousarray(scale, dtype=np.int32)
    n = boxes.shape[0]
    kept = np.empty(max(n, 1), dtype=np.int32)
    k = (_lib.orc_fuse_grid if grid else _lib.orc_fuse)(boxes.ctypes.data_as(_dp), cls.ctypes.data_as(_ip), conf.ctypes.data_as(_fp),
                      scale.ctypes.data_as(_ip), n, int(n_scales), iou_partner, conf_low, conf_high,
                      kept.ctypes.data_as(_ip))
    return kept[:k]
