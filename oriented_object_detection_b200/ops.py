"""Device-level Python API: torch tensors in, torch tensors out, every call goes through the C ABI.

torch is plumbing here (device memory, streams); all arithmetic is in ``libgeomap_b200.so``.
Each function names the reference function it stands in for (file:line of
``/root/reference``).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib as L


def _require_cuda() -> None:
    if not torch.cuda.is_available():
        raise RuntimeError("geomap_b200 needs an sm_100 CUDA device; there is no CPU fallback")


def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


_workspaces = {}
_scope = [""]


class workspace_scope:
    """Calls inside the ``with`` block take their scratch buffers from a private namespace.  A CUDA graph
    bakes the addresses of the buffers its kernels used into its nodes; a capture therefore runs inside its
    own scope and keeps the tensors (:func:`workspaces_of`), so later eager calls that grow the shared
    buffers cannot free memory a graph still replays into."""

    def __init__(self, name: str):
        self.name = name

    def __enter__(self):
        _scope.append(self.name)
        return self

    def __exit__(self, *exc):
        _scope.pop()
        return False


def workspaces_of(scope: str):
    return [t for (dev, name), t in _workspaces.items() if name.startswith(scope + "/")]


def _workspace(name: str, nbytes: int, device) -> torch.Tensor:
    """Grow-only scratch buffer per (device, purpose); kernels are stream-ordered so reuse is safe."""
    if _scope[-1]:
        name = _scope[-1] + "/" + name
    key = (str(device), name)
    t = _workspaces.get(key)
    if t is None or t.numel() < nbytes:
        t = torch.empty(int(nbytes * 1.25) + 1024, dtype=torch.uint8, device=device)
        _workspaces[key] = t
    return t


def release_workspaces() -> None:
    _workspaces.clear()


# ----------------------------------------------------------------------------- a1: tile plan

@dataclass
class TilePlan:
    """Overlapped tile plan of one map (or of one rank's row band).  Detect_OBB.py:210-223."""
    H: int
    W: int
    tile_size: int
    overlap: int
    rows: int              # tile rows in this plan
    cols: int
    row_begin: int
    tiles: np.ndarray      # structured: y0, x0, h, w, px_off
    total_px: int
    dev: Optional[torch.Tensor] = None   # the same records on the device (uint8 view)

    @property
    def n(self) -> int:
        return int(self.tiles.shape[0])

    @property
    def max_tile(self) -> int:
        return int(max(self.tiles["h"].max(), self.tiles["w"].max())) if self.n else 0

    def to(self, device) -> "TilePlan":
        if self.dev is None or self.dev.device != torch.device(device):
            self.dev = torch.from_numpy(self.tiles.view(np.uint8).copy()).to(device)
        return self


TILE_DTYPE = np.dtype([("y0", "<i4"), ("x0", "<i4"), ("h", "<i4"), ("w", "<i4"), ("px_off", "<i8")])
assert TILE_DTYPE.itemsize == C.sizeof(L.gm_tile)


def make_plan(H: int, W: int, tile_size: int, overlap: int, row_begin: int = 0, row_end: int = -1,
              device=None) -> TilePlan:
    rows, cols, total = C.c_int32(), C.c_int32(), C.c_int64()
    n_all = L.lib.gm_tile_plan_count(H, W, tile_size, overlap, C.byref(rows), C.byref(cols), C.byref(total))
    if n_all < 0:
        raise L.GmError(int(n_all), "gm_tile_plan_count")
    if row_end < 0 or row_end > rows.value:
        row_end = rows.value
    count = (row_end - row_begin) * cols.value
    arr = np.zeros(max(count, 0), dtype=TILE_DTYPE)
    tot = C.c_int64()
    got = L.lib.gm_tile_plan_fill(H, W, tile_size, overlap, row_begin, row_end,
                                  arr.ctypes.data_as(C.POINTER(L.gm_tile)), arr.shape[0], C.byref(tot))
    if got < 0:
        raise L.GmError(int(got), "gm_tile_plan_fill")
    plan = TilePlan(H, W, tile_size, overlap, row_end - row_begin, cols.value, row_begin, arr, int(tot.value))
    if device is not None:
        plan.to(device)
    return plan


def plan_from_arrays(H: int, W: int, y0, x0, h, w, device=None, tile_size: int = 0, overlap: int = 0) -> TilePlan:
    """Tile batch from coordinate arrays (e.g. a rank's tile range of several maps stacked in one buffer)."""
    n = len(y0)
    arr = np.zeros(n, dtype=TILE_DTYPE)
    arr["y0"], arr["x0"], arr["h"], arr["w"] = y0, x0, h, w
    px = np.asarray(h, dtype=np.int64) * np.asarray(w, dtype=np.int64)
    arr["px_off"] = np.concatenate([[0], np.cumsum(px)[:-1]]) if n else np.zeros(0, np.int64)
    plan = TilePlan(H, W, tile_size, overlap, 0, 0, 0, arr, int(px.sum()))
    if device is not None:
        plan.to(device)
    return plan


def plan_from_tiles(H: int, W: int, tiles: Sequence[Tuple[int, int, int, int]], device=None) -> TilePlan:
    """Arbitrary list of (y0, x0, h, w) crops treated as a tile batch."""
    arr = np.zeros(len(tiles), dtype=TILE_DTYPE)
    off = 0
    for i, (y0, x0, h, w) in enumerate(tiles):
        arr[i] = (y0, x0, h, w, off)
        off += h * w
    plan = TilePlan(H, W, 0, 0, 0, 0, 0, arr, off)
    if device is not None:
        plan.to(device)
    return plan


# ----------------------------------------------------------------------------- a2: 3-ch gather

def tile_gather(map_bgr: torch.Tensor, plan: TilePlan, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Packed BGR tiles of ``map_bgr`` (uint8 [H,W,3], CUDA).  build_multich 3-ch, Detect_OBB.py:92-93."""
    _require_cuda()
    assert map_bgr.dtype == torch.uint8 and map_bgr.is_cuda and map_bgr.is_contiguous()
    H, W = int(map_bgr.shape[0]), int(map_bgr.shape[1])
    plan.to(map_bgr.device)
    if out is None:
        out = torch.empty(3 * plan.total_px, dtype=torch.uint8, device=map_bgr.device)
    L.check(L.lib.gm_tile_gather_u8(_ptr(map_bgr), H, W, _ptr(plan.dev), plan.n, plan.max_tile, _ptr(out), _stream()),
            "gm_tile_gather_u8")
    return out


# ----------------------------------------------------------------------------- a3: DT-Edge

def dtedge_build(map_bgr: torch.Tensor, plan: TilePlan, params: Optional[L.gm_dtedge_params] = None,
                 out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Packed [R,G,B,DT-Edge] tiles.  build_multich 4-ch, Detect_OBB.py:95-133 / Train_OBB.py:615-664."""
    _require_cuda()
    assert map_bgr.dtype == torch.uint8 and map_bgr.is_cuda and map_bgr.is_contiguous()
    H, W = int(map_bgr.shape[0]), int(map_bgr.shape[1])
    plan.to(map_bgr.device)
    if params is None:
        params = L.make_params()
    if out is None:
        out = torch.empty(4 * plan.total_px, dtype=torch.uint8, device=map_bgr.device)
    need = L.lib.gm_dtedge_workspace_bytes(plan.total_px, plan.n)
    ws = _workspace("dtedge", need, map_bgr.device)
    L.check(L.lib.gm_dtedge_build_u8(_ptr(map_bgr), H, W, _ptr(plan.dev), plan.n, plan.max_tile, plan.total_px,
                                     C.byref(params), _ptr(out), _ptr(ws), ws.numel(), _stream()),
            "gm_dtedge_build_u8")
    return out


# ----------------------------------------------------------------------------- host maps: chunked upload overlapped with the build

_side_streams = {}


def _side_stream(device, name: str) -> torch.cuda.Stream:
    key = (str(torch.device(device)), name)
    if key not in _side_streams:
        _side_streams[key] = torch.cuda.Stream(device=device)
    return _side_streams[key]


def build_tiles_from_host(map_host: torch.Tensor, plan: TilePlan, channels: int = 4,
                          params: Optional[L.gm_dtedge_params] = None, out: Optional[torch.Tensor] = None,
                          map_dev: Optional[torch.Tensor] = None, n_chunks: int = 8, n_build_streams: int = 3):
    """Tile batch of a map that lives in (pinned) HOST memory: ``map_host`` uint8 [H,W,3] BGR.

    The map is uploaded in chunks of tile rows on a copy stream while the current stream builds the
    tiles of the chunks that have already arrived (tiles are independent, Detect_OBB.py:210-225, and a
    tile row needs only its own pixel rows), so the PCIe copy hides the build instead of preceding it.
    ``plan.tiles['y0']`` are rows of ``map_host``.  Returns (packed tiles, device copy of the map).
    """
    _require_cuda()
    assert channels in (3, 4), f"Unsupported out_channels={channels}"
    assert map_host.dtype == torch.uint8 and not map_host.is_cuda and map_host.is_contiguous()
    H, W = int(map_host.shape[0]), int(map_host.shape[1])
    dev = plan.dev.device if plan.dev is not None else torch.device("cuda", torch.cuda.current_device())
    plan.to(dev)
    if map_dev is None:
        map_dev = torch.empty((H, W, 3), dtype=torch.uint8, device=dev)
    if out is None:
        out = torch.empty(channels * plan.total_px, dtype=torch.uint8, device=dev)
    if channels == 4:
        if params is None:
            params = L.make_params()
        need = L.lib.gm_dtedge_workspace_bytes(plan.total_px, plan.n)
        ws = _workspace("dtedge", need, dev)
    # chunks = contiguous tile ranges of (nearly) equal size; the tiles of a plan are ordered by non-decreasing y0 (tile
    # rows of one map, or the bands of several maps stacked in one buffer), so "the pixel rows a range needs" is a prefix
    n_tiles = plan.n
    n_chunks = max(1, min(n_chunks, n_tiles))
    y_bottom = plan.tiles["y0"].astype(np.int64) + plan.tiles["h"].astype(np.int64)
    assert n_tiles == 0 or bool(np.all(np.diff(plan.tiles["y0"].astype(np.int64)) >= 0)), "tiles must be ordered by y0"
    main = torch.cuda.current_stream(dev)
    side = _side_stream(dev, "copy")
    builders = [_side_stream(dev, f"build{i}") for i in range(max(1, n_build_streams))]
    side.wait_stream(main)                     # earlier readers of map_dev are done before it is overwritten
    for b in builders:
        b.wait_stream(main)
    y_done = 0
    tile_bytes = C.sizeof(L.gm_tile)
    for k in range(n_chunks):
        t0, t1 = (n_tiles * k) // n_chunks, (n_tiles * (k + 1)) // n_chunks
        if t1 <= t0:
            continue
        y_end = int(y_bottom[t0:t1].max()) if k + 1 < n_chunks else H
        if y_end > y_done:
            with torch.cuda.stream(side):
                map_dev[y_done:y_end].copy_(map_host[y_done:y_end], non_blocking=True)
            y_done = y_end
        bs = builders[k % len(builders)]
        bs.wait_stream(side)
        if channels == 4:
            L.check(L.lib.gm_dtedge_build_range_u8(_ptr(map_dev), H, W, _ptr(plan.dev), plan.n, plan.max_tile,
                                                   plan.total_px, t0, t1 - t0, C.byref(params), _ptr(out), _ptr(ws),
                                                   ws.numel(), C.c_void_p(bs.cuda_stream)),
                    "gm_dtedge_build_range_u8")
        else:
            L.check(L.lib.gm_tile_gather_u8(_ptr(map_dev), H, W, C.c_void_p(plan.dev.data_ptr() + t0 * tile_bytes),
                                            t1 - t0, plan.max_tile, _ptr(out), C.c_void_p(bs.cuda_stream)),
                    "gm_tile_gather_u8")
    for b in builders:
        main.wait_stream(b)
    main.wait_stream(side)
    return out, map_dev


DTEDGE_STAGES = ("grad", "select_grad", "edge_open", "chamfer", "select_dist", "tail")


def dtedge_build_timed(map_bgr: torch.Tensor, plan: TilePlan, params: Optional[L.gm_dtedge_params] = None,
                       out: Optional[torch.Tensor] = None):
    """:func:`dtedge_build` with CUDA events between its kernels; returns (out, {stage: ms})."""
    _require_cuda()
    H, W = int(map_bgr.shape[0]), int(map_bgr.shape[1])
    plan.to(map_bgr.device)
    if params is None:
        params = L.make_params()
    if out is None:
        out = torch.empty(4 * plan.total_px, dtype=torch.uint8, device=map_bgr.device)
    need = L.lib.gm_dtedge_workspace_bytes(plan.total_px, plan.n)
    ws = _workspace("dtedge", need, map_bgr.device)
    ms = (C.c_float * len(DTEDGE_STAGES))()
    L.check(L.lib.gm_dtedge_build_timed(_ptr(map_bgr), H, W, _ptr(plan.dev), plan.n, plan.max_tile, plan.total_px,
                                        C.byref(params), _ptr(out), _ptr(ws), ws.numel(), _stream(), ms),
            "gm_dtedge_build_timed")
    return out, {k: float(ms[i]) for i, k in enumerate(DTEDGE_STAGES)}


def dtedge_debug_views(plan: TilePlan, device) -> dict:
    """Intermediates of the last :func:`dtedge_build` on ``device`` (parity tests): per tile
    S (uint32), chamfer field (uint32, 16.16) and the opened-edge zero mask (bool)."""
    ws = _workspaces[(str(torch.device(device)), "dtedge")]
    S, Z, T = C.c_void_p(), C.c_void_p(), C.c_void_p()
    L.check(L.lib.gm_dtedge_workspace_views(_ptr(ws), plan.total_px, plan.n, C.byref(S), C.byref(Z), C.byref(T)),
            "gm_dtedge_workspace_views")
    base = ws.data_ptr()
    torch.cuda.synchronize()
    raw = ws.cpu().numpy()

    def view(ptr, count, dtype):
        o = ptr.value - base
        return raw[o:o + count * np.dtype(dtype).itemsize].view(dtype)

    S_all = view(S, plan.total_px, np.uint32)
    T_all = view(T, plan.total_px, np.uint32)
    zwords = plan.total_px // 32 + plan.n * (L.GM_MAX_TILE + 1) + 2
    Z_all = view(Z, zwords, np.uint32)
    out = {"S": [], "t": [], "zero": []}
    for ti, t in enumerate(plan.tiles):
        h, w, off = int(t["h"]), int(t["w"]), int(t["px_off"])
        out["S"].append(S_all[off:off + h * w].reshape(h, w))
        out["t"].append(T_all[off:off + h * w].reshape(h, w))
        wpr = (w + 31) // 32
        zo = (off >> 5) + ti * (L.GM_MAX_TILE + 1)
        words = Z_all[zo:zo + h * wpr].reshape(h, wpr)
        bits = np.unpackbits(words.view(np.uint8), axis=1, bitorder="little")[:, :w]
        out["zero"].append(bits.astype(bool))
    return out


# ----------------------------------------------------------------------------- a10: rotated IoU

def _boxes64(b: torch.Tensor) -> torch.Tensor:
    assert b.is_cuda and b.shape[-1] == 8
    return b.to(torch.float64).contiguous()


def rotated_iou_pairs(boxes_a: torch.Tensor, boxes_b: torch.Tensor, idx_a: Optional[torch.Tensor] = None,
                      idx_b: Optional[torch.Tensor] = None) -> torch.Tensor:
    """IoU of (a[idx_a[p]], b[idx_b[p]]) (row p of each when no index).  compute_polygon_iou, Detect_OBB.py:144-154."""
    _require_cuda()
    a, b = _boxes64(boxes_a), _boxes64(boxes_b)
    if idx_a is not None:
        idx_a = idx_a.to(torch.int32).contiguous()
        idx_b = idx_b.to(torch.int32).contiguous()
        n = idx_a.numel()
    else:
        n = a.shape[0]
        assert b.shape[0] == n
    out = torch.empty(n, dtype=torch.float32, device=a.device)
    L.check(L.lib.gm_rotated_iou_pairs(_ptr(a), _ptr(b), _ptr(idx_a), _ptr(idx_b), n, _ptr(out), _stream()),
            "gm_rotated_iou_pairs")
    return out


def rotated_iou_pairs_f64(boxes_a: torch.Tensor, boxes_b: torch.Tensor, idx_a: Optional[torch.Tensor] = None,
                          idx_b: Optional[torch.Tensor] = None) -> torch.Tensor:
    """The same pair list in float64: the reference's own arithmetic (compute_polygon_iou on Python floats,
    Detect_OBB.py:144-154), concave simple quads included."""
    _require_cuda()
    a, b = _boxes64(boxes_a), _boxes64(boxes_b)
    if idx_a is not None:
        idx_a = idx_a.to(torch.int32).contiguous()
        idx_b = idx_b.to(torch.int32).contiguous()
        n = idx_a.numel()
    else:
        n = a.shape[0]
        assert b.shape[0] == n
    out = torch.empty(n, dtype=torch.float64, device=a.device)
    L.check(L.lib.gm_rotated_iou_pairs_f64(_ptr(a), _ptr(b), _ptr(idx_a), _ptr(idx_b), n, _ptr(out), _stream()),
            "gm_rotated_iou_pairs_f64")
    return out


def rotated_iou_matrix(boxes_a: torch.Tensor, boxes_b: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _require_cuda()
    a, b = _boxes64(boxes_a), _boxes64(boxes_b)
    if out is None:
        out = torch.empty((a.shape[0], b.shape[0]), dtype=torch.float32, device=a.device)
    L.check(L.lib.gm_rotated_iou_matrix(_ptr(a), a.shape[0], _ptr(b), b.shape[0], _ptr(out), _stream()),
            "gm_rotated_iou_matrix")
    return out


def rotated_iou_matrix_sum(boxes_a: torch.Tensor, boxes_b: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Column sums (sum over boxes_a for every box of boxes_b) of the dense IoU matrix without storing it
    (FP32-throughput measurement)."""
    _require_cuda()
    a, b = _boxes64(boxes_a), _boxes64(boxes_b)
    if out is None:
        out = torch.empty(b.shape[0], dtype=torch.float64, device=a.device)
    L.check(L.lib.gm_rotated_iou_matrix_sum(_ptr(a), a.shape[0], _ptr(b), b.shape[0], _ptr(out), _stream()),
            "gm_rotated_iou_matrix_sum")
    return out


def ffma_peak(iters: int = 4096) -> float:
    _require_cuda()
    v = C.c_double()
    L.check(L.lib.gm_ffma_peak(iters, C.byref(v), _stream()), "gm_ffma_peak")
    return float(v.value)


# ----------------------------------------------------------------------------- a11: NMS

def threshold_adjacent_stats(reset: bool = False) -> dict:
    """Pairs whose ``IoU >= thr`` decision needed float64 since the last reset (see ``gm_threshold_adjacent_stats``):
    ``{"float64_decided": a, "within_1e-5": b}``.  Synchronises the device."""
    _require_cuda()
    c = (C.c_uint64 * 2)()
    L.check(L.lib.gm_threshold_adjacent_stats(c, 1 if reset else 0), "gm_threshold_adjacent_stats")
    return {"float64_decided": int(c[0]), "within_1e-5": int(c[1])}


def nms_global(boxes: torch.Tensor, cls: torch.Tensor, conf: torch.Tensor, iou_thr: float, max_class: Optional[int] = None,
               edge_capacity: int = 0, sync: bool = True):
    """Exact class-wise greedy rotated NMS.  merge_detections, Detect_OBB.py:176-200.

    Returns (order, keep, kept_idx): the stable confidence-descending permutation, the keep
    flag per input box, and the kept input indices in output order.  With ``sync=False`` the
    kept list is returned padded together with the device count (no host read).
    """
    _require_cuda()
    b = _boxes64(boxes)
    n = b.shape[0]
    dev = b.device
    cls = cls.to(torch.int32).contiguous()
    conf = conf.to(torch.float32).contiguous()
    if max_class is None:
        max_class = int(cls.max().item()) if n else 0
    order = torch.empty(n, dtype=torch.int32, device=dev)
    keep = torch.empty(n, dtype=torch.uint8, device=dev)
    kept = torch.empty(n, dtype=torch.int32, device=dev)
    cnt = torch.zeros(1, dtype=torch.int64, device=dev)
    cap = edge_capacity
    while True:
        need = L.lib.gm_nms_workspace_bytes(n, cap)
        ws = _workspace("merge", need, dev)
        L.check(L.lib.gm_nms_global(_ptr(b), _ptr(cls), _ptr(conf), n, int(max_class), float(iou_thr), cap,
                                    _ptr(order), _ptr(keep), _ptr(kept), _ptr(cnt), _ptr(ws), ws.numel(), _stream()),
                "gm_nms_global")
        if not sync:
            return order, keep, kept, cnt
        k = int(cnt.item())
        if k >= 0:
            return order, keep, kept[:k]
        cap = -k + 1024          # pairs found exceeded the edge buffer: rerun with room for all


def tile_postprocess(boxes_local: torch.Tensor, cls: torch.Tensor, conf: torch.Tensor, tile_id: torch.Tensor,
                     plan: TilePlan, margin_px: int, angle_class: int, iou_merge: float, max_class: int,
                     edge_capacity: int = 0, sync: bool = True, max_per_tile: int = 0):
    """Remap + border filter + strike angle + per-tile NMS.  detect_symbols body, Detect_OBB.py:228-264.

    ``max_per_tile``: the caller's bound on the detections of one tile (a detector's max_det, Ultralytics: 300).  With
    0 < max_per_tile <= 320 the per-tile NMS runs as one CTA per tile (score-sorted bit-mask sweep) instead of the general
    engine; same results, and a tile above the bound is still exact, only slow.  0 = no promise: the general engine.

    Returns dict(boxes float64 [m,8], cls, conf, angle float64, src) in the reference's list order.
    ``sync=False``: no host read - the arrays keep their full input length n, the first ``count`` rows
    (device int64[1], key "count") are the survivors; a negative count means the pair buffer overflowed.
    """
    _require_cuda()
    dev = boxes_local.device
    n = boxes_local.shape[0]
    bl = boxes_local.to(torch.float32).contiguous()
    cls = cls.to(torch.int32).contiguous()
    conf = conf.to(torch.float32).contiguous()
    tile_id = tile_id.to(torch.int32).contiguous()
    plan.to(dev)
    ob = torch.empty((n, 8), dtype=torch.float64, device=dev)
    oc = torch.empty(n, dtype=torch.int32, device=dev)
    of = torch.empty(n, dtype=torch.float32, device=dev)
    oa = torch.empty(n, dtype=torch.float64, device=dev)
    osrc = torch.empty(n, dtype=torch.int32, device=dev)
    cnt = torch.zeros(1, dtype=torch.int64, device=dev)
    cap = edge_capacity
    while True:
        need = L.lib.gm_tile_postprocess_workspace_bytes(n, cap)
        ws = _workspace("merge", need, dev)
        L.check(L.lib.gm_tile_postprocess_bounded(_ptr(bl), _ptr(cls), _ptr(conf), _ptr(tile_id), n, _ptr(plan.dev), plan.n,
                                                  int(max_class), int(margin_px), int(angle_class), float(iou_merge), cap,
                                                  int(max_per_tile), _ptr(ob), _ptr(oc), _ptr(of), _ptr(oa), _ptr(osrc),
                                                  _ptr(cnt), _ptr(ws), ws.numel(), _stream()), "gm_tile_postprocess")
        if not sync:
            return {"boxes": ob, "cls": oc, "conf": of, "angle": oa, "src": osrc, "count": cnt}
        k = int(cnt.item())
        if k >= 0:
            break
        cap = -k + 1024
    return {"boxes": ob[:k], "cls": oc[:k], "conf": of[:k], "angle": oa[:k], "src": osrc[:k]}


# ----------------------------------------------------------------------------- (e) cross-band exchange records

BAND_RECORD_BYTES = 80


def band_pack(rec: dict, count: torch.Tensor, capacity: int) -> torch.Tensor:
    """Survivors of one band (padded arrays + device count, ``tile_postprocess(sync=False)``) as ``capacity`` fixed-size
    records, rows >= count blank.  uint8 [capacity, 80]."""
    _require_cuda()
    dev = rec["conf"].device
    boxes = _boxes64(rec["boxes"])
    cls = rec["cls"].to(torch.int32).contiguous()
    conf = rec["conf"].to(torch.float32).contiguous()
    angle = rec["angle"].to(torch.float64).contiguous() if "angle" in rec else None
    count = count.to(device=dev, dtype=torch.int64).contiguous()
    out = torch.empty((capacity, BAND_RECORD_BYTES), dtype=torch.uint8, device=dev)
    L.check(L.lib.gm_band_pack(_ptr(boxes), _ptr(cls), _ptr(conf), _ptr(angle), int(conf.shape[0]), _ptr(count), int(capacity),
                               _ptr(out), _stream()), "gm_band_pack")
    return out


def band_unpack(records: torch.Tensor, world: int, rank: int, with_angle: bool = True) -> dict:
    """Gathered records -> SoA arrays; ``cls_owned`` = class where class % world == rank, else -1."""
    _require_cuda()
    assert records.is_cuda and records.dtype == torch.uint8 and records.is_contiguous() and records.shape[1] == BAND_RECORD_BYTES
    dev = records.device
    total = int(records.shape[0])
    out = {"boxes": torch.empty((total, 8), dtype=torch.float64, device=dev),
           "cls": torch.empty(total, dtype=torch.int32, device=dev),
           "cls_owned": torch.empty(total, dtype=torch.int32, device=dev),
           "conf": torch.empty(total, dtype=torch.float32, device=dev),
           "n_valid": torch.empty(1, dtype=torch.int64, device=dev)}
    if with_angle:
        out["angle"] = torch.empty(total, dtype=torch.float64, device=dev)
    L.check(L.lib.gm_band_unpack(_ptr(records), total, int(world), int(rank), _ptr(out["boxes"]), _ptr(out["cls"]),
                                 _ptr(out["cls_owned"]), _ptr(out["conf"]), _ptr(out.get("angle")), _ptr(out["n_valid"]),
                                 _stream()), "gm_band_unpack")
    return out


def band_mask_keep(keep: torch.Tensor, cls_owned: torch.Tensor) -> torch.Tensor:
    """In place: keep[i] = 0 for rows this rank does not own."""
    _require_cuda()
    assert keep.dtype == torch.uint8 and keep.is_contiguous() and cls_owned.dtype == torch.int32
    L.check(L.lib.gm_band_mask_keep(_ptr(keep), _ptr(cls_owned), int(keep.numel()), _stream()), "gm_band_mask_keep")
    return keep


def band_extract(order: torch.Tensor, keep: torch.Tensor, fields: dict) -> dict:
    """Kept rows in the stable confidence order, compacted: padded arrays + ``n_out`` (device int64[1])."""
    _require_cuda()
    dev = keep.device
    total = int(keep.numel())
    order = order.to(torch.int32).contiguous()
    out = {"boxes": torch.empty((total, 8), dtype=torch.float64, device=dev),
           "cls": torch.empty(total, dtype=torch.int32, device=dev),
           "conf": torch.empty(total, dtype=torch.float32, device=dev),
           "index": torch.empty(total, dtype=torch.int64, device=dev),
           "n_out": torch.empty(1, dtype=torch.int64, device=dev)}
    if "angle" in fields:
        out["angle"] = torch.empty(total, dtype=torch.float64, device=dev)
    need = L.lib.gm_band_extract_workspace_bytes(total)
    ws = _workspace("band", need, dev)
    L.check(L.lib.gm_band_extract(_ptr(order), _ptr(keep), total, _ptr(fields["boxes"]), _ptr(fields["cls"]), _ptr(fields["conf"]),
                                  _ptr(fields.get("angle")), _ptr(out["boxes"]), _ptr(out["cls"]), _ptr(out["conf"]),
                                  _ptr(out.get("angle")), _ptr(out["index"]), _ptr(out["n_out"]), _ptr(ws), ws.numel(), _stream()),
            "gm_band_extract")
    return out


# ----------------------------------------------------------------------------- (e) seam-band exchange

SEAM_EDGE_OVERFLOW, SEAM_CAPACITY_OVERFLOW, SEAM_EXTENT_EXCEEDED, SEAM_INPUT_OVERFLOW, SEAM_CHAIN_ESCAPES = 1, 4, 8, 16, 32


def seam_status_text(status: int) -> str:
    names = {SEAM_EDGE_OVERFLOW: "more overlapping pairs than edge_capacity", SEAM_CAPACITY_OVERFLOW: "a rank deferred more boxes than "
             "seam_capacity", SEAM_EXTENT_EXCEEDED: "a box is larger than extent_bound", SEAM_INPUT_OVERFLOW: "the per-tile stage "
             "overflowed its pair buffer", SEAM_CHAIN_ESCAPES: "a chain of overlaps leaves the restricted rank range"}
    return "; ".join(v for k, v in names.items() if status & k) or "ok"


def band_merge_local(rec: dict, count: torch.Tensor, max_class: int, iou_thr: float, rects, extent_bound: float, world: int,
                     seam_capacity: int, edge_capacity: int = 0):
    """Global NMS of this rank's OWN band with the boxes that depend on another band deferred (``gm_band_merge_local``;
    merge_detections over one band of Detect_OBB.py:291).  ``rec``: padded survivor arrays of ``tile_postprocess(sync=False)``
    (rows at or beyond ``count`` are blanked in place).  Returns (seam records uint8 [(seam_capacity + 1), 80], workspace)
    - the workspace holds the local verdicts for :func:`band_merge_finish`."""
    _require_cuda()
    dev = rec["conf"].device
    boxes, cls, conf = rec["boxes"], rec["cls"], rec["conf"]
    assert boxes.dtype == torch.float64 and boxes.is_contiguous() and cls.dtype == torch.int32 and conf.dtype == torch.float32
    n = int(conf.shape[0])
    count = count.to(device=dev, dtype=torch.int64).contiguous()
    r = np.ascontiguousarray(np.asarray(rects, dtype=np.float32).reshape(-1, 4))
    need = L.lib.gm_band_merge_workspace_bytes(n, int(world), int(seam_capacity), int(edge_capacity))
    ws = _workspace("seam", need, dev)
    out = torch.empty((seam_capacity + 1, BAND_RECORD_BYTES), dtype=torch.uint8, device=dev)
    L.check(L.lib.gm_band_merge_local(_ptr(boxes), _ptr(cls), _ptr(conf), n, _ptr(count), int(max_class), float(iou_thr),
                                      int(edge_capacity), r.ctypes.data_as(C.POINTER(C.c_float)), int(r.shape[0]),
                                      float(extent_bound), int(world), int(seam_capacity), _ptr(out), _ptr(ws), ws.numel(),
                                      _stream()), "gm_band_merge_local")
    return out, ws


def band_merge_finish(gathered: torch.Tensor, world: int, rank: int, seam_capacity: int, rec: dict, max_class: int,
                      iou_thr: float, ws: torch.Tensor, edge_capacity: int = 0, blocks=None, outside_rects=None,
                      extent_bound: float = 0.0, out: Optional[dict] = None) -> dict:
    """Seam verdicts + this rank's kept records in stable confidence order (``gm_band_merge_finish``).  Padded arrays
    ("boxes", "cls", "conf", "angle", "src" = row of ``rec``) and ``meta`` = device int64[4]
    {kept rows, status bits of all ranks, seam rows of all ranks, survivors of all ranks}.  ``blocks`` = (b0, b1): resolve
    only the seam boxes of ranks [b0, b1) with ``outside_rects`` covering the box centres of the ranks left out."""
    _require_cuda()
    dev = rec["conf"].device
    n = int(rec["conf"].shape[0])
    assert gathered.is_cuda and gathered.dtype == torch.uint8 and gathered.is_contiguous()
    assert gathered.numel() == world * (seam_capacity + 1) * BAND_RECORD_BYTES
    angle = rec.get("angle")
    if out is None:
        out = {"boxes": torch.empty((n, 8), dtype=torch.float64, device=dev), "cls": torch.empty(n, dtype=torch.int32, device=dev),
               "conf": torch.empty(n, dtype=torch.float32, device=dev), "src": torch.empty(n, dtype=torch.int32, device=dev),
               "meta": torch.empty(4, dtype=torch.int64, device=dev)}
        if angle is not None:
            out["angle"] = torch.empty(n, dtype=torch.float64, device=dev)
    b0, bn = (int(blocks[0]), int(blocks[1]) - int(blocks[0])) if blocks is not None else (0, 0)
    r = np.ascontiguousarray(np.asarray(outside_rects if outside_rects is not None else [], dtype=np.float32).reshape(-1, 4))
    L.check(L.lib.gm_band_merge_finish(_ptr(gathered), int(world), int(rank), int(seam_capacity), b0, bn,
                                       r.ctypes.data_as(C.POINTER(C.c_float)), int(r.shape[0]), float(extent_bound),
                                       _ptr(rec["boxes"]), _ptr(rec["cls"]),
                                       _ptr(rec["conf"]), _ptr(angle), n, int(max_class), float(iou_thr), int(edge_capacity),
                                       _ptr(out["boxes"]), _ptr(out["cls"]), _ptr(out["conf"]), _ptr(out.get("angle")),
                                       _ptr(out["src"]), _ptr(out["meta"]), _ptr(ws), ws.numel(), _stream()),
            "gm_band_merge_finish")
    return out


# ----------------------------------------------------------------------------- a12: fusion

def fuse_scales(boxes: torch.Tensor, cls: torch.Tensor, conf: torch.Tensor, scale_id: torch.Tensor, n_scales: int,
                max_class: Optional[int] = None, iou_partner: float = 0.40, conf_low: float = 0.25,
                conf_high: float = 0.70, edge_capacity: int = 0) -> torch.Tensor:
    """Kept input indices in output order.  cross_scale_consensus_filter, Detect_OBB.py:347-423."""
    _require_cuda()
    b = _boxes64(boxes)
    n = b.shape[0]
    dev = b.device
    cls = cls.to(torch.int32).contiguous()
    conf = conf.to(torch.float32).contiguous()
    scale_id = scale_id.to(torch.int32).contiguous()
    if max_class is None:
        max_class = int(cls.max().item()) if n else 0
    kept = torch.empty(n, dtype=torch.int32, device=dev)
    cnt = torch.zeros(1, dtype=torch.int64, device=dev)
    cap = edge_capacity
    while True:
        need = L.lib.gm_fuse_workspace_bytes(n, cap)
        ws = _workspace("merge", need, dev)
        L.check(L.lib.gm_fuse_scales(_ptr(b), _ptr(cls), _ptr(conf), _ptr(scale_id), n, int(n_scales), int(max_class),
                                     float(iou_partner), float(conf_low), float(conf_high), cap,
                                     _ptr(kept), _ptr(cnt), _ptr(ws), ws.numel(), _stream()), "gm_fuse_scales")
        k = int(cnt.item())
        if k >= 0:
            return kept[:k]
        cap = -k + 1024


# ----------------------------------------------------------------------------- a5: decode

def letterbox_shape(tile_h: int, tile_w: int, net_size: int, stride: int = 32, auto: bool = True):
    """(new_h, new_w, top, left, out_h, out_w) of Ultralytics' LetterBox for a tile (oracle/letterbox.py)."""
    v = [C.c_int32() for _ in range(6)]
    L.check(L.lib.gm_letterbox_shape(int(tile_h), int(tile_w), int(net_size), int(stride), int(bool(auto)),
                                     *[C.byref(x) for x in v]), "gm_letterbox_shape")
    return tuple(int(x.value) for x in v)


def letterbox_tiles(packed: torch.Tensor, plan: TilePlan, tile_index: torch.Tensor, channels: int, net_size: int,
                    stride: int = 32, auto: bool = True, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Network input of the tiles ``tile_index`` (all of one size) of a packed batch: float32
    [n, C, out_h, out_w] = LetterBox + BGR->RGB (3 channels) + CHW + /255, the predictor's pre-processing
    behind ``model(net_input, conf=...)`` (Detect_OBB.py:76-85), one launch for the whole group."""
    _require_cuda()
    assert packed.is_cuda and packed.dtype == torch.uint8 and channels in (3, 4)
    dev = packed.device
    plan.to(dev)
    idx = tile_index.to(device=dev, dtype=torch.int64)
    n = int(idx.numel())
    recs = plan.dev.view(-1, C.sizeof(L.gm_tile))[idx].contiguous()           # the group's tile records
    hs = plan.tiles["h"][tile_index.cpu().numpy()] if n else np.zeros(0, np.int32)
    ws_ = plan.tiles["w"][tile_index.cpu().numpy()] if n else np.zeros(0, np.int32)
    assert n == 0 or (hs.min() == hs.max() and ws_.min() == ws_.max()), "tiles of one letterbox call share their size"
    th, tw = (int(hs[0]), int(ws_[0])) if n else (1, 1)
    _, _, _, _, oh, ow = letterbox_shape(th, tw, net_size, stride, auto)
    if out is None:
        out = torch.empty((n, channels, oh, ow), dtype=torch.float32, device=dev)
    L.check(L.lib.gm_letterbox_tiles(_ptr(packed), int(channels), _ptr(recs), n, th, tw, int(net_size), int(stride),
                                     int(bool(auto)), _ptr(out), _stream()), "gm_letterbox_tiles")
    return out


def decode_tiles(head: torch.Tensor, plan: TilePlan, net_size, conf_thr: float = 0.25, iou_probiou: float = 0.7,
                 max_det: int = 300):
    """Ultralytics OBB predictor tail for a batch of tiles.  head: float32 [n_tiles, 4+nc+1, A];
    ``net_size``: side of a square network input, or (net_h, net_w) of a rect-letterboxed one.

    Returns (boxes_local [n_tiles*max_det, 8], cls, conf, count [n_tiles]); tile t owns slots
    [t*max_det, t*max_det + count[t]).
    """
    _require_cuda()
    assert head.is_cuda and head.dtype == torch.float32 and head.dim() == 3
    head = head.contiguous()
    nt, C_, A = head.shape
    assert nt == plan.n
    nc = C_ - 5
    dev = head.device
    plan.to(dev)
    boxes = torch.zeros((nt * max_det, 8), dtype=torch.float32, device=dev)
    cls = torch.zeros(nt * max_det, dtype=torch.int32, device=dev)
    conf = torch.zeros(nt * max_det, dtype=torch.float32, device=dev)
    count = torch.zeros(nt, dtype=torch.int32, device=dev)
    need = L.lib.gm_decode_workspace_bytes(nt, A)
    ws = _workspace("decode", need, dev)
    net_h, net_w = (int(net_size), int(net_size)) if np.isscalar(net_size) else (int(net_size[0]), int(net_size[1]))
    L.check(L.lib.gm_decode_tiles(_ptr(head), nt, nc, A, _ptr(plan.dev), net_h, net_w, float(conf_thr),
                                  float(iou_probiou), int(max_det), _ptr(boxes), _ptr(cls), _ptr(conf), _ptr(count),
                                  _ptr(ws), ws.numel(), _stream()), "gm_decode_tiles")
    return boxes, cls, conf, count


def compact_decoded(boxes: torch.Tensor, cls: torch.Tensor, conf: torch.Tensor, count: torch.Tensor, max_det: int):
    """Slots of :func:`decode_tiles` -> flat per-tile lists (tile_id non-decreasing). Index plumbing only."""
    nt = count.numel()
    slot = torch.arange(max_det, device=count.device, dtype=torch.int32).unsqueeze(0)
    valid = (slot < count.unsqueeze(1)).reshape(-1)
    tile_id = torch.arange(nt, device=count.device, dtype=torch.int32).repeat_interleave(max_det)
    return boxes[valid], cls[valid], conf[valid], tile_id[valid]
