"""Multi-GPU path: one process per GPU, maps sharded by tile-row band.

Tiles are independent for every pixel stage and for the per-tile NMS, so a rank needs only the
map rows of its band (band rows + ``overlap`` extra) and no halo exchange.  Cross-tile coupling
exists only in the fusion and the global merge (Detect_OBB.py:291, :176-200).

Seam-band exchange (:func:`merge_bands_seam_device`, the path ``bench.py`` runs): every rank resolves
the class-wise greedy NMS of its OWN band; a box whose verdict can depend on another band - it may
overlap a foreign box, or a higher-priority neighbour of it is such a box and nothing kept
suppresses it first - is deferred.  Only the deferred boxes (the seam band) travel, in ONE
fixed-capacity all_gather (NCCL over NVLink on GPUs, gloo in the CPU tests); every rank resolves the
gathered seam set (identical on all ranks) and applies the verdicts to its own rows.  Per-rank work
is its band plus the seam set, not ``world`` bands.  The union of the ranks' kept lists, merged by
(confidence desc, rank, local order), is the single-rank result - members and order.

Full exchange (:func:`merge_bands_device`, round 1, kept as the reference formulation and as the
fallback when a seam bound is exceeded): all survivors are gathered and the NMS is sharded by class.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional, Tuple

import time

import torch
import torch.distributed as dist

_PROFILE: dict = {}


def band_rows(total_rows: int, world: int, rank: int) -> Tuple[int, int]:
    """Tile rows [r0, r1) of ``rank``: an even split, the first ``total_rows % world`` ranks get one more."""
    base, extra = divmod(total_rows, world)
    r0 = rank * base + min(rank, extra)
    return r0, r0 + base + (1 if rank < extra else 0)


def band_pixel_rows(H: int, tile_size: int, overlap: int, r0: int, r1: int) -> Tuple[int, int]:
    """Map rows [y0, y1) a band of tile rows touches."""
    step = max(1, tile_size - overlap)
    if r1 <= r0:
        return 0, 0
    return r0 * step, min((r1 - 1) * step + tile_size, H)


# ----------------------------------------------------------------------------- host placement

def parse_cpulist(text: str) -> list:
    """Linux cpulist syntax ("0-15,32-47") -> sorted list of CPU numbers."""
    cpus = set()
    for part in text.strip().split(","):
        part = part.strip()
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-", 1)
            cpus.update(range(int(a), int(b) + 1))
        else:
            cpus.add(int(part))
    return sorted(cpus)


def gpu_local_cpus(device_index: int, sysfs_root: str = "/sys") -> Optional[list]:
    """CPUs of the NUMA node the GPU's PCIe root port hangs off (sysfs ``local_cpulist``), or None if unknown."""
    import os
    try:
        p = torch.cuda.get_device_properties(device_index)
        bdf = f"{int(p.pci_domain_id):04x}:{int(p.pci_bus_id):02x}:{int(p.pci_device_id):02x}.0"
    except Exception:
        return None
    try:
        with open(os.path.join(sysfs_root, "bus", "pci", "devices", bdf, "local_cpulist")) as fh:
            cpus = parse_cpulist(fh.read())
        return cpus or None
    except (OSError, ValueError):
        return None


def bind_host_to_gpu(device_index: int, sysfs_root: str = "/sys", cpus: Optional[list] = None) -> dict:
    """Pin the calling process to the CPUs local to its GPU, BEFORE it allocates pinned host buffers.

    One process per GPU on a two-socket box: without a binding Linux places a rank's pinned map buffer on whatever
    node the process happened to run on, and every rank whose GPU sits on the other socket drags its 200 MB per step
    across the socket interconnect - the N = 8 end-to-end number of round 1 (120 GB/s aggregate against ~55 GB/s per
    GPU link) is that limit, not PCIe.  With the affinity set first, first-touch puts the buffer on the GPU's node.
    Never raises: returns what it did ({"cpus": n, "first": c0, "last": c1} or {"unchanged": reason})."""
    import os
    try:
        if cpus is None:
            cpus = gpu_local_cpus(device_index, sysfs_root)
        if not cpus:
            return {"unchanged": "no local_cpulist for the device"}
        allowed = os.sched_getaffinity(0)
        use = sorted(set(cpus) & allowed)
        if not use:
            return {"unchanged": "local CPUs outside the allowed set"}
        if set(use) == set(allowed):
            return {"unchanged": "already local", "cpus": len(use)}
        os.sched_setaffinity(0, use)
        return {"cpus": len(use), "first": use[0], "last": use[-1]}
    except Exception as e:      # noqa: BLE001 - placement is an optimisation, never a failure
        return {"unchanged": f"{type(e).__name__}: {e}"}


# ----------------------------------------------------------------------------- seam-band exchange

def tile_range(total_tiles: int, world: int, rank: int) -> Tuple[int, int]:
    """Tiles [t0, t1) of ``rank`` in row-major order: an even split of the TILES (a row band whose first and last tile
    row may be partial), so every rank builds the same number of tiles whatever the number of tile rows."""
    base, extra = divmod(total_tiles, world)
    t0 = rank * base + min(rank, extra)
    return t0, t0 + base + (1 if rank < extra else 0)


def foreign_center_rects(H: int, W: int, tile_size: int, overlap: int, t_begin: int, t_end: int, margin: int) -> list:
    """Closed rectangles (x0, y0, x1, y1) covering every position the CENTRE of a detection of another rank can take,
    for the rank that owns tiles [t_begin, t_end) of the row-major plan.  A detection survives the border filter only if
    its centre lies in the safe region of its tile, ``margin <= c - tile origin <= tile dim - margin`` (Detect_OBB.py:
    167-174, :242-249); the union of the safe regions of the tiles before (after) the range is covered by at most two
    rectangles: the full tile rows and the partial row.  With the filter off (``margin <= 0``) centres are unconstrained:
    one rectangle covering everything, i.e. every box is a seam candidate (still exact, nothing saved)."""
    step = max(1, tile_size - overlap)
    rows, cols = -(-H // step), -(-W // step)
    n = rows * cols
    if t_begin <= 0 and t_end >= n:
        return []
    if margin <= 0:
        return [(-3.0e38, -3.0e38, 3.0e38, 3.0e38)]
    m = float(margin)

    def row_y(r):
        y0 = r * step
        return y0, min(y0 + tile_size, H)

    def col_x(c):
        x0 = c * step
        return x0, min(x0 + tile_size, W)

    rects = []
    rb, cb = divmod(max(t_begin, 0), cols)
    if rb > 0:                                   # full tile rows above
        rects.append((m, m, W - m, row_y(rb - 1)[1] - m))
    if cb > 0:                                   # tiles to the left in the first (partial) tile row
        y0, y1 = row_y(rb)
        rects.append((m, y0 + m, col_x(cb - 1)[1] - m, y1 - m))
    re, ce = divmod(min(t_end, n), cols)
    if ce > 0:                                   # tiles to the right in the last (partial) tile row
        y0, y1 = row_y(re)
        rects.append((col_x(ce)[0] + m, y0 + m, W - m, y1 - m))
        re += 1
    if re < rows:                                # full tile rows below
        rects.append((m, row_y(re)[0] + m, W - m, H - m))
    return rects


def seam_scope(H: int, W: int, tile_size: int, overlap: int, margin: int, world: int, rank: int, ranges=None, reach: int = 1) -> dict:
    """The part of the gathered seam set a rank resolves: the blocks of ranks [rank - reach, rank + reach] and the
    rectangles covering the box centres of every rank OUTSIDE that range (``foreign_center_rects`` of the union tile range).
    ``ranges`` = the (t0, t1) of every rank (default: :func:`tile_range` of the plan).  With the default reach a rank's work
    on the seam set is three blocks whatever the number of ranks; a chain of overlaps that leaves the range is detected
    (``GM_SEAM_CHAIN_ESCAPES``) and that rank alone falls back to all blocks - no collective, it holds every record."""
    step = max(1, tile_size - overlap)
    n = (-(-H // step)) * (-(-W // step))
    if ranges is None:
        ranges = [tile_range(n, world, r) for r in range(world)]
    b0, b1 = max(0, rank - reach), min(world, rank + reach + 1)
    if b0 == 0 and b1 == world:
        return {"blocks": (0, world), "rects": []}
    return {"blocks": (b0, b1), "rects": foreign_center_rects(H, W, tile_size, overlap, ranges[b0][0], ranges[b1 - 1][1], margin)}


def box_reach(boxes: torch.Tensor) -> torch.Tensor:
    """Chebyshev distance of the farthest corner from the centre the border filter tests (mean of the four corners,
    Detect_OBB.py:159-165), per box: the quantity ``extent_bound`` of the seam exchange bounds."""
    b = boxes.to(torch.float64)
    cx = 0.25 * ((b[:, 0] + b[:, 2]) + (b[:, 4] + b[:, 6]))
    cy = 0.25 * ((b[:, 1] + b[:, 3]) + (b[:, 5] + b[:, 7]))
    return torch.maximum((b[:, 0::2] - cx[:, None]).abs().amax(1), (b[:, 1::2] - cy[:, None]).abs().amax(1))


def agree_seam_bounds(rec: Dict[str, torch.Tensor], count: torch.Tensor, iou_thr: float, max_class: int, rects,
                      slack: float = 1.25, group=None) -> Tuple[float, int]:
    """(extent_bound, seam_capacity) for :func:`merge_bands_seam_device`, agreed ONCE from a representative batch (set-up
    time: two small all_reduce(MAX) and two host reads): the largest box reach over all ranks, and the largest number of
    boxes any rank defers with that bound, both with ``slack``.  Every later call verifies them (status bits)."""
    world, _ = _world(group)
    dev = rec["conf"].device
    n = int(count.reshape(-1)[0].item())
    reach = box_reach(rec["boxes"][:max(n, 0)]).max().reshape(1) if n > 0 else torch.zeros(1, dtype=torch.float64, device=dev)
    reach = reach.to(torch.float64)
    if world > 1:
        dist.all_reduce(reach, op=dist.ReduceOp.MAX, group=group)
    bound = float(reach.item()) * 1.02 + 1.0
    from . import ops
    with ops.workspace_scope("agree"):
        clone = {k: v.clone() for k, v in rec.items() if k in ("boxes", "cls", "conf")}
        cap = int(rec["conf"].shape[0])
        send, _ = ops.band_merge_local(clone, count, max_class, iou_thr, rects, bound, world, cap)
        deferred = send[0].view(torch.int64)[0:1].clone()
    if world > 1:
        dist.all_reduce(deferred, op=dist.ReduceOp.MAX, group=group)
    ops._workspaces.pop((str(dev), "agree/seam"), None)
    return bound, int(int(deferred.item()) * slack) + 1024


def seam_candidates(boxes: torch.Tensor, rects, bound: float) -> torch.Tensor:
    """Tensor mirror of ``k_seam_candidates`` (gloo tests on CPU tensors): the outward-rounded fp32 AABB of a box, grown
    by ``bound``, reaches a rectangle of foreign centres."""
    b = boxes.to(torch.float64)
    lo = torch.stack([b[:, 0::2].amin(1), b[:, 1::2].amin(1)], 1)
    hi = torch.stack([b[:, 0::2].amax(1), b[:, 1::2].amax(1)], 1)
    lo32, hi32 = lo.to(torch.float32), hi.to(torch.float32)
    lo32 = torch.where(lo32.to(torch.float64) > lo, torch.nextafter(lo32, torch.full_like(lo32, -float("inf"))), lo32)
    hi32 = torch.where(hi32.to(torch.float64) < hi, torch.nextafter(hi32, torch.full_like(hi32, float("inf"))), hi32)
    bound32 = torch.tensor(bound, dtype=torch.float32)
    c = torch.zeros(b.shape[0], dtype=torch.bool, device=b.device)
    for (x0, y0, x1, y1) in rects:
        r = torch.tensor([x0, y0, x1, y1], dtype=torch.float32)
        c |= ((lo32[:, 0] - bound32) <= r[2]) & (r[0] <= (hi32[:, 0] + bound32)) & \
             ((lo32[:, 1] - bound32) <= r[3]) & (r[1] <= (hi32[:, 1] + bound32))
    return c


def merge_bands_seam_device(rec: Dict[str, torch.Tensor], count: torch.Tensor, seam_capacity: int, iou_thr: float,
                            max_class: int, rects, extent_bound: float, edge_capacity: int = 0,
                            local_fn: Optional[Callable] = None, seam_fn: Optional[Callable] = None,
                            group=None, scope: Optional[dict] = None) -> Dict[str, torch.Tensor]:
    """Device part of the cross-band merge with a SEAM-BAND exchange: fixed shapes, no host read (capturable in a CUDA
    graph), ONE collective.

    ``rec``: this rank's per-tile-NMS survivors, padded arrays with ``count`` valid rows (``ops.tile_postprocess(
    sync=False)``); rows beyond are blanked in place.  ``rects`` = :func:`foreign_center_rects` of this rank's tile range;
    ``extent_bound`` >= the larger AABB side of any box on any rank and ``seam_capacity`` >= the deferred boxes of any
    rank - both agreed once (e.g. from a first pass) and verified by every call: a violated bound comes back as a status
    bit in ``meta`` and the result must be recomputed (``merge_bands_device`` is the bound-free formulation).

    ``scope`` (:func:`seam_scope`): resolve only the seam boxes of the neighbouring ranks; exactness is guarded by a
    second deferral (a chain of overlaps that leaves the neighbourhood sets a status bit and :func:`merge_bands_seam_finish`
    repeats the seam phase on all blocks, locally).

    Returns this rank's kept records, padded: "boxes", "cls", "conf", "angle", "src" (row of ``rec``) and ``meta`` =
    int64[4] {kept rows, status bits of all ranks, seam rows of all ranks, survivors of all ranks}.

    CPU tensors (gloo tests): ``local_fn(boxes, cls, conf, candidate) -> (order, state)`` must implement the deferring
    greedy NMS (state 1 kept / 2 suppressed / 3 deferred; order = stable confidence-descending permutation) and
    ``seam_fn(boxes, cls, conf) -> (order, keep)`` the plain one; there is no CPU implementation in this package."""
    world, rank = _world(group)
    dev = rec["conf"].device
    if rec["conf"].is_cuda and local_fn is None:
        from . import ops
        send, ws = ops.band_merge_local(rec, count, max_class, iou_thr, rects, extent_bound, world, seam_capacity, edge_capacity)
        if world > 1:
            recv = torch.empty((world * (seam_capacity + 1), send.shape[1]), dtype=torch.uint8, device=dev)
            dist.all_gather_into_tensor(recv, send, group=group)
        else:
            recv = send
        blocks, orects = (scope["blocks"], scope["rects"]) if scope is not None and tuple(scope["blocks"]) != (0, world) else (None, None)
        out = ops.band_merge_finish(recv, world, rank, seam_capacity, rec, max_class, iou_thr, ws, edge_capacity,
                                    blocks=blocks, outside_rects=orects, extent_bound=extent_bound)
        # what a local repeat of the seam phase needs (merge_bands_seam_finish, chain-escape fallback)
        out["_redo"] = (recv, world, rank, seam_capacity, rec, max_class, iou_thr, ws, edge_capacity)
        return out
    if local_fn is None or seam_fn is None:
        raise RuntimeError("merge_bands_seam_device on CPU tensors needs local_fn and seam_fn (tests); the product path is CUDA")
    n = int(rec["conf"].shape[0])
    cnt = int(count.reshape(-1)[0].item())
    valid = torch.arange(n) < max(cnt, 0)
    boxes = torch.where(valid[:, None], rec["boxes"].to(torch.float64), torch.full((n, 8), float("nan"), dtype=torch.float64))
    cls = torch.where(valid, rec["cls"].to(torch.int32), torch.full((n,), -1, dtype=torch.int32))
    conf = torch.where(valid, rec["conf"].to(torch.float32), torch.full((n,), float("-inf"), dtype=torch.float32))
    cand = seam_candidates(boxes, rects, extent_bound) & valid
    order, state = local_fn(boxes, cls, conf, cand)
    state = torch.where(valid, state.to(torch.uint8), torch.full((n,), 2, dtype=torch.uint8))
    deferred = torch.nonzero(state == 3).squeeze(1)                      # ascending row = list order
    ext = torch.where(valid, box_reach(boxes), torch.zeros(n, dtype=torch.float64))
    status = (4 if deferred.numel() > seam_capacity else 0) | (8 if n and float(ext.max()) > extent_bound else 0) | (16 if cnt < 0 else 0)
    k = min(int(deferred.numel()), seam_capacity)
    # the same record layout as the kernels: 10 eight-byte words per row, row 0 = header
    send = torch.zeros((seam_capacity + 1, 10), dtype=torch.int64)
    send[0, 0], send[0, 1], send[0, 3] = int(deferred.numel()), status, max(cnt, 0)
    send[1:, :8] = torch.full((seam_capacity, 8), float("nan"), dtype=torch.float64).view(torch.int64)
    send[1:, 8] = -1
    send[1:, 9] = -1
    d = deferred[:k]
    send[1:k + 1, :8] = boxes[d].view(torch.int64)
    packed = torch.stack([cls[d], conf[d].view(torch.int32)], 1).contiguous().view(torch.int64).reshape(-1)
    send[1:k + 1, 8] = packed
    send[1:k + 1, 9] = d.to(torch.int64)
    if world > 1:
        recv = torch.empty((world * (seam_capacity + 1), 10), dtype=torch.int64)
        dist.all_gather_into_tensor(recv, send, group=group)
    else:
        recv = send
    blocks = recv.reshape(world, seam_capacity + 1, 10)
    head, rows = blocks[:, 0], blocks[:, 1:].reshape(world * seam_capacity, 10)
    u_boxes = rows[:, :8].contiguous().view(torch.float64)
    cc = rows[:, 8].contiguous().view(torch.int32).reshape(-1, 2)
    u_cls, u_conf, u_src = cc[:, 0].contiguous(), cc[:, 1].contiguous().view(torch.float32), rows[:, 9]
    chain = 0
    if u_cls.numel():
        mine = slice(rank * seam_capacity, (rank + 1) * seam_capacity)
        live = u_cls[mine] >= 0
        keep_mine = None
        if scope is not None and tuple(scope["blocks"]) != (0, world):
            b0, b1 = scope["blocks"]
            sub = slice(b0 * seam_capacity, b1 * seam_capacity)
            taint = seam_candidates(u_boxes[sub], scope["rects"], extent_bound) & (u_cls[sub] >= 0)
            _, st_sub = local_fn(u_boxes[sub], u_cls[sub], u_conf[sub], taint)
            st_mine = st_sub[(rank - b0) * seam_capacity:(rank - b0 + 1) * seam_capacity]
            if bool((st_mine[live] == 3).any()):
                chain = 32                               # a chain of overlaps leaves the neighbourhood: all blocks (local)
            else:
                keep_mine = (st_mine == 1)
        if keep_mine is None:
            _, keep_u = seam_fn(u_boxes, u_cls, u_conf)
            keep_mine = keep_u[mine].to(torch.bool)
        state[u_src[mine][live]] = torch.where(keep_mine[live], torch.tensor(1, dtype=torch.uint8), torch.tensor(2, dtype=torch.uint8))
    order = order.to(torch.int64)
    kept = order[(state[order] == 1)]
    m = int(kept.numel())
    pad = torch.zeros(n - m, dtype=torch.int64)
    idx = torch.cat([kept, pad])
    out = {k2: rec[k2][idx] for k2 in ("boxes", "cls", "conf", "angle") if k2 in rec}
    out["src"] = idx.to(torch.int32)
    st_all = 0
    for v in head[:, 1].tolist():
        st_all |= int(v)
    out["meta"] = torch.tensor([m, st_all, int(head[:, 0].sum()), int(head[:, 3].sum())], dtype=torch.int64)
    out["_chain_fallbacks"] = 1 if chain else 0
    return out


def merge_bands_seam_finish(dev_out: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Host part: the ONE host read (4 integers), the bound checks and the slicing of the padded arrays."""
    m, status, n_seam, n_valid = (int(v) for v in dev_out["meta"].tolist())
    fallbacks = int(dev_out.get("_chain_fallbacks", 0))
    if status == 32 and "_redo" in dev_out:
        # only this rank's restricted seam phase was inconclusive: repeat it on the records of all ranks (already here)
        from . import ops
        recv, world, rank, cap, rec, max_class, iou_thr, ws, edge_cap = dev_out["_redo"]
        keep = {k: v for k, v in dev_out.items() if not k.startswith("_")}
        ops.band_merge_finish(recv, world, rank, cap, rec, max_class, iou_thr, ws, edge_cap, out=keep)
        m, status, n_seam, n_valid = (int(v) for v in keep["meta"].tolist())
        fallbacks = 1
    if status:
        from . import ops
        raise SeamBoundExceeded(status, ops.seam_status_text(status))
    out = {k: v[:m] for k, v in dev_out.items() if k != "meta" and not k.startswith("_")}
    out["n_valid"], out["n_seam"], out["chain_fallbacks"] = n_valid, n_seam, fallbacks
    return out


class SeamBoundExceeded(RuntimeError):
    """A bound of the seam exchange (seam_capacity, extent_bound, edge_capacity) did not hold for this input; every rank
    raises alike (the status travels with the exchange).  Rerun with the bound raised, or use ``merge_bands_padded``."""

    def __init__(self, status: int, text: str):
        super().__init__(f"seam exchange bound exceeded ({status}): {text}")
        self.status = status


def gather_merged(out: Dict[str, torch.Tensor], group=None) -> Dict[str, torch.Tensor]:
    """Every rank's kept records (``merge_bands_seam_finish``) -> the ONE merged list in the reference's order on every
    rank: stable confidence-descending over the concatenation in rank order (= list order).  A second collective, for
    consumers that need the whole list in one place (rendering, Excel); not part of the hot path."""
    keys = [k for k in ("boxes", "cls", "conf", "angle") if k in out]
    got = allgather_records({k: out[k] for k in keys}, group=group)
    order = torch.sort(got["conf"], descending=True, stable=True)[1]
    return {k: got[k][order] for k in keys}


def _world(group=None) -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def allgather_records(rec: Dict[str, torch.Tensor], capacity: Optional[int] = None, group=None) -> Dict[str, torch.Tensor]:
    """Concatenate every rank's detection records in rank order.

    Two collectives: the per-rank counts, then ONE padded all_gather of the records packed row-wise
    into bytes (all fields of a record side by side).  The padded length is agreed from the gathered
    counts, so ranks never disagree on the message size; ``capacity`` is only an optional upper bound
    that raises (on every rank alike) when some rank exceeds it.  ``rec`` fields share their first
    dimension."""
    world, _ = _world(group)
    n = next(iter(rec.values())).shape[0]
    if world == 1:
        return {k: v for k, v in rec.items()}
    dev = next(iter(rec.values())).device
    counts_t = torch.empty(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts_t, torch.tensor([n], dtype=torch.int64, device=dev), group=group)
    counts = [int(c) for c in counts_t.tolist()]
    if capacity is not None and max(counts) > capacity:
        raise ValueError(f"{max(counts)} records on one rank exceed the all_gather capacity {capacity}")
    cap = max(max(counts), 1)
    # pack: every field viewed as bytes [n, row_bytes], concatenated along dim 1
    cols, layout = [], []
    for k, v in rec.items():
        flat = v.contiguous().reshape(n, -1) if n else v.reshape(0, int(torch.tensor(v.shape[1:]).prod()) if v.dim() > 1 else 1)
        b = flat.view(torch.uint8).reshape(n, -1) if n else torch.empty((0, flat.shape[1] * v.element_size()), dtype=torch.uint8, device=dev)
        layout.append((k, v.dtype, tuple(v.shape[1:]), b.shape[1]))
        cols.append(b)
    row_bytes = sum(l[3] for l in layout)
    send = torch.zeros((cap, row_bytes), dtype=torch.uint8, device=dev)
    if n:
        send[:n] = torch.cat(cols, dim=1)
    recv = torch.empty((world * cap, row_bytes), dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(recv, send, group=group)
    rows = torch.cat([recv[r * cap:r * cap + c] for r, c in enumerate(counts)], dim=0)
    out, off = {}, 0
    for k, dtype, tail, nb in layout:
        out[k] = rows[:, off:off + nb].contiguous().view(dtype).reshape((rows.shape[0],) + tail)
        off += nb
    return out


def merge_sharded_by_class(boxes: torch.Tensor, cls: torch.Tensor, conf: torch.Tensor, iou_thr: float, max_class: int,
                           nms_fn: Optional[Callable] = None, group=None) -> torch.Tensor:
    """Global class-wise NMS of records every rank holds identically; rank r resolves the classes
    c with c % world == r.  Returns the kept indices in the reference's output order (stable
    confidence-descending), identical on every rank."""
    world, rank = _world(group)
    if nms_fn is None:
        from . import ops
        nms_fn = lambda b, c, f: ops.nms_global(b, c, f, iou_thr, max_class=max_class)[2]   # noqa: E731
    n = boxes.shape[0]
    dev = boxes.device
    if world == 1:
        return nms_fn(boxes, cls, conf).to(torch.int64)
    mine = torch.nonzero((cls % world) == rank).squeeze(1)          # ascending: keeps list order for ties
    kept_local = nms_fn(boxes[mine], cls[mine], conf[mine]).to(torch.int64) if mine.numel() else mine
    kept = mine[kept_local]
    got = allgather_records({"idx": kept}, group=group)
    allk = got["idx"]
    allk, _ = torch.sort(allk)                                       # list order first ...
    order = torch.sort(conf[allk], descending=True, stable=True)[1]  # ... then stable by confidence
    return allk[order]


def merge_bands_device(rec: Dict[str, torch.Tensor], count: torch.Tensor, capacity: int, iou_thr: float, max_class: int,
                       nms_fn: Optional[Callable] = None, group=None) -> Dict[str, torch.Tensor]:
    """Device part of the cross-band merge: fixed shapes, no host read, so the whole call can be captured in a
    CUDA graph (:class:`CapturedCall`).

    ``rec`` holds this rank's per-tile-NMS survivors in arrays of length >= ``capacity`` of which the first
    ``count`` (device int64[1]) rows are valid (``ops.tile_postprocess(sync=False)``); ``capacity`` must be the
    same on every rank (e.g. the largest per-rank input size, agreed once).  Steps:
      1. rows >= count are blanked (NaN corners, class -1, confidence -inf) and every field is packed into one
         byte matrix; one fixed-size all_gather (no count exchange: nobody needs the counts on the host);
      2. every rank runs the exact class-wise NMS over ALL world*capacity rows with the classes it does not own
         masked to -1 (such rows, like the blank ones, fall into a group nobody queries);
      3. the keep flags of the owned classes are combined with one all_reduce(MAX);
      4. the kept rows are compacted in stable confidence-descending order by a prefix sum + scatter.
    Returns padded arrays of world*capacity rows ("boxes", "cls", "conf", "angle", "index" = position
    r*capacity + i) whose first ``meta[0]`` rows are the merged records, and ``meta`` = device int64[3]:
    {kept rows, status (negative: a pair buffer overflowed), survivors of all bands that entered the merge}.
    """
    world, rank = _world(group)
    keys = [k for k in ("boxes", "cls", "conf", "angle") if k in rec]
    dev = rec["conf"].device
    prof = _PROFILE.get("sink")                     # debugging aid: {"sink": list} makes every section synchronous and timed

    def _mark(name):
        if prof is not None:
            torch.cuda.synchronize()
            prof.append((name, time.perf_counter()))

    _mark("start")
    if nms_fn is None and rec["conf"].is_cuda:
        # device path: one kernel each for blank + pack, unpack + class mask, and the ordered compaction
        from . import ops
        send = ops.band_pack(rec, count, capacity)
        _mark("pack")
        if world > 1:
            recv = torch.empty((world * capacity, send.shape[1]), dtype=torch.uint8, device=dev)
            dist.all_gather_into_tensor(recv, send, group=group)
        else:
            recv = send
        _mark("all_gather")
        u = ops.band_unpack(recv, world, rank, with_angle="angle" in rec)
        _mark("unpack")
        order, keep, _, n_kept = ops.nms_global(u["boxes"], u["cls_owned"], u["conf"], iou_thr, max_class=max_class, sync=False)
        _mark("nms")
        ops.band_mask_keep(keep, u["cls_owned"])
        if world > 1:
            dist.all_reduce(keep, op=dist.ReduceOp.MAX, group=group)
        _mark("all_reduce")
        x = ops.band_extract(order, keep, u)
        out = {k: x[k] for k in keys}
        out["index"] = x["index"]
        status = torch.minimum(n_kept.reshape(1).to(torch.int64), count.reshape(1).to(device=dev, dtype=torch.int64))
        out["meta"] = torch.cat([x["n_out"], status, u["n_valid"]])
        _mark("extract")
        return out
    valid = torch.arange(capacity, device=dev) < count.to(dev)
    fields = {}
    for k in keys:
        v = rec[k][:capacity]
        if v.shape[0] < capacity:                                   # shorter input buffer: pad up to the agreed capacity
            v = torch.cat([v, torch.zeros((capacity - v.shape[0],) + tuple(v.shape[1:]), dtype=v.dtype, device=dev)])
        m = valid.reshape((-1,) + (1,) * (v.dim() - 1))
        blank = float("nan") if k == "boxes" else (-1 if k == "cls" else (float("-inf") if k == "conf" else 0.0))
        fields[k] = torch.where(m, v, torch.full_like(v, blank))
    _mark("blank")
    if world > 1:
        cols = [fields[k].contiguous().reshape(capacity, -1).view(torch.uint8).reshape(capacity, -1) for k in keys]
        widths = [c.shape[1] for c in cols]
        send = torch.cat(cols, dim=1).contiguous()
        recv = torch.empty((world * capacity, send.shape[1]), dtype=torch.uint8, device=dev)
        _mark("pack")
        dist.all_gather_into_tensor(recv, send, group=group)
        _mark("all_gather")
        off = 0
        for k, wd in zip(keys, widths):
            tail = tuple(fields[k].shape[1:])
            fields[k] = recv[:, off:off + wd].contiguous().view(fields[k].dtype).reshape((world * capacity,) + tail)
            off += wd
    _mark("unpack")
    cls_all = fields["cls"]
    mine = (cls_all >= 0) & ((cls_all % world) == rank)
    cls_mine = torch.where(mine, cls_all, torch.full_like(cls_all, -1))
    if nms_fn is None:
        from . import ops
        order, keep, _, n_kept = ops.nms_global(fields["boxes"], cls_mine, fields["conf"], iou_thr, max_class=max_class, sync=False)
    else:
        order, keep = nms_fn(fields["boxes"], cls_mine, fields["conf"])
        n_kept = torch.zeros(1, dtype=torch.int64, device=dev)
    _mark("nms")
    keep = (keep.to(torch.uint8) * mine.to(torch.uint8)).contiguous()
    if world > 1:
        dist.all_reduce(keep, op=dist.ReduceOp.MAX, group=group)
    _mark("all_reduce")
    order = order.to(torch.int64)
    total = order.shape[0]
    flags = keep[order].to(torch.bool)
    pos = torch.cumsum(flags.to(torch.int64), 0) - 1
    slot = torch.where(flags, pos, torch.full_like(pos, total))      # dropped rows all land in the spare last slot
    kept_pad = torch.zeros(total + 1, dtype=torch.int64, device=dev)
    kept_pad.scatter_(0, slot, order)
    index = kept_pad[:total]
    out = {k: fields[k][index] for k in keys}
    out["index"] = index
    status = torch.minimum(n_kept.reshape(1).to(torch.int64), count.reshape(1).to(torch.int64))
    out["meta"] = torch.cat([flags.sum().reshape(1), status, (cls_all >= 0).sum().reshape(1)])
    _mark("extract")
    return out


def merge_bands_finish(dev_out: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Host part of the merge: the ONE host read (3 integers) and the slicing of the padded arrays."""
    m, status, n_valid = (int(v) for v in dev_out["meta"].tolist())
    if status < 0:
        raise RuntimeError("pair buffer overflow in the padded merge: rerun with a larger edge capacity")
    out = {k: v[:m] for k, v in dev_out.items() if k != "meta"}
    out["n_valid"] = n_valid
    return out


def merge_bands_padded(rec: Dict[str, torch.Tensor], count: torch.Tensor, capacity: int, iou_thr: float, max_class: int,
                       nms_fn: Optional[Callable] = None, group=None) -> Dict[str, torch.Tensor]:
    """The whole cross-band merge with ONE host read: :func:`merge_bands_device` then :func:`merge_bands_finish`.
    Members and order equal the single-rank ``merge_detections`` of the concatenated band lists.
    Returns the kept records (same keys as ``rec`` minus bookkeeping) plus "index" (position r*capacity + i)."""
    return merge_bands_finish(merge_bands_device(rec, count, capacity, iou_thr, max_class, nms_fn, group))


class CapturedCall:
    """``fn()`` recorded once into a CUDA graph and replayed: the detection path of a step is ~125 launches of
    small kernels plus a few dozen tensor ops, so the launch overhead (host and device side) is most of its
    time; a replay issues them as one graph.  ``fn`` must be free of host reads and use fixed shapes
    (``ops.tile_postprocess(sync=False)`` + :func:`merge_bands_device`); it reads its inputs from tensors the
    caller keeps and overwrites in place.  Capture happens on ``stream`` (its priority is recorded in the kernel
    nodes) inside a private workspace scope.  If the capture fails (e.g. a driver that cannot capture one of the
    calls) the object falls back to calling ``fn`` eagerly and says so in ``captured``."""

    def __init__(self, fn: Callable[[], Dict[str, torch.Tensor]], stream: Optional["torch.cuda.Stream"] = None,
                 warmup: int = 2):
        from . import ops
        self.fn = fn
        self.scope = f"graph{id(self):x}"
        self.stream = stream if stream is not None else torch.cuda.Stream()
        self.captured = False
        self.error = None
        self.launches = 0
        self.stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.stream), ops.workspace_scope(self.scope):
            for _ in range(max(1, warmup)):                      # lazy one-time initialisation happens here, not under capture
                self.out = fn()
        self.stream.synchronize()
        from . import _lib
        try:
            self.graph = torch.cuda.CUDAGraph()
            before = _lib.lib.gm_launch_count()
            with ops.workspace_scope(self.scope), torch.cuda.graph(self.graph, stream=self.stream):
                self.out = fn()
            self.launches = int(_lib.lib.gm_launch_count() - before)    # library kernels per replay (bench.py gpu_launches)
            self.captured = True
        except Exception as e:                                      # noqa: BLE001 - any capture failure means "run eagerly"
            self.error = repr(e)
            self.graph = None
            torch.cuda.synchronize()
        self._keep = ops.workspaces_of(self.scope)

    def __call__(self) -> Dict[str, torch.Tensor]:
        """Enqueue one run on the current stream; returns the (static) output tensors."""
        from . import ops
        if self.captured:
            self.graph.replay()
            return self.out
        with ops.workspace_scope(self.scope):
            self.out = self.fn()
        return self.out
