"""Multi-GPU path: one process per GPU, maps sharded by tile-row band.

Tiles are independent for every pixel stage and for the per-tile NMS, so a rank needs only the
map rows of its band (band rows + ``overlap`` extra) and no halo exchange.  Cross-tile coupling
exists only in the fusion and the global merge: the survivors of all bands are exchanged with
one fixed-capacity all_gather (NCCL over NVLink on GPUs, gloo in the CPU tests), the class-wise
global NMS is sharded by class (classes are independent, Detect_OBB.py:193), and the per-class
keep lists are exchanged with a second all_gather.  The merged result is identical - members
and order - to the single-rank result.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional, Tuple

import torch
import torch.distributed as dist


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
