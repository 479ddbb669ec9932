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


def allgather_records(rec: Dict[str, torch.Tensor], capacity: int, group=None) -> Dict[str, torch.Tensor]:
    """Concatenate every rank's detection records in rank order (one padded all_gather per field
    plus one for the counts).  ``rec`` fields share their first dimension; ``capacity`` bounds it."""
    world, _ = _world(group)
    n = next(iter(rec.values())).shape[0]
    if world == 1:
        return {k: v for k, v in rec.items()}
    if n > capacity:
        raise ValueError(f"{n} records exceed the all_gather capacity {capacity}")
    dev = next(iter(rec.values())).device
    counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([n], dtype=torch.int64, device=dev), group=group)
    counts = [int(c.item()) for c in counts]
    out = {}
    for k, v in rec.items():
        pad = torch.zeros((capacity,) + tuple(v.shape[1:]), dtype=v.dtype, device=dev)
        pad[:n] = v
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad, group=group)
        out[k] = torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)
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
    pad = torch.full((n,), -1, dtype=torch.int64, device=dev)
    pad[:kept.numel()] = kept
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    allk = torch.cat(parts)
    allk = allk[allk >= 0]
    allk, _ = torch.sort(allk)                                       # list order first ...
    order = torch.sort(conf[allk], descending=True, stable=True)[1]  # ... then stable by confidence
    return allk[order]
