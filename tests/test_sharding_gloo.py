"""world_size-2 and world_size-4 gloo runs of the multi-GPU host logic on CPU: band split, padded all_gather of
detection records, class-sharded global merge - against the single-rank oracle result."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import geom_c
        from oriented_object_detection_b200 import sharding, synth
        boxes, cls, conf = synth.synthetic_obbs(400, 1500, 1500, n_classes=5, seed=9)
        conf[::11] = conf[5]
        n = len(conf)
        # every rank owns a contiguous slice of the list (its band's survivors)
        # (unequal slices: ranks must agree on the message size by themselves)
        cut = [0, n // 3, n] if world == 2 else [sharding.band_rows(n, world, r)[0] for r in range(world)] + [n]
        r0, r1 = cut[rank], cut[rank + 1]
        angle = np.arange(n, dtype=np.float64) * 0.5
        rec = {"boxes": torch.from_numpy(boxes[r0:r1]), "cls": torch.from_numpy(cls[r0:r1]),
               "conf": torch.from_numpy(conf[r0:r1]), "angle": torch.from_numpy(angle[r0:r1])}
        allr = sharding.allgather_records(rec)
        assert np.array_equal(allr["boxes"].numpy(), boxes) and np.array_equal(allr["cls"].numpy(), cls)
        assert np.array_equal(allr["conf"].numpy(), conf) and np.array_equal(allr["angle"].numpy(), angle)
        assert allr["boxes"].dtype == torch.float64 and allr["cls"].dtype == torch.int32
        # an empty rank and a too-small capacity (raises on every rank, no hang)
        empty = {k: v[:0] for k, v in rec.items()} if rank == 1 else rec
        part = sharding.allgather_records(empty)
        rest = np.concatenate([conf[cut[r]:cut[r + 1]] for r in range(world) if r != 1])
        assert part["conf"].shape[0] == len(rest) and np.array_equal(part["conf"].numpy(), rest)
        try:
            sharding.allgather_records(rec, capacity=3)
            raise AssertionError("capacity overflow not reported")
        except ValueError:
            pass

        def nms_fn(b, c, f):
            return torch.from_numpy(geom_c.nms(b.numpy(), c.numpy(), f.numpy(), 0.4)[1].astype(np.int64))

        kept = sharding.merge_sharded_by_class(allr["boxes"], allr["cls"], allr["conf"], 0.4, 4, nms_fn=nms_fn)
        want = geom_c.nms(boxes, cls, conf, 0.4)[1]
        ok = kept.tolist() == want.tolist()

        # the one-sync padded path: buffers longer than the valid count, capacity agreed beforehand
        def padded_nms(b, c, f):
            b, c, f = b.numpy(), c.numpy(), f.numpy()
            live = np.nonzero(c >= 0)[0]
            o, k = geom_c.nms(b[live], c[live], f[live], 0.4)
            order = np.argsort(-f.astype(np.float64), kind="stable")       # the real op returns the full stable conf-desc order
            keep = np.ones(len(f), np.uint8); keep[live] = 0; keep[live[k]] = 1
            return torch.from_numpy(order.astype(np.int64)), torch.from_numpy(keep)

        cap = n                                                             # >= every rank's count
        m = r1 - r0
        pad = lambda a: torch.from_numpy(np.concatenate([a[r0:r1], np.full((7,) + a.shape[1:], 3, dtype=a.dtype)]))
        rec2 = {"boxes": pad(boxes), "cls": pad(cls), "conf": pad(conf), "angle": pad(angle)}
        got = sharding.merge_bands_padded(rec2, torch.tensor([m]), cap, 0.4, 4, nms_fn=padded_nms)
        # padded positions r*cap + i map back to the concatenated list
        pos = got["index"].numpy()
        back = np.array([cut[int(p) // cap] + int(p) % cap for p in pos], dtype=np.int64)
        ok = ok and back.tolist() == want.tolist() and np.array_equal(got["boxes"].numpy(), boxes[want])
        ok = ok and np.array_equal(got["angle"].numpy(), angle[want])
        q.put((rank, ok, len(want)))
    finally:
        dist.destroy_process_group()


def test_two_rank_merge_equals_single_rank():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res) and res[0][2] > 100


def test_four_rank_merge_equals_single_rank():
    """The same host logic at world_size 4 (even split of the list; rank 1 empty in the second gather)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 4, port, q)) for r in range(4)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res) and res[0][2] > 100


def test_band_split_covers_all_rows():
    sys.path.insert(0, ROOT)
    from oriented_object_detection_b200 import sharding
    for total, world in ((52, 8), (168, 8), (26, 4), (3, 8), (1, 2)):
        spans = [sharding.band_rows(total, world, r) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1
    assert [sharding.band_rows(52, 8, r)[1] - sharding.band_rows(52, 8, r)[0] for r in range(8)] == [7, 7, 7, 7, 6, 6, 6, 6]
    assert sharding.band_pixel_rows(16384, 416, 100, 7, 14) == (7 * 316, 13 * 316 + 416)
    assert sharding.band_pixel_rows(16384, 416, 100, 46, 52) == (46 * 316, 16384)


def test_host_placement_helpers(tmp_path, monkeypatch):
    """cpulist parsing and the NUMA binding of a rank to its GPU's CPUs (fake sysfs; the call never raises)."""
    import os
    from oriented_object_detection_b200 import sharding
    assert sharding.parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    assert sharding.parse_cpulist("") == []
    # no CUDA device here: the lookup reports "unknown" and the binding leaves the process alone
    before = os.sched_getaffinity(0)
    assert sharding.gpu_local_cpus(0) is None or isinstance(sharding.gpu_local_cpus(0), list)
    r = sharding.bind_host_to_gpu(0, sysfs_root=str(tmp_path))
    assert "unchanged" in r and os.sched_getaffinity(0) == before
    # an explicit CPU list: intersected with the allowed set, applied, reported
    cpus = sorted(before)
    if len(cpus) >= 2:
        try:
            r = sharding.bind_host_to_gpu(0, cpus=cpus[:1] + [10 ** 6])
            assert r == {"cpus": 1, "first": cpus[0], "last": cpus[0]} and os.sched_getaffinity(0) == {cpus[0]}
        finally:
            os.sched_setaffinity(0, before)
    assert sharding.bind_host_to_gpu(0, cpus=[10 ** 6]) == {"unchanged": "local CPUs outside the allowed set"}
    assert sharding.bind_host_to_gpu(0, cpus=cpus)["unchanged"] == "already local"
