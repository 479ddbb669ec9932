#!/usr/bin/env python
"""Benchmark of the tiled-detection hot path (BASELINE.json metric: map Mpx/s).

    python bench.py --gpus 1 --steps K --warmup W                      (native arm, one GPU)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference --gpus N --steps K --warmup W     (CPU reference arm)

One step = one pass of the hot path over one synthetic map band per rank:
  4-channel [R,G,B,DT-Edge] tiling at 416/100 (BASELINE config 3: 8192^2 -> 676 tiles per GPU)
  -> tile->map remap + border filter + strike angle + per-tile rotated NMS of the band's synthetic
     per-tile detections (BASELINE config 2: ~100k OBBs, 15 classes, from the 8192^2 tiling)
  -> [N > 1: all_gather of the survivors over NCCL] -> class-wise exact greedy global NMS.
Weak scaling: the map is (8192*N) x 8192, each rank owns a band of tile rows (~8192 px rows).
`value` = map pixels of all ranks / max-over-ranks device time, inputs resident in HBM.
`e2e`   = the same with the band's pixels and detections copied from pinned host memory and the
          merged records copied back inside the timed region, every step; two steps are in flight
          (double-buffered device map), so a step's upload overlaps the previous step's build.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

MAP_SIDE = 16384                  # BASELINE config 5: 16384^2 maps
MAPS_PER_STEP = 8                 # maps batched per step: 64 maps = 8 steps
TILE, OVERLAP, MARGIN = 416, 100, 20
N_CLASSES = 15
OBJECTS_PER_BAND = 59000          # CPU sample: -> ~100k per-tile detections on the 8192^2 plan (1.7 copies per object)
OBJECTS_PER_MAP = 236000          # the same density on a 16384^2 map -> ~400k per-tile detections per map
IOU_MERGE = 0.4
MAX_DET_PER_TILE = int(os.environ.get("GM_BENCH_MAX_DET", "300"))   # the detector's max_det (Ultralytics default): the per-tile NMS may rely on it
METRIC = "map Mpx/s (tile+DT-Edge+merge)"
SAMPLE_TILES = 7                  # CPU arms: a 7x7-tile sub-map (2312^2 px) of the same workload


def _traffic(kernel: str):
    """DRAM bytes per launch of `kernel` from the newest committed ncu capture (profiles/rNN_traffic.json), or None."""
    import glob
    for p in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_traffic.json")), reverse=True):
        with open(p) as fh:
            v = json.load(fh).get("bytes_per_launch", {}).get(kernel)
        if v is not None:
            return v, os.path.basename(p)
    return None, None


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------- clocks sampler

class ClockSampler:
    """Samples SM clock and throttle reasons through NVML every ~10 ms while running."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.01)

    def start(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr is not None:
            self._thr.join(timeout=1.0)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ----------------------------------------------------------------------------- CPU arms (oracle = the checker, timed here as the baseline)
# Nothing on this path touches the CUDA library: the tile plan comes from oracle.geometry.tile_plan and the synthetic
# generators are loaded from synth.py by file path (importing the package would map libgeomap_b200.so).

_REF = {}


def _reference_build_multich():
    """The reference's OWN build_multich when its two scripts travelled with the push (oracle/_ref/, git-ignored, filled
    by `make -C oracle` from /root/reference in the build container) or /root/reference is present: the lifted function
    (oracle/lift_reference.py executes the reference source, nothing is restated).  Otherwise the OpenCV/numpy port."""
    if "fn" not in _REF:
        from oracle import lift_reference as LR
        try:
            import cv2
            cv2.ipp.setUseIPP(False)
            mod = LR.load_detect(4) if LR.reference_available() else None
        except Exception:                                   # noqa: BLE001
            mod = None
        if mod is not None:
            _REF["fn"], _REF["kind"] = (lambda crop: mod.build_multich(crop, 4)), "reference"
        else:
            from oracle import pixel_cv
            _REF["fn"], _REF["kind"] = (lambda crop: pixel_cv.build_multich(crop, 4)), "port"
    return _REF["fn"], _REF["kind"]


def _cpu_tile_job(crop):
    import cv2
    cv2.setNumThreads(1)
    fn, kind = _reference_build_multich()
    return fn(crop).shape[0], kind


def _load_synth():
    import importlib.util
    spec = importlib.util.spec_from_file_location("gm_synth_standalone", os.path.join(ROOT, "oriented_object_detection_b200", "synth.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class _HostPlan:
    """The fields synth.synthetic_tile_dets reads, from the oracle's tile plan."""

    def __init__(self, H, W, tile, overlap):
        import numpy as np
        from oracle import geometry as G
        t = G.tile_plan(H, W, tile, overlap)
        step = max(1, tile - overlap)
        self.H, self.W, self.tile_size, self.overlap, self.row_begin = H, W, tile, overlap, 0
        self.rows, self.cols = -(-H // step), -(-W // step)
        self.tiles = np.zeros(len(t), dtype=[("y0", "<i4"), ("x0", "<i4"), ("h", "<i4"), ("w", "<i4")])
        for k, name in enumerate(("y0", "x0", "h", "w")):
            self.tiles[name] = [v[k] for v in t]


def _cpu_sample_inputs(seed: int):
    """The top-left SAMPLE_TILES x SAMPLE_TILES tiles of one 16384^2 map of the workload: pixels + their detections."""
    import numpy as np
    synth = _load_synth()
    side = (SAMPLE_TILES - 1) * (TILE - OVERLAP) + TILE
    img = synth.synthetic_map(MAP_SIDE, MAP_SIDE, seed, "cpu", row0=0, rows=side, col0=0, cols=side).numpy()
    plan_full = _HostPlan(MAP_SIDE, MAP_SIDE, TILE, OVERLAP)
    local, cls, conf, tid = synth.synthetic_tile_dets(plan_full, OBJECTS_PER_MAP, N_CLASSES, seed=0, margin=MARGIN)
    r, c = tid // plan_full.cols, tid % plan_full.cols
    sel = (r < SAMPLE_TILES) & (c < SAMPLE_TILES)
    tiles = [(int(t["y0"]), int(t["x0"]), int(t["h"]), int(t["w"])) for t in plan_full.tiles
             if t["y0"] // (TILE - OVERLAP) < SAMPLE_TILES and t["x0"] // (TILE - OVERLAP) < SAMPLE_TILES]
    return img, tiles, (local[sel], cls[sel], conf[sel], tid[sel]), plan_full


def _cpu_step(pool, img, tiles, dets, plan_full):
    """Reference path on the sample: per-tile DT-Edge (all host cores), remap + filter + per-tile NMS,
    global NMS (C restatement of merge_detections - far faster than the reference's Python loop)."""
    import numpy as np
    from oracle import geom_c
    crops = [np.ascontiguousarray(img[y:y + h, x:x + w]) for (y, x, h, w) in tiles]
    kinds = {k for _, k in pool.map(_cpu_tile_job, crops, chunksize=1)}
    local, cls, conf, tid = dets
    gb, gc, gf = [], [], []
    for t in np.unique(tid):
        s = np.nonzero(tid == t)[0]
        tl = plan_full.tiles[t]
        b = local[s].astype(np.float64)
        b[:, 0::2] += float(tl["x0"]); b[:, 1::2] += float(tl["y0"])
        cx = b[:, 0::2].sum(1) / 4.0 - float(tl["x0"]); cy = b[:, 1::2].sum(1) / 4.0 - float(tl["y0"])
        ok = (cx >= MARGIN) & (cx <= tl["w"] - MARGIN) & (cy >= MARGIN) & (cy <= tl["h"] - MARGIN)
        b, c, f = b[ok], cls[s][ok], conf[s][ok]
        _, kept = geom_c.nms(b, c, f, IOU_MERGE)
        gb.append(b[kept]); gc.append(c[kept]); gf.append(f[kept])
    if gb:
        geom_c.nms(np.concatenate(gb), np.concatenate(gc), np.concatenate(gf), IOU_MERGE)
    return kinds


def cpu_baseline(steps: int, warmup: int, seed: int = 1000):
    """Times the CPU path on the bounded sample; returns (Mpx/s, cores, sample description, ms/step, kind)."""
    import concurrent.futures as cf
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    img, tiles, dets, plan_full = _cpu_sample_inputs(seed)
    side = img.shape[0]
    with cf.ProcessPoolExecutor(max_workers=cores, mp_context=mp.get_context("spawn")) as pool:
        list(pool.map(_cpu_tile_job, [img[:64, :64].copy()] * cores))           # start the workers
        for _ in range(warmup):
            _cpu_step(pool, img, tiles, dets, plan_full)
        t0 = time.perf_counter()
        for _ in range(steps):
            kinds = _cpu_step(pool, img, tiles, dets, plan_full)
        dt = (time.perf_counter() - t0) / max(steps, 1)
    kind = "reference" if kinds == {"reference"} else "port"
    what = ("the reference's own build_multich (Detect_OBB.py:87-133, lifted from oracle/_ref)" if kind == "reference"
            else "OpenCV/numpy port of build_multich (the reference scripts did not travel)")
    desc = (f"{SAMPLE_TILES}x{SAMPLE_TILES} tiles ({side}x{side} px) of one {MAP_SIDE}^2 map of the workload + their {len(dets[2])} "
            f"per-tile detections, {steps} steps; {what} over {cores} processes; merge_detections = C restatement "
            f"(shapely absent; the reference's Python loop is orders of magnitude slower)")
    return side * side / 1e6 / dt, cores, desc, dt * 1e3, kind


# ----------------------------------------------------------------------------- native arm

def _rank_workload(args, world, rank, dev):
    """BASELINE config 5: `maps` synthetic 16384^2 maps per step, every map sharded over the ranks by a balanced
    row-major TILE RANGE (a row band whose first / last tile row may be partial).  A rank stacks its band of every map of
    the batch in one buffer, so one launch sequence serves the whole batch."""
    import numpy as np
    import torch
    from oriented_object_detection_b200 import ops, sharding, synth
    H = W = args.map_side
    M = args.maps
    full = ops.make_plan(H, W, TILE, OVERLAP)
    cols = full.cols
    t0, t1 = sharding.tile_range(full.n, world, rank)
    nt = t1 - t0
    r_first, r_last = t0 // cols, (t1 - 1) // cols
    y0, y1 = sharding.band_pixel_rows(H, TILE, OVERLAP, r_first, r_last + 1)
    band_h = y1 - y0
    mine = full.tiles[t0:t1]
    rep = lambda a: np.tile(np.asarray(a), M)
    plan_geo = ops.plan_from_arrays(H, W, rep(mine["y0"]), rep(mine["x0"]), rep(mine["h"]), rep(mine["w"]), device=dev,
                                    tile_size=TILE, overlap=OVERLAP)                       # map coordinates (remap, border filter)
    y_stack = np.concatenate([mine["y0"] - y0 + m * band_h for m in range(M)])
    plan_px = ops.plan_from_arrays(M * band_h, W, y_stack, rep(mine["x0"]), rep(mine["h"]), rep(mine["w"]), device=dev,
                                   tile_size=TILE, overlap=OVERLAP)                        # rows of the stacked band buffer
    map_stack = torch.empty((M * band_h, W, 3), dtype=torch.uint8, device=dev)
    for m in range(M):
        map_stack[m * band_h:(m + 1) * band_h] = synth.synthetic_map(H, W, 1000 + m, dev, row0=y0, rows=band_h)
    # the detections of a map are generated for the WHOLE map from its seed and then cut to the rank's tiles, so the set -
    # and therefore the merged result - does not depend on how many ranks share the map (config.merged_checksum)
    parts = []
    for m in range(M):
        local, cls, conf, tid = synth.synthetic_tile_dets(full, args.objects_per_map, N_CLASSES, seed=m, margin=MARGIN)
        sel = (tid >= t0) & (tid < t1)
        parts.append((local[sel], cls[sel], conf[sel], (tid[sel].astype(np.int64) - t0 + m * nt).astype(np.int32)))
    dets = tuple(np.ascontiguousarray(np.concatenate([p[k] for p in parts])) for k in range(4))
    rects = sharding.foreign_center_rects(H, W, TILE, OVERLAP, t0, t1, MARGIN)
    scope = sharding.seam_scope(H, W, TILE, OVERLAP, MARGIN, world, rank, reach=1)
    return dict(scope=scope, H=H, W=W, M=M, full=full, t0=t0, t1=t1, nt=nt, y0=y0, y1=y1, band_h=band_h, plan_geo=plan_geo, plan_px=plan_px,
                map_stack=map_stack, dets=dets, rects=rects)


def native(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import __graft_entry__ as entry

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # NCCL's log (version banner included) defaults to stdout
        torch.cuda.set_device(local_rank)
        import datetime
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank),
                                timeout=datetime.timedelta(seconds=300))
    else:
        torch.cuda.set_device(0)
    dev = torch.device("cuda", torch.cuda.current_device())
    if rank == 0:
        entry.build()
    if world > 1:
        dist.barrier()
    from oriented_object_detection_b200 import _lib, ops, sharding

    placement = {"unchanged": "single GPU"}
    if world > 1 and os.environ.get("GM_BIND_NUMA", "1") != "0":
        placement = sharding.bind_host_to_gpu(local_rank)
        print(f"[bench] rank {rank}: host placement {placement}", file=sys.stderr, flush=True)

    wl = _rank_workload(args, world, rank, dev)
    H, W, M, nt = wl["H"], wl["W"], wl["M"], wl["nt"]
    plan_geo, plan_px, map_stack, rects = wl["plan_geo"], wl["plan_px"], wl["map_stack"], wl["rects"]
    local, cls, conf, tid = wl["dets"]
    n_dets = len(conf)
    h_map = map_stack.cpu().pin_memory()
    h_det = [torch.from_numpy(a).pin_memory() for a in (local, cls, conf, tid)]
    d_det = [t.to(dev) for t in h_det]
    out4 = torch.empty(4 * plan_px.total_px, dtype=torch.uint8, device=dev)
    KEY_CLASSES = M * N_CLASSES                      # class key of the batched merge: map * N_CLASSES + class

    def tile_stage():
        """remap / border filter / strike angle / per-tile NMS of the whole batch, then the merge's class key."""
        pp = ops.tile_postprocess(d_det[0], d_det[1], d_det[2], d_det[3], plan_geo, MARGIN, 1, IOU_MERGE,
                                  max_class=N_CLASSES - 1, sync=False, max_per_tile=MAX_DET_PER_TILE)
        src = pp["src"].clamp_(0, max(n_dets - 1, 0)).to(torch.int64)       # rows beyond the count are not data
        map_of = torch.div(d_det[3][src], nt, rounding_mode="floor").to(torch.int32)
        rec = {"boxes": pp["boxes"], "cls": pp["cls"] + map_of * N_CLASSES, "conf": pp["conf"], "angle": pp["angle"]}
        return rec, pp["count"]

    # bounds of the seam exchange, agreed once from this batch (two small all_reduce at set-up; verified by every step)
    rec0, count0 = tile_stage()
    extent_bound, seam_cap = sharding.agree_seam_bounds(rec0, count0, IOU_MERGE, KEY_CLASSES - 1, rects)
    survivors_rank = int(count0.item())
    del rec0, count0
    cap_out = n_dets + 16
    h_out = {"boxes": torch.empty((cap_out, 8), dtype=torch.float64).pin_memory(),
             "cls": torch.empty(cap_out, dtype=torch.int32).pin_memory(),
             "conf": torch.empty(cap_out, dtype=torch.float32).pin_memory(),
             "angle": torch.empty(cap_out, dtype=torch.float64).pin_memory()}
    result = {}

    # small latency-bound kernels: schedule their CTAs first (GM_DET_PRIORITY=0 puts them on a normal-priority stream)
    det_stream = torch.cuda.Stream(device=dev, priority=int(os.environ.get("GM_DET_PRIORITY", "-1")))
    det_stream.wait_stream(torch.cuda.current_stream())      # inputs above were produced on the current stream
    # diagnostic (GM_BUILD_PRIORITY): the build on its own stream with that priority instead of the current stream
    build_stream = (torch.cuda.Stream(device=dev, priority=int(os.environ["GM_BUILD_PRIORITY"]))
                    if os.environ.get("GM_BUILD_PRIORITY") else None)

    def merge_device():
        """per-tile stage -> local NMS of the band with the seam deferred -> ONE all_gather of the seam records -> seam
        verdicts -> this rank's kept records: fixed shapes, no host read (sharding.merge_bands_seam_device)."""
        rec, count = tile_stage()
        return sharding.merge_bands_seam_device(rec, count, seam_cap, IOU_MERGE, KEY_CLASSES - 1, rects, extent_bound,
                                                scope=wl["scope"])

    use_graph = args.graph in ("on", "auto")
    merge_call = sharding.CapturedCall(merge_device, stream=det_stream) if use_graph else merge_device
    graph_state = ("captured" if merge_call.captured else f"eager (capture failed: {merge_call.error})") if use_graph else "eager"

    def merge_enqueue(from_host: bool):
        if from_host:
            for d, h in zip(d_det, h_det):
                d.copy_(h, non_blocking=True)
        return merge_call()

    def merge_finish(dev_out, from_host: bool):
        """the one host read (4 integers) [-> this rank's merged records to pinned host memory]."""
        rec = sharding.merge_bands_seam_finish(dev_out)
        m = int(rec["conf"].shape[0])
        result["kept"], result["survivors_all"], result["seam_all"] = m, rec["n_valid"], rec["n_seam"]
        result["rec"] = rec
        result["fallbacks"] = result.get("fallbacks", 0) + rec["chain_fallbacks"]
        if from_host:
            for k in h_out:
                h_out[k][:m].copy_(rec[k], non_blocking=True)
        return m

    def merge_path(from_host: bool):
        return merge_finish(merge_enqueue(from_host), from_host)

    n_chunks = args.chunks if args.chunks > 0 else max(13, plan_px.n // 169)
    e2e_streams = [torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)]
    map_bufs = [map_stack, torch.empty_like(map_stack)]
    out_bufs = [out4, torch.empty_like(out4)]
    in_flight = [None, None]

    def step(from_host: bool, i: int = 0):
        # The pixel path and the detection path of one step have no data dependence (in the reference the CNN sits
        # between them), so the small, latency-bound merge runs on its own stream beside the build.
        if not from_host:
            main = torch.cuda.current_stream()
            if os.environ.get("GM_BENCH_ONLY", "") != "det":        # diagnostic: one of the two paths alone
                if build_stream is not None:
                    build_stream.wait_stream(main)
                    with torch.cuda.stream(build_stream):
                        ops.dtedge_build(map_stack, plan_px, out=out4)
                    main.wait_stream(build_stream)
                else:
                    ops.dtedge_build(map_stack, plan_px, out=out4)
            if os.environ.get("GM_BENCH_ONLY", "") == "build":
                return 0
            with torch.cuda.stream(det_stream):
                kept = merge_path(False)
            main.wait_stream(det_stream)
            return kept
        slot = i % 2
        if in_flight[slot] is not None:
            in_flight[slot].synchronize()                    # step i - 2 is complete: its buffers are free again
        with torch.cuda.stream(e2e_streams[slot]):
            main = torch.cuda.current_stream()
            with torch.cuda.stream(det_stream):
                pending = merge_enqueue(True)
            ops.build_tiles_from_host(h_map, plan_px, 4, out=out_bufs[slot], map_dev=map_bufs[slot], n_chunks=n_chunks)
            with torch.cuda.stream(det_stream):
                kept = merge_finish(pending, True)           # one host read per step
            main.wait_stream(det_stream)
            in_flight[slot] = torch.cuda.Event()
            in_flight[slot].record(main)
        return kept

    def timed(from_host: bool, steps: int, warmup: int):
        cur = torch.cuda.current_stream()

        def fork():
            for st in e2e_streams:
                st.wait_stream(cur)

        def join():
            for st in e2e_streams:
                cur.wait_stream(st)

        fork()
        for i in range(warmup):
            step(from_host, i)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = _lib.lib.gm_launch_count()
        e0.record()
        fork()
        for i in range(steps):
            step(from_host, i)
        join()
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        replayed = merge_call.launches if (use_graph and merge_call.captured) else 0    # kernels inside the graph, per replay
        return float(ms.item()) / steps, (_lib.lib.gm_launch_count() - l0) // max(steps, 1) + replayed

    sampler = ClockSampler(torch.cuda.current_device())
    sampler.start()
    ms_dev, launches = timed(False, args.steps, max(args.warmup, 3))
    ms_e2e, _ = timed(True, args.steps, max(args.warmup, 3))
    clocks = sampler.stop()

    # ---- per-kernel breakdown of the DT-Edge build (CUDA events between its kernels) and of the merge
    stage = {k: 0.0 for k in ops.DTEDGE_STAGES}
    reps = 3
    for _ in range(reps):
        _, ms = ops.dtedge_build_timed(map_stack, plan_px, out=out4)
        for k in stage:
            stage[k] += ms[k] / reps
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def best_ms(fn, n=reps):
        best = 1e9
        for _ in range(n):
            torch.cuda.synchronize()
            e0.record()
            r = fn()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return best, r

    ms_tilepp, _ = best_ms(lambda: tile_stage())
    GATHER_REPS = 4
    out3 = torch.empty(3 * plan_px.total_px, dtype=torch.uint8, device=dev)
    ms_gather, _ = best_ms(lambda: [ops.tile_gather(map_stack, plan_px, out=out3) for _ in range(GATHER_REPS)])
    ms_gather /= GATHER_REPS
    del out3
    # the whole detection path as the step runs it (graph replay + the host read), back to back, nothing beside it
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0w = time.perf_counter()
    for _ in range(reps):
        merge_path(False)
    torch.cuda.synchronize()
    ms_merge_wall = (time.perf_counter() - t0w) * 1e3 / reps
    ops.threshold_adjacent_stats(reset=True)
    merge_path(False)
    torch.cuda.synchronize()
    adjacent = ops.threshold_adjacent_stats()
    # H2D ceiling of this box with every rank copying at once: the same pinned band, the same chunking, nothing else
    copy_stream = torch.cuda.Stream(device=dev)
    rows_per = max(1, h_map.shape[0] // n_chunks)
    def bare_upload():
        with torch.cuda.stream(copy_stream):
            for ya in range(0, h_map.shape[0], rows_per):
                map_bufs[1][ya:ya + rows_per].copy_(h_map[ya:ya + rows_per], non_blocking=True)
        copy_stream.synchronize()
    bare_upload()
    if world > 1:
        dist.barrier()
    tb = time.perf_counter()
    for _ in range(3):
        bare_upload()
    h2d_s = torch.tensor([(time.perf_counter() - tb) / 3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(h2d_s, op=dist.ReduceOp.MAX)
    h2d_probe_gbs = world * h_map.numel() / float(h2d_s.item()) / 1e9

    hbm_peak, peak_src = _peaks()
    band_px = int(map_stack.shape[0]) * W
    top = max(stage, key=stage.get)
    # algorithmic bytes of the dominant DT-Edge kernel: grad reads the band once (3 B/px) and writes the
    # gradient energy once (4 B per tile pixel); the whole build: 3 B/map px + 4 B/tile px.
    alg = {"grad": 3 * band_px + 4 * plan_px.total_px, "select_grad": 4 * plan_px.total_px,
           "edge_open": 4 * plan_px.total_px + plan_px.total_px // 8, "chamfer": plan_px.total_px // 8 + 4 * plan_px.total_px,
           "select_dist": 4 * plan_px.total_px, "tail": 3 * band_px + 8 * plan_px.total_px + 4 * plan_px.total_px}
    achieved = alg[top] / (stage[top] * 1e-3) / 1e9
    build_ms = sum(stage.values())
    traffic, traffic_file = _traffic("k_" + top)
    roofline = {"bound": "hbm", "kernel": "k_" + top, "achieved": round(achieved, 1), "peak": hbm_peak, "unit": "GB/s",
                "frac": round(achieved / hbm_peak, 4),
                "traffic": (int(traffic * plan_px.total_px / 114318864) if traffic else None), "peak_source": peak_src,
                "traffic_note": f"ncu dram bytes of this kernel per launch on the 8192^2 plan (profiles/{traffic_file}), scaled by tile pixels",
                "algorithmic_bytes_per_launch": alg[top], "kernel_ms": round(stage[top], 4),
                "dtedge_build_ms": round(build_ms, 4),
                "dtedge_build_frac_of_hbm": round((3 * band_px + 4 * plan_px.total_px) / (build_ms * 1e-3) / 1e9 / hbm_peak, 4),
                "stages_ms": {k: round(v, 4) for k, v in stage.items()},
                "tile_stage_ms": round(ms_tilepp, 4), "merge_path_wall_ms": round(ms_merge_wall, 4),
                "tile_gather3_ms": round(ms_gather, 4),
                "tile_gather3_frac_of_hbm": round((3 * band_px + 3 * plan_px.total_px) / (ms_gather * 1e-3) / 1e9 / hbm_peak, 4),
                "note": "the build is issue bound (~340 thread-instructions per tile pixel over six kernels; k_grad 81 % issue active) - the HBM fraction is an upper-bound view; kernels that are memory bound in this step: chamfer, edge_open, tail (72-82 % of the measured peak for the bytes they move)"}

    # ---- rotated IoU throughput (dense matrix, no early-out), EVERY rank; fraction of the measured FFMA peak and of nominal
    iou = None
    if not args.no_iou:
        nb = 8192
        sel = slice(0, nb)
        bxh = local[sel].astype(np.float64)
        bxh[:, 0::2] += plan_geo.tiles["x0"][tid[sel]][:, None]
        bxh[:, 1::2] += plan_geo.tiles["y0"][tid[sel]][:, None]
        bx = torch.from_numpy(bxh).to(dev)
        rs = torch.empty(nb, dtype=torch.float64, device=dev)
        for _ in range(2):
            ops.rotated_iou_matrix_sum(bx, bx, out=rs)
        e0.record()
        for _ in range(5):
            ops.rotated_iou_matrix_sum(bx, bx, out=rs)
        e1.record(); torch.cuda.synchronize()
        ms_iou = e0.elapsed_time(e1) / 5
        ffma = ops.ffma_peak(8192)
        gp = nb * nb / (ms_iou * 1e-3) / 1e9
        per_rank = torch.tensor([gp], dtype=torch.float64, device=dev)
        allr = [torch.zeros_like(per_rank) for _ in range(world)]
        if world > 1:
            dist.all_gather(allr, per_rank)
        else:
            allr = [per_rank]
        vals = [float(v.item()) for v in allr]
        iou = {"gpairs_per_s": round(sum(vals), 2), "per_rank_min": round(min(vals), 2), "per_rank_max": round(max(vals), 2),
               "pairs_per_rank": nb * nb, "ms": round(ms_iou, 4), "flop_per_pair": 210,
               "achieved_tflops_per_gpu": round(gp * 210 / 1e3, 2), "ffma_peak_tflops_measured": round(ffma, 1),
               "frac_of_measured_ffma": round(gp * 210 / 1e3 / ffma, 4), "nominal_fp32_tflops": 74.4,
               "frac_of_nominal_fp32": round(gp * 210 / 1e3 / 74.4, 4),
               "workload": "dense 8192 x 8192 matrix per rank over the first 8192 per-tile OBBs of the rank's tiles (fp32 tile-local "
                           "corners + tile offset), no early-out, checksum per column"}
        if rank == 0:
            # error of the fp32 kernel against the reference's float64 arithmetic (gm_rotated_iou_pairs_f64), per IoU
            # bucket, over every overlapping pair of a 2048 x 2048 block of the same boxes (untimed)
            sub = bx[:2048]
            mat = ops.rotated_iou_matrix(sub, sub)
            ii, jj = torch.nonzero(mat > 0, as_tuple=True)
            ref64 = ops.rotated_iou_pairs_f64(sub, sub, ii, jj)
            got = mat[ii, jj].to(torch.float64)
            err = (got - ref64).abs()
            buckets = {}
            for lo, hi in ((0.0, 0.01), (0.01, 0.05), (0.05, 0.3), (0.3, 0.5), (0.5, 1.01)):
                msk = (ref64 >= lo) & (ref64 < hi)
                if int(msk.sum().item()):
                    buckets[f"[{lo},{min(hi, 1.0)}{']' if hi > 1 else ')'}"] = {
                        "pairs": int(msk.sum().item()), "max_abs": float(f"{err[msk].max().item():.3g}"),
                        # a relative error means nothing where the float64 value itself may be 0 (touching boxes)
                        "max_rel": (float(f"{(err[msk] / ref64[msk]).max().item():.3g}") if lo > 0 else None)}
            iou["error_vs_float64"] = {"pairs": int(ii.numel()), "max_abs": float(f"{err.max().item():.3g}"), "by_iou": buckets,
                                       "note": "fp32 dense kernel against the float64 pair kernel on all overlapping pairs of a "
                                               "2048 x 2048 block; decisions within 1e-4 of a threshold are redone in float64"}

    extras = {}
    if world == 1 and not args.no_extras:
        extras = _extras(dev)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, cores, desc, _, kind = cpu_baseline(steps=2, warmup=1)
        cpu = {"value": round(v, 3), "unit": "Mpx/s", "cores": cores, "kind": kind, "sample": desc}

    tot = torch.tensor([n_dets, survivors_rank, result.get("kept", 0)], dtype=torch.int64, device=dev)
    # order-independent fingerprint of the merged records of ALL ranks: the same for every number of ranks
    rec = result.get("rec")
    fp = torch.zeros(2, dtype=torch.float64, device=dev)
    if rec is not None and rec["conf"].numel():
        fp[0] = rec["conf"].to(torch.float64).sum()
        fp[1] = (rec["boxes"].sum(dim=1) * (rec["cls"].to(torch.float64) + 1.0)).sum()
    if world > 1:
        dist.all_reduce(tot)
        dist.all_reduce(fp)
    n_dets_all, survivors_all, kept_all = (int(v) for v in tot.tolist())
    merged_checksum = [kept_all, round(float(fp[0].item()), 3), round(float(fp[1].item()), 1)]
    if rank == 0:
        total_px = M * H * W
        line = {
            "metric": METRIC, "value": round(total_px / 1e6 / (ms_dev * 1e-3), 1), "unit": "Mpx/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(ms_dev, 4),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u8/int32 (pixels), f32 pair-local + f64 decisions (geometry)", "data": "synthetic",
            "config": {"workload": f"c5: {M} synthetic {H}x{W} BGR maps per step (64 maps = {64 // M if 64 % M == 0 else 64 / M} steps), every "
                                   f"map sharded over {world} rank(s) by a balanced tile range (row band): 4-ch [R,G,B,DT-Edge] "
                                   f"tiling {TILE}/{OVERLAP} ({wl['full'].n} tiles per map) + remap/border filter/per-tile NMS of "
                                   f"{n_dets_all} synthetic per-tile OBBs ({N_CLASSES} classes) + exact greedy global NMS"
                                   f"{' with ONE all_gather of the seam-band detections' if world > 1 else ''}",
                       "map": [H, W], "maps_per_step": M, "tile": TILE, "overlap": OVERLAP, "tiles_per_rank": plan_px.n,
                       "tile_px_per_rank": plan_px.total_px, "detections": n_dets_all,
                       "survivors_after_tile_nms": survivors_all, "merged": kept_all, "merged_checksum": merged_checksum,
                       "seam_rows_exchanged": result.get("seam_all"), "seam_capacity_per_rank": seam_cap,
                       "seam_blocks_resolved_per_rank": int(wl["scope"]["blocks"][1] - wl["scope"]["blocks"][0]),
                       "seam_chain_fallbacks_rank0": result.get("fallbacks", 0),
                       "seam_extent_bound_px": round(extent_bound, 2),
                       "parallelism": f"tile-range row band x{world}, 1 collective per step" if world > 1 else "single GPU",
                       "host_placement": placement, "detection_path": graph_state,
                       "l2": "working set per step (map bands + >10 GB of stage buffers) exceeds the 126 MB L2; no flush needed"},
            "e2e": {"value": round(total_px / 1e6 / (ms_e2e * 1e-3), 1), "unit": "Mpx/s", "ms_per_step": round(ms_e2e, 4),
                    "h2d_bytes_per_step": int(world * (h_map.numel() + sum(t.numel() * t.element_size() for t in h_det))),
                    "d2h_bytes_per_step": int(kept_all * (64 + 4 + 4 + 8)),
                    "h2d_probe_gbs": round(h2d_probe_gbs, 1),
                    "h2d_achieved_gbs": round(world * h_map.numel() / (ms_e2e * 1e-3) / 1e9, 1),
                    "note": "h2d_probe = all ranks uploading their pinned band at once with nothing else running (same chunking); "
                            "per-rank bytes x world (rank 0's band size)"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "iou": iou,
            "threshold_adjacent_pairs": dict(adjacent, iou_threshold=IOU_MERGE, note="per pass of the detection path on rank 0; "
                                             "decided on the float64 IoU, like the reference"),
        }
        line.update(extras)
        _emit(line)
    if world > 1:
        # A process group whose collectives sit inside a live CUDA graph can block in its destructor; the JSON line
        # is out, so tear down with a deadline and leave regardless.
        def _shutdown():
            try:
                torch.cuda.synchronize()
                if use_graph and getattr(merge_call, "graph", None) is not None:
                    merge_call.graph.reset()
                dist.barrier()
                dist.destroy_process_group()
            except Exception:                       # noqa: BLE001
                pass
        th = threading.Thread(target=_shutdown, daemon=True)
        th.start()
        th.join(20.0)
        sys.stdout.flush(); sys.stderr.flush()
        os._exit(0)


def _extras(dev):
    """Single-GPU legs outside the timed step (BASELINE configs 4 and 1, the Otsu binarisation): each reports its own time;
    a failing leg reports its error instead of taking the line down."""
    import numpy as np
    import torch
    from oriented_object_detection_b200 import _lib, ops, synth
    out = {}
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]

    def ms_of(fn, n=3):
        best, r = 1e9, None
        for _ in range(n):
            torch.cuda.synchronize(); ev[0].record(); r = fn(); ev[1].record(); torch.cuda.synchronize()
            best = min(best, ev[0].elapsed_time(ev[1]))
        return best, r
    try:        # c4: dual-scale 128/30 + 416/100 late fusion on a 16384^2 map, ~1 M candidate boxes
        H = W = 16384
        sets = []
        for (ts, ov, mg) in ((128, 30, 10), (416, 100, 20)):
            plan = ops.make_plan(H, W, ts, ov, device=dev)
            local, cls, conf, tid = synth.synthetic_tile_dets(plan, 400000, N_CLASSES, seed=1, margin=mg)
            d = [torch.from_numpy(a).to(dev) for a in (local, cls, conf, tid)]
            ms_t, pp = ms_of(lambda: ops.tile_postprocess(d[0], d[1], d[2], d[3], plan, mg, 1, IOU_MERGE, max_class=N_CLASSES - 1,
                                                          max_per_tile=MAX_DET_PER_TILE), 2)
            sets.append((pp, ms_t, len(conf)))
        boxes = torch.cat([s[0]["boxes"] for s in sets]); cls = torch.cat([s[0]["cls"] for s in sets])
        conf = torch.cat([s[0]["conf"] for s in sets])
        sid = torch.cat([torch.full((s[0]["conf"].shape[0],), k, dtype=torch.int32, device=dev) for k, s in enumerate(sets)])
        ms_f, fused = ms_of(lambda: ops.fuse_scales(boxes, cls, conf, sid, 2, max_class=N_CLASSES - 1))
        fi = fused.to(torch.int64)
        fb, fc, ff = boxes[fi].contiguous(), cls[fi].contiguous(), conf[fi].contiguous()
        ms_n, (_, _, kept) = ms_of(lambda: ops.nms_global(fb, fc, ff, IOU_MERGE, max_class=N_CLASSES - 1))
        n = int(conf.shape[0])
        fusion = {"workload": "c4: 16384^2 map, 128/30 (28,224 tiles) + 416/100 (2,704 tiles) per-tile survivors -> cross_scale_consensus_filter "
                              "-> merge_detections", "candidate_boxes": n, "raw_detections": [s[2] for s in sets],
                  "tile_stage_ms": [round(s[1], 3) for s in sets], "fusion_ms": round(ms_f, 3), "fused": int(fused.numel()),
                  "global_nms_ms": round(ms_n, 3), "merged": int(kept.numel()),
                  "boxes_per_s": round(n / ((ms_f + ms_n) * 1e-3), 1)}
        try:    # CPU side: the C restatement of the two reference loops on a bounded window of the same set
            from oracle import geom_c
            hb, hc, hf, hs = boxes.cpu().numpy(), cls.cpu().numpy(), conf.cpu().numpy(), sid.cpu().numpy()
            win = (hb[:, 0] < 4096) & (hb[:, 1] < 4096)
            t0 = time.perf_counter()
            k = geom_c.fuse(np.ascontiguousarray(hb[win]), np.ascontiguousarray(hc[win]), np.ascontiguousarray(hf[win]),
                            np.ascontiguousarray(hs[win]), 2, grid=True)
            geom_c.nms(np.ascontiguousarray(hb[win][k]), np.ascontiguousarray(hc[win][k]), np.ascontiguousarray(hf[win][k]), IOU_MERGE, grid=True)
            dt = time.perf_counter() - t0
            fusion["cpu_sample"] = {"boxes": int(win.sum()), "seconds": round(dt, 3), "boxes_per_s": round(int(win.sum()) / dt, 1),
                                    "kind": "port (C restatement with a uniform-grid candidate index, 1 core; the reference's "
                                            "Python + shapely double loop is O(n^2))",
                                    "extrapolation": "window = 1/16 of the map area; the gridded loops are ~linear in boxes"}
        except Exception as e:          # noqa: BLE001
            fusion["cpu_sample"] = {"error": repr(e)}
        out["fusion"] = fusion
    except Exception as e:              # noqa: BLE001
        out["fusion"] = {"error": repr(e)}
    try:        # Otsu binarisation (DT_BIN_METHOD = "otsu", Detect_OBB.py:109-111): k_otsu_grad takes the select_grad slot
        plan = ops.make_plan(8192, 8192, TILE, OVERLAP, device=dev)
        m8 = synth.synthetic_map(8192, 8192, 1000, dev)
        o8 = torch.empty(4 * plan.total_px, dtype=torch.uint8, device=dev)
        p = _lib.make_params(flags=_lib.bin_method_flags("otsu"))
        acc = {}
        for _ in range(3):
            _, ms = ops.dtedge_build_timed(m8, plan, params=p, out=o8)
            for k, v in ms.items():
                acc[k] = min(acc.get(k, 1e9), v)
        out["otsu"] = {"workload": "c3 with DT_BIN_METHOD='otsu' (8192^2, 676 tiles)", "k_otsu_grad_ms": round(acc["select_grad"], 4),
                       "build_ms": round(sum(acc.values()), 4)}
        # the small-tile plan of config 4 (128/30: 7,056 tiles of an 8192^2 map - one-CTA-per-tile stages in a different regime)
        plan128 = ops.make_plan(8192, 8192, 128, 30, device=dev)
        o128 = torch.empty(4 * plan128.total_px, dtype=torch.uint8, device=dev)
        acc = {}
        for _ in range(3):
            _, ms = ops.dtedge_build_timed(m8, plan128, out=o128)
            for k, v in ms.items():
                acc[k] = min(acc.get(k, 1e9), v)
        hbm_peak, _ = _peaks()
        alg128 = 3 * 8192 * 8192 + 4 * plan128.total_px
        out["dtedge_128"] = {"workload": "4-ch DT-Edge tiling of an 8192^2 map at 128/30 (7,056 tiles, the small scale of config 4)",
                             "stages_ms": {k: round(v, 4) for k, v in acc.items()}, "build_ms": round(sum(acc.values()), 4),
                             "frac_of_hbm": round(alg128 / (sum(acc.values()) * 1e-3) / 1e9 / hbm_peak, 4)}
        del m8, o8, o128
    except Exception as e:              # noqa: BLE001
        out["otsu"] = {"error": repr(e)}
    try:        # the per-tile protocol of the reference: build_multich(crop, 4) on ONE host crop per call (Detect_OBB.py:76-78, :87)
        from oriented_object_detection_b200 import detect
        crop = synth.synthetic_map_numpy(416, 416, seed=5)
        for _ in range(3):
            detect.build_multich(crop, 4)
        t0 = time.perf_counter()
        for _ in range(20):
            r4 = detect.build_multich(crop, 4)
        per_call = (time.perf_counter() - t0) / 20
        out["build_multich_host"] = {"workload": "detect.build_multich(uint8[416,416,3] host crop, 4): upload, six kernels, download, synchronous "
                                                 "(gm_build_multich_host), 20 calls",
                                     "ms_per_crop": round(per_call * 1e3, 3), "mpx_per_s": round(416 * 416 / per_call / 1e6, 1),
                                     "checksum": int(r4[..., 3].astype(np.int64).sum())}
    except Exception as e:              # noqa: BLE001
        out["build_multich_host"] = {"error": repr(e)}
    try:        # c1: one 807 x 895 map (the size of Input/Test1.png) through the drop-in entry points, 3-ch 416/100, random-init YOLO11n-OBB
        import tempfile
        import cv2
        from oriented_object_detection_b200 import detect
        from oriented_object_detection_b200.predictor import TilePredictor
        from oriented_object_detection_b200.yolo11_obb import random_init_yolo11_obb
        with tempfile.TemporaryDirectory() as td:
            path = os.path.join(td, "Test1.png")
            cv2.imwrite(path, synth.synthetic_map_numpy(807, 895, seed=1))
            saved = (detect.tile_sizes, detect.overlaps, detect.models, detect.channels, detect.output_tag)
            try:
                torch.manual_seed(0)
                detect.tile_sizes, detect.overlaps, detect.channels, detect.output_tag = [416], [100], 3, "_OFFLINE-RANDOM-INIT"
                detect.models = [TilePredictor(random_init_yolo11_obb("n", len(detect.CLASS_NAMES), 3, 416, seed=0), 416)]
                import contextlib, io
                times = []
                for _ in range(3):
                    torch.cuda.synchronize(); t0 = time.perf_counter()
                    with contextlib.redirect_stdout(io.StringIO()):
                        detect.process_image(path, td)
                    torch.cuda.synchronize(); times.append((time.perf_counter() - t0) * 1e3)
                out["c1"] = {"workload": "c1: process_image on a synthetic 807x895 map (Test1.png's size), single scale 416/100, 3 channels, "
                                         "random-init YOLO11n-OBB (PyTorch forward), JPG + XLSX written", "ms_first": round(times[0], 1),
                             "ms_best": round(min(times), 1), "detections": len(detect.all_dets_per_image.get(path, []))}
            finally:
                detect.tile_sizes, detect.overlaps, detect.models, detect.channels, detect.output_tag = saved
    except Exception as e:              # noqa: BLE001
        out["c1"] = {"error": repr(e)}
    return out


# ----------------------------------------------------------------------------- reference arm

def reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import subprocess
    mk = os.path.join(ROOT, "oracle", "Makefile")
    if os.path.isfile(mk):                  # the oracle's C restatement (+ oracle/_ref when /root/reference is here); no CUDA library
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
    steps, warmup = max(1, min(args.steps, 200)), max(1, min(args.warmup, 5))     # EXACTLY K steps (a step = 0.2-0.4 s of host work)
    v, cores, desc, ms, kind = cpu_baseline(steps=steps, warmup=warmup)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    line = {"impl": "reference", "metric": METRIC, "value": round(v, 3), "unit": "Mpx/s", "n_gpus": world,
            "steps": steps, "warmup": warmup, "ms_per_step": round(ms, 2), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u8/f32/f64 (OpenCV + numpy), f64 (geometry)", "data": "synthetic",
            "config": {"workload": "c5 on the host cores, bounded sample per step: " + desc},
            "cpu_baseline": {"value": round(v, 3), "unit": "Mpx/s", "cores": cores, "kind": kind, "sample": desc},
            "e2e": {"value": round(v, 3), "unit": "Mpx/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    _emit(line)


_REAL_STDOUT = None


def _protect_stdout():
    """stdout carries exactly ONE JSON line.  Libraries (NCCL's banner, a stray print in an extension) write to fd 1 as
    well, so fd 1 is pointed at stderr for the whole run and the line goes out through a private copy of the real one."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def _emit(line: dict):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    _protect_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--chunks", type=int, default=0, help="chunks of the pipelined host upload (e2e); 0 = one per ~169 tiles")
    ap.add_argument("--maps", type=int, default=MAPS_PER_STEP, help="maps batched per step")
    ap.add_argument("--map-side", type=int, default=MAP_SIDE)
    ap.add_argument("--objects-per-map", type=int, default=OBJECTS_PER_MAP)
    ap.add_argument("--no-extras", action="store_true", help="skip the single-GPU c4 / c1 / Otsu legs")
    ap.add_argument("--no-iou", action="store_true", help="skip the dense rotated-IoU throughput leg")
    ap.add_argument("--graph", default="auto", choices=["auto", "on", "off"],
                    help="replay the detection path (per-tile NMS + merge) as a CUDA graph; auto = on")
    args = ap.parse_args()
    if args.impl == "reference":
        reference(args)
    else:
        native(args)


if __name__ == "__main__":
    main()
