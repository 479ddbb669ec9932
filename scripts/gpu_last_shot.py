#!/usr/bin/env python
"""A GPU check that fits into ~25 s of box time (one process, no pytest start-up): the changed rotated-IoU slab form
(throughput on the bench workload + the two device parity tests of tests/test_gpu_geom.py) and the Otsu branch
(golden vectors, degenerate tiles, one stage-by-stage plan).  Every result is appended to gpurun_out/last_shot.log
as soon as it exists, so a call that is cut off still reports what it reached."""
import json
import os
import sys
import time
import traceback

T0 = time.time()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
LOG = open(os.path.join(ROOT, "gpurun_out", "last_shot.log"), "w")


def log(**kw):
    kw["t"] = round(time.time() - T0, 2)
    LOG.write(json.dumps(kw) + "\n")
    LOG.flush()
    os.fsync(LOG.fileno())
    print(json.dumps(kw), flush=True)


def step(name, fn):
    try:
        r = fn()
        log(step=name, ok=True, **(r or {}))
    except Exception as e:          # noqa: BLE001 - report and go on to the next check
        log(step=name, ok=False, error=repr(e)[:400], tb=traceback.format_exc()[-800:])


import numpy as np  # noqa: E402
import torch  # noqa: E402

log(step="import torch", ok=True)
from oriented_object_detection_b200 import _lib, detect, ops, synth  # noqa: E402
from oracle import pixel as P  # noqa: E402

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
GOLD = os.path.join(ROOT, "tests", "golden")
pixel_golden = np.load(os.path.join(GOLD, "pixel_golden.npz"))
otsu_golden = np.load(os.path.join(GOLD, "otsu_golden.npz"))


def iou_throughput():
    plan = ops.make_plan(8192, 8192, 416, 100, device=dev)
    local, cls, conf, tid = synth.synthetic_tile_dets(plan, 59000, 15, seed=0, margin=20)
    nb = 8192
    bxh = local[:nb].astype(np.float64)
    bxh[:, 0::2] += plan.tiles["x0"][tid[:nb]][:, None]
    bxh[:, 1::2] += plan.tiles["y0"][tid[:nb]][:, None]
    bx = torch.from_numpy(bxh).to(dev)
    rs = torch.empty(nb, dtype=torch.float64, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ffma = ops.ffma_peak(8192)
    res = {"ffma_peak_tflops": round(ffma, 1)}
    for variant in (0, 1, 2, 3, 4):          # (rows per CTA, unroll): 0 = 256/1 (default), 1 = 64/1, 2 = 128/1, 3 = 256/2, 4 = 64/2
        os.environ["GM_IOU_VARIANT"] = str(variant)
        for _ in range(2):
            ops.rotated_iou_matrix_sum(bx, bx, out=rs)
        e0.record()
        for _ in range(5):
            ops.rotated_iou_matrix_sum(bx, bx, out=rs)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        gp = nb * nb / (ms * 1e-3) / 1e9
        res[f"v{variant}"] = {"gpairs_per_s": round(gp, 2), "ms": round(ms, 4),
                              "frac_of_measured_ffma": round(gp * 210 / 1e3 / ffma, 4), "checksum": float(rs.sum().item())}
    os.environ.pop("GM_IOU_VARIANT")
    return res


def iou_parity():
    import test_gpu_geom as T
    T.test_iou_pairs_matrix_and_checksum(dev)
    T.test_window_iou_coincident_edges_and_general_quads_gpu(dev)
    T.test_iou_degenerate_cases_and_host_api(dev)


def otsu_golden_vectors():
    detect.DT_BIN_METHOD = "otsu"
    try:
        bad = {}
        for k in pixel_golden.files:
            if k.startswith("in_"):
                got = detect.build_multich(pixel_golden[k], 4)
                n = int((got != otsu_golden["out_" + k[3:]]).sum())
                if n:
                    bad[k] = n
    finally:
        detect.DT_BIN_METHOD = "percentile"
    assert not bad, bad
    return {"crops": sum(k.startswith("in_") for k in pixel_golden.files)}


def otsu_rest():
    import test_gpu_zy_otsu as T
    T.test_otsu_degenerate_tiles_and_train_twin(dev, pixel_golden, otsu_golden)
    T.test_otsu_stage_by_stage(dev, 300, 300, 256, 64)
    T.test_otsu_stage_by_stage(dev, 700, 820, 128, 30)


def nms_quick():
    import test_gpu_geom as T
    T.test_nms_random_sets_against_oracle(dev, 900, 15, 4000, 0)
    T.test_tile_postprocess_synthetic_against_oracle(dev)


step("iou_throughput", iou_throughput)
step("otsu_golden_vectors", otsu_golden_vectors)
step("iou_parity", iou_parity)
step("otsu_rest", otsu_rest)
step("nms_quick", nms_quick)
log(step="done", ok=True)
