"""The random-init YOLO11n-OBB (yolo11_obb.py, the entry script's offline default; BASELINE config 1) behind the
device-resident batched predictor and through the root Detect_OBB.py script.  Sorts after the tests of the kernels it
sits between."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_yolo11n_behind_the_batched_predictor(cuda_dev):
    from oriented_object_detection_b200 import detect, ops, synth
    from oriented_object_detection_b200.predictor import TilePredictor
    from oriented_object_detection_b200.yolo11_obb import random_init_yolo11_obb
    H, W, ts, ov, ch, nc = 807, 895, 416, 100, 3, 12
    img = synth.synthetic_map_numpy(H, W, seed=2)
    pred = TilePredictor(random_init_yolo11_obb("n", nc, ch, ts, seed=0), ts, batch=8, device=cuda_dev)
    plan = ops.make_plan(H, W, ts, ov, device=cuda_dev)
    packed = ops.tile_gather(torch.from_numpy(img).to(cuda_dev), plan)
    b, c, f, tid = pred.predict_tiles(packed, plan, ch, 0.25)
    assert pred.net.calibrated
    assert b.shape[0] > 0 and b.shape[1] == 8 and bool(torch.isfinite(b).all())
    assert bool((tid[1:] >= tid[:-1]).all()) and int(c.min()) >= 0 and int(c.max()) < nc
    assert float(f.min()) > 0.25 and float(f.max()) <= 1.0
    same = tid[1:] == tid[:-1]
    assert bool((f[1:][same] <= f[:-1][same]).all())                 # confidence-descending inside a tile
    detect.channels = 3
    dets = detect.detect_symbols(img, pred, ts, ov)
    assert len(dets) > 0 and all(len(d) == 11 and isinstance(d[8], int) for d in dets)


def test_drop_in_script_with_yolo11n(cuda_dev, tmp_path, capsys, monkeypatch):
    import cv2
    import Detect_OBB as script
    from oriented_object_detection_b200 import detect, synth
    monkeypatch.setenv("GM_OFFLINE_MODEL", "yolo11n")
    inp, outp = tmp_path / "Input", tmp_path / "Output"
    inp.mkdir()
    cv2.imwrite(str(inp / "Test1.png"), synth.synthetic_map_numpy(807, 895, seed=1))
    script.input_dir, script.output_dir = str(inp), str(outp)
    script.calculate_metrics = False
    script.channels = 3
    detect.all_dets_per_image.clear()
    script.main()
    assert (outp / f"Test1_detected{script.OFFLINE_TAG}.jpg").exists() and (outp / f"Test1{script.OFFLINE_TAG}.xlsx").exists()
    out = capsys.readouterr().out
    assert "random-init YOLO11n-OBB" in out and "Processing Test1.png" in out
