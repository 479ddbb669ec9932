"""The root Detect_OBB.py entry script fails loudly, like the reference (YOLO("best128.pt") raises on a missing
checkpoint, Detect_OBB.py:26): the random-init predictors are an explicit opt-in and tag their outputs."""
import pytest


def test_load_models_raises_without_checkpoints(monkeypatch, tmp_path):
    import Detect_OBB as script
    monkeypatch.delenv("GM_OFFLINE_MODEL", raising=False)
    monkeypatch.chdir(tmp_path)                      # no best128.pt / best416.pt here
    with pytest.raises(FileNotFoundError, match="best128.pt"):
        script.load_models()
    monkeypatch.setenv("GM_OFFLINE_MODEL", "resnet")
    with pytest.raises(ValueError):
        script.load_models()


def test_offline_outputs_are_tagged():
    import Detect_OBB as script
    from oriented_object_detection_b200 import detect
    assert script.OFFLINE_TAG and detect.output_tag == "" or detect.output_tag == script.OFFLINE_TAG
