import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a) GPU; run with `pytest -m gpu`")
    try:
        import cv2
        cv2.ipp.setUseIPP(False)      # parity is defined against the IPP-off OpenCV path
    except Exception:
        pass


@pytest.fixture(scope="session")
def built_lib():
    """Builds the CUDA library (nvcc cross-compiles without a GPU) and imports the package."""
    import __graft_entry__ as g
    g.build()
    import oriented_object_detection_b200 as pkg
    return pkg


@pytest.fixture(scope="session")
def pixel_golden():
    return np.load(os.path.join(GOLDEN, "pixel_golden.npz"))


@pytest.fixture(scope="session")
def otsu_golden():
    return np.load(os.path.join(GOLDEN, "otsu_golden.npz"))


@pytest.fixture(scope="session")
def xlsx_rows():
    with open(os.path.join(GOLDEN, "xlsx_rows.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def merge_golden():
    with open(os.path.join(GOLDEN, "merge_golden.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def cuda_dev(built_lib):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from oriented_object_detection_b200 import _lib
    assert _lib.lib.gm_device_check() == 0, "not an sm_100 device"
    return torch.device("cuda:0")
