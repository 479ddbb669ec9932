"""Run the reference's OWN functions in this container.  TEST INFRASTRUCTURE ONLY.

``/root/reference/Detect_OBB.py`` cannot be imported: it needs ``ultralytics`` and
``shapely`` (neither installed nor in the offline wheelhouse), loads ``best128.pt`` at
import (Detect_OBB.py:26) and runs its main loop at module level (:745-761).  Instead the
file is parsed with ``ast``; imports (minus the two missing packages), constant
assignments (minus ``models`` / directories) and every ``def`` are kept, everything else is
dropped, and the result is exec'd into a fresh module whose ``Polygon`` / ``Point`` / ``YOLO``
names are injected.  The function bodies that run are the reference's, unmodified.

Used only to (a) validate the restatements in ``oracle/`` and (b) generate the vectors in
``tests/golden/`` (``tests/golden/make_golden.py``).  ``/root/reference`` does not exist on
the GPU box, so nothing that runs there imports this module's ``load_*`` functions.
"""
from __future__ import annotations

import ast
import os
import types

REFERENCE_ROOT = os.environ.get("GEOMAP_REFERENCE_ROOT", "/root/reference")
if not os.path.isfile(os.path.join(REFERENCE_ROOT, "Detect_OBB.py")):
    # GPU box: the two scripts as `make -C oracle` copied them (git-ignored oracle/_ref/; Input/ and Output/ do not travel,
    # so only the lifted FUNCTIONS are available there, not the reference's fixtures)
    _shipped = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
    if os.path.isfile(os.path.join(_shipped, "Detect_OBB.py")):
        REFERENCE_ROOT = _shipped

_DROP_IMPORTS = {"ultralytics", "shapely.geometry", "shapely"}
_DROP_ASSIGN = {"models", "input_dir", "output_dir"}


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "Detect_OBB.py"))


def fixtures_available() -> bool:
    """The reference's own images / spreadsheets (Input/, Output/): only in the build container."""
    return reference_available() and os.path.isfile(os.path.join(REFERENCE_ROOT, "Input", "Test1.png"))


def _lift(path: str, inject: dict) -> types.ModuleType:
    with open(path, "r", encoding="utf-8") as fh:
        tree = ast.parse(fh.read(), filename=path)
    body = []
    for node in tree.body:
        if isinstance(node, ast.Import):
            node.names = [a for a in node.names if a.name not in _DROP_IMPORTS]
            if node.names:
                body.append(node)
        elif isinstance(node, ast.ImportFrom):
            if node.module not in _DROP_IMPORTS:
                body.append(node)
        elif isinstance(node, ast.Assign):
            names = {t.id for t in node.targets if isinstance(t, ast.Name)}
            if not (names & _DROP_ASSIGN):
                body.append(node)
        elif isinstance(node, ast.FunctionDef):
            body.append(node)
        elif isinstance(node, ast.Expr) and isinstance(getattr(node, "value", None), ast.Constant):
            body.append(node)  # docstring
    mod = types.ModuleType("lifted_" + os.path.splitext(os.path.basename(path))[0])
    mod.__dict__.update(inject)
    code = compile(ast.Module(body=body, type_ignores=[]), path, "exec")
    exec(code, mod.__dict__)
    return mod


def load_detect(channels: int = 3) -> types.ModuleType:
    """The reference's Detect_OBB functions with the float64 Polygon stand-in injected."""
    from .geometry import Point, Polygon

    mod = _lift(os.path.join(REFERENCE_ROOT, "Detect_OBB.py"),
                {"Polygon": Polygon, "Point": Point, "YOLO": None})
    mod.channels = channels
    return mod


def load_train() -> types.ModuleType:
    return _lift(os.path.join(REFERENCE_ROOT, "Train_OBB.py"), {"YOLO": None})


class FakeBoxes:
    """Satisfies the three attributes Detect_OBB.py:228-231 reads from one detection."""

    def __init__(self, corners, cls, conf):
        import torch
        self.xyxyxyxy = torch.as_tensor(corners, dtype=torch.float32).reshape(1, 4, 2)
        self.cls = torch.as_tensor([cls], dtype=torch.float32)
        self.conf = torch.as_tensor([conf], dtype=torch.float32)


class FakeResult:
    def __init__(self, dets):
        self.obb = dets


class FakeModel:
    """Callable ``(ndarray, conf=) -> [result]``; ``fn(crop, conf)`` returns (corners[n,4,2], cls[n], conf[n])."""

    def __init__(self, fn):
        self.fn = fn
        self.calls = []

    def __call__(self, img, conf=0.25):
        self.calls.append((tuple(img.shape), str(img.dtype), bool(img.flags["C_CONTIGUOUS"]), conf))
        corners, cls, cf = self.fn(img, conf)
        return [FakeResult([FakeBoxes(corners[i], int(cls[i]), float(cf[i])) for i in range(len(cf))])]
