"""Generates the committed golden vectors from the REFERENCE ITSELF (run in the build container,
where /root/reference exists; the GPU box only reads the generated files).

  python tests/golden/make_golden.py

* pixel_golden.npz   - crops of Input/Test1.png / Test2.png and what the reference's own
                        build_multich(crop, 4) (lifted, cv2 4.13 IPP off, numpy 2.3) returns, plus
                        the Train twin's CHW output for one crop.
* xlsx_rows.json     - the 44 detection rows of the reference's Output/Test{1,2}.xlsx.
* merge_golden.json  - the lifted reference merge_detections / cross_scale_consensus_filter /
                        detect_symbols run on seeded synthetic detections (float64 Polygon stand-in
                        for shapely): kept indices in output order.
"""
import json
import os
import sys
import zipfile
import xml.etree.ElementTree as ET

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import cv2  # noqa: E402

cv2.ipp.setUseIPP(False)
from oracle import lift_reference as LR  # noqa: E402
from oracle import geometry as G  # noqa: E402

REF = LR.REFERENCE_ROOT


def read_xlsx(path):
    ns = {"m": "http://schemas.openxmlformats.org/spreadsheetml/2006/main"}
    z = zipfile.ZipFile(path)
    root = ET.fromstring(z.read("xl/worksheets/sheet1.xml"))
    rows = []
    for row in root.find("m:sheetData", ns):
        vals = []
        for c in row:
            if c.get("t") == "inlineStr":
                vals.append(c.find("m:is/m:t", ns).text)
            else:
                vals.append(float(c.find("m:v", ns).text))
        rows.append(vals)
    return rows[0], rows[1:]


def synth_dets(rng, n_obj, n_cls, extent, scale_tag=None):
    """Objects with 1-3 jittered copies each, fp32 values widened to Python floats."""
    dets = []
    for _ in range(n_obj):
        cx, cy = rng.uniform(0, extent, 2)
        w, h = rng.uniform(12, 100), rng.uniform(11, 97)
        th = rng.uniform(-np.pi / 4, 3 * np.pi / 4)
        cls = int(rng.integers(0, n_cls))
        conf = rng.uniform(0.2, 1.0)
        for _ in range(int(rng.integers(1, 4))):
            c, s = np.cos(th + rng.normal(0, 0.03)), np.sin(th + rng.normal(0, 0.03))
            ww, hh = w * rng.uniform(0.95, 1.05), h * rng.uniform(0.95, 1.05)
            x, y = cx + rng.normal(0, 2.0), cy + rng.normal(0, 2.0)
            v1, v2 = np.array([ww / 2 * c, ww / 2 * s]), np.array([-hh / 2 * s, hh / 2 * c])
            ctr = np.array([x, y])
            pts = np.concatenate([ctr + v1 + v2, ctr + v1 - v2, ctr - v1 - v2, ctr - v1 + v2]).astype(np.float32)
            cf = float(np.float32(np.clip(conf + rng.uniform(-0.05, 0.05), 0.01, 0.999)))
            dets.append(tuple(float(v) for v in pts) + (cls, cf, 0.0))
    order = rng.permutation(len(dets))
    return [dets[i] for i in order]


def main():
    det4 = LR.load_detect(4)
    train = LR.load_train()
    # ---------------- pixel vectors
    crops = {}
    t1 = cv2.imread(os.path.join(REF, "Input", "Test1.png"))
    t2 = cv2.imread(os.path.join(REF, "Input", "Test2.png"))
    picks = {
        "t1_416_ragged": t1[632:807, 632:895],      # 175 x 263
        "t1_128_full": t1[294:422, 392:520],
        "t1_128_noedge": t1[0:128, 0:128],           # the open removes every edge -> saturated DT
        "t1_128_sliver": t1[784:807, 882:895],       # 23 x 13
        "t2_416_crop": t2[316:732, 316:560],         # 416 x 244
        "t2_128_ragged": t2[980:1028, 980:1056],     # 48 x 76
        "const_5x7": np.full((5, 7, 3), 93, np.uint8),
        "row_1x40": t2[500:501, 300:340],
        "col_33x1": t2[500:533, 300:301],
    }
    out = {}
    for k, crop in picks.items():
        crop = np.ascontiguousarray(crop)
        out["in_" + k] = crop
        out["out_" + k] = det4.build_multich(crop, 4)
    out["chw_t1_128_full"] = train.build_4ch_CHW_from_bgr_dtedge(np.ascontiguousarray(picks["t1_128_full"]),
                                                                 sigmas=(0, 0.6, 1.2, 2.4))
    out["chw_default_sigmas_t2_128_ragged"] = train.build_4ch_CHW_from_bgr_dtedge(
        np.ascontiguousarray(picks["t2_128_ragged"]))
    np.savez_compressed(os.path.join(HERE, "pixel_golden.npz"), **out)

    # ---------------- xlsx rows
    rows = {}
    for name in ("Test1", "Test2"):
        cols, r = read_xlsx(os.path.join(REF, "Output", name + ".xlsx"))
        rows[name] = {"columns": cols, "rows": r}
    with open(os.path.join(HERE, "xlsx_rows.json"), "w") as fh:
        json.dump(rows, fh)

    # ---------------- merge / fusion / detect_symbols through the lifted reference
    det3 = LR.load_detect(3)
    gold = {}
    rng = np.random.default_rng(20251018)
    cases = {"sparse": (120, 15, 1500.0), "dense": (150, 3, 600.0), "mapscale": (120, 15, 16000.0)}
    gold["merge"] = {}
    for name, (n_obj, n_cls, extent) in cases.items():
        dets = synth_dets(rng, n_obj, n_cls, extent)
        work = list(dets)
        kept = det3.merge_detections(work, 0.4)
        ident = {id(d): i for i, d in enumerate(dets)}
        gold["merge"][name] = {"dets": dets, "sorted": [ident[id(d)] for d in work],
                               "kept": [ident[id(d)] for d in kept]}
    gold["fusion"] = {}
    for name, (n_obj, n_cls, extent) in {"two": (120, 6, 900.0), "three": (100, 4, 700.0)}.items():
        scales = [128, 416] if name == "two" else [64, 128, 416]
        base = synth_dets(rng, n_obj, n_cls, extent)
        by_scale = {s: [] for s in scales}
        for d in base:
            by_scale[scales[int(rng.integers(0, len(scales)))]].append(d)
        flat = [d for s in sorted(scales) for d in by_scale[s]]
        ident = {id(d): i for i, d in enumerate(flat)}
        kept = det3.cross_scale_consensus_filter({s: list(v) for s, v in by_scale.items()})
        gold["fusion"][name] = {"scales": scales, "by_scale": {str(s): by_scale[s] for s in scales},
                                "kept": [ident[id(d)] for d in kept]}
    # detect_symbols with a fake model that reports seeded tile-local boxes
    H, W, ts, ov = 700, 820, 128, 30
    img = np.zeros((H, W, 3), np.uint8)
    rng2 = np.random.default_rng(77)

    def fake(crop, conf):
        h, w = crop.shape[:2]
        n = int(rng2.integers(0, 7))
        corners, cls, cf = [], [], []
        for _ in range(n):
            cx, cy = rng2.uniform(0, w), rng2.uniform(0, h)
            bw, bh, th = rng2.uniform(10, 40), rng2.uniform(8, 30), rng2.uniform(-0.7, 2.3)
            c, s = np.cos(th), np.sin(th)
            v1, v2 = np.array([bw / 2 * c, bw / 2 * s]), np.array([-bh / 2 * s, bh / 2 * c])
            ctr = np.array([cx, cy])
            pts = np.stack([ctr + v1 + v2, ctr + v1 - v2, ctr - v1 - v2, ctr - v1 + v2]).astype(np.float32)
            for _ in range(int(rng2.integers(1, 3))):
                corners.append(pts + rng2.normal(0, 0.8, pts.shape).astype(np.float32))
                cls.append(int(rng2.integers(0, 3)))
                cf.append(float(rng2.uniform(0.25, 1)))
        o = np.argsort(-np.asarray(cf)) if cf else []
        rec.append({"corners": [np.asarray(corners[i]).reshape(-1).tolist() for i in o],
                    "cls": [cls[i] for i in o], "conf": [float(np.float32(cf[i])) for i in o]})
        return [corners[i] for i in o], [cls[i] for i in o], [cf[i] for i in o]

    rec = []
    model = LR.FakeModel(lambda crop, conf: tuple(np.asarray(a) if len(a) else np.zeros((0,)) for a in fake(crop, conf)))
    dets = det3.detect_symbols(img, model, ts, ov)
    gold["detect_symbols"] = {"H": H, "W": W, "tile": ts, "overlap": ov, "per_tile": rec,
                              "calls": [list(c[0]) for c in model.calls], "out": dets}
    with open(os.path.join(HERE, "merge_golden.json"), "w") as fh:
        json.dump(gold, fh)
    print("golden written:", {k: len(v) for k, v in gold.items()})


if __name__ == "__main__":
    main()
