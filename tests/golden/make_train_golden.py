"""Golden vectors of the training tilers (Train_OBB.py:44-146, :290-428) from the REFERENCE ITSELF: the lifted
functions run on a seeded synthetic dataset in a temp directory; the label files, list files and the empty-tile
metadata they write are recorded.

  python tests/golden/make_train_golden.py        (build container only: needs /root/reference)
"""
import hashlib
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import cv2  # noqa: E402

from oracle import lift_reference as LR  # noqa: E402
from oriented_object_detection_b200 import synth  # noqa: E402

IMAGES = {"mapA.png": (700, 900, 31), "mapB.jpg": (416, 500, 32), "tiny.png": (100, 90, 33)}


def make_dataset(root):
    """Images + YOLO-OBB label files (normalised), deterministic."""
    img_dir, lbl_dir = os.path.join(root, "images"), os.path.join(root, "labels")
    os.makedirs(img_dir); os.makedirs(lbl_dir)
    rng = np.random.default_rng(2024)
    for name, (H, W, seed) in IMAGES.items():
        cv2.imwrite(os.path.join(img_dir, name), synth.synthetic_map_numpy(H, W, seed=seed))
        lines = ["# a comment line"]
        for _ in range(0 if name == "tiny.png" else 90):
            cx, cy = rng.uniform(-10, W + 10), rng.uniform(-10, H + 10)
            w, h, th = rng.uniform(10, 120), rng.uniform(8, 90), rng.uniform(-0.7, 2.3)
            c, s = np.cos(th), np.sin(th)
            v1, v2 = np.array([w / 2 * c, w / 2 * s]), np.array([-h / 2 * s, h / 2 * c])
            q = np.stack([np.array([cx, cy]) + a for a in (v1 + v2, v1 - v2, -v1 - v2, -v1 + v2)])
            q = q / np.array([W, H])
            lines.append(f"{int(rng.integers(0, 12))} " + " ".join(f"{v:.6f}" for v in q.reshape(-1)))
        lines.insert(5, "3 0.5 0.5 oops") if len(lines) > 5 else None
        with open(os.path.join(lbl_dir, os.path.splitext(name)[0] + ".txt"), "w") as fh:
            fh.write("\n".join(lines) + "\n")
    return img_dir, lbl_dir


def snapshot(out_img, out_lbl):
    labels = {}
    for fn in sorted(os.listdir(out_lbl)):
        with open(os.path.join(out_lbl, fn)) as fh:
            labels[fn] = fh.read()
    images = {fn: hashlib.sha1(open(os.path.join(out_img, fn), "rb").read()).hexdigest() for fn in sorted(os.listdir(out_img))}
    return {"labels": labels, "images": images}


def main():
    ref = LR.load_train()
    gold = {"images": IMAGES, "runs": {}}
    with tempfile.TemporaryDirectory() as root:
        img_dir, lbl_dir = make_dataset(root)
        for name, (ts, ov) in {"enum_128_50": (128, 50), "enum_416_100": (416, 100)}.items():
            oi, ol = os.path.join(root, name, "img"), os.path.join(root, name, "lbl")
            res = ref.enumerate_and_save_nonempty_tiles(img_dir, lbl_dir, oi, ol, os.path.join(root, name + ".txt"),
                                                        tile_size=ts, overlap=ov,
                                                        empty_meta_path=os.path.join(root, name + "_empty.json"))
            snap = snapshot(oi, ol)
            snap["result"] = {"P_total": res["P_total"], "E_total": res["E_total"]}
            snap["list"] = [os.path.basename(p.strip()) for p in open(os.path.join(root, name + ".txt"))]
            snap["empty"] = json.load(open(os.path.join(root, name + "_empty.json")))["empty"]
            gold["runs"][name] = snap
        for name, (ts, ov, frac) in {"crop_256_64_auto": (256, 64, None), "crop_128_0_half": (128, 0, 0.5)}.items():
            oi, ol = os.path.join(root, name, "img"), os.path.join(root, name, "lbl")
            ref.crop_images_and_labels(img_dir, lbl_dir, oi, ol, os.path.join(root, "unused.txt"),
                                       os.path.join(root, name + ".txt"), tile_size=ts, overlap=ov,
                                       keep_empty_fraction=frac, rng_seed=7)
            snap = snapshot(oi, ol)
            snap["list"] = [os.path.basename(p.strip()) for p in open(os.path.join(root, name + ".txt"))]
            gold["runs"][name] = snap
    with open(os.path.join(HERE, "train_golden.json"), "w") as fh:
        json.dump(gold, fh)
    print({k: (len(v["labels"]), len(v["images"])) for k, v in gold["runs"].items()})


if __name__ == "__main__":
    main()
