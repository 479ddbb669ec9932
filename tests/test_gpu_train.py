"""f3: the training tilers' mirror (oriented_object_detection_b200.train) against what the reference's own
functions wrote for the same seeded dataset (tests/golden/train_golden.json, make_train_golden.py): label
tables (text, byte for byte), tile images, list files and empty-tile metadata."""
import hashlib
import json
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def gold():
    with open(os.path.join(HERE, "golden", "train_golden.json")) as fh:
        return json.load(fh)


@pytest.fixture()
def dataset(tmp_path):
    sys.path.insert(0, os.path.join(HERE, "golden"))
    from make_train_golden import make_dataset
    return make_dataset(str(tmp_path)) + (str(tmp_path),)


def _snapshot(out_img, out_lbl):
    labels = {fn: open(os.path.join(out_lbl, fn)).read() for fn in sorted(os.listdir(out_lbl))}
    images = {fn: hashlib.sha1(open(os.path.join(out_img, fn), "rb").read()).hexdigest() for fn in sorted(os.listdir(out_img))}
    return labels, images


@pytest.mark.parametrize("name,ts,ov", [("enum_128_50", 128, 50), ("enum_416_100", 416, 100)])
def test_enumerate_and_save_nonempty_tiles(cuda_dev, gold, dataset, name, ts, ov):
    from oriented_object_detection_b200 import train
    img_dir, lbl_dir, root = dataset
    oi, ol = os.path.join(root, name, "img"), os.path.join(root, name, "lbl")
    res = train.enumerate_and_save_nonempty_tiles(img_dir, lbl_dir, oi, ol, os.path.join(root, name + ".txt"), tile_size=ts,
                                                  overlap=ov, empty_meta_path=os.path.join(root, name + "_empty.json"))
    want = gold["runs"][name]
    labels, images = _snapshot(oi, ol)
    assert labels == want["labels"]
    assert images == want["images"]
    assert {"P_total": res["P_total"], "E_total": res["E_total"]} == want["result"]
    assert sorted(os.path.basename(p.strip()) for p in open(os.path.join(root, name + ".txt"))) == sorted(want["list"])
    key = lambda e: (e["image_file"], e["tile_id"])
    assert sorted(json.load(open(os.path.join(root, name + "_empty.json")))["empty"], key=key) == sorted(want["empty"], key=key)


@pytest.mark.parametrize("name,ts,ov,frac", [("crop_256_64_auto", 256, 64, None), ("crop_128_0_half", 128, 0, 0.5)])
def test_crop_images_and_labels(cuda_dev, gold, dataset, name, ts, ov, frac):
    from oriented_object_detection_b200 import train
    img_dir, lbl_dir, root = dataset
    oi, ol = os.path.join(root, name, "img"), os.path.join(root, name, "lbl")
    train.crop_images_and_labels(img_dir, lbl_dir, oi, ol, os.path.join(root, "unused.txt"), os.path.join(root, name + ".txt"),
                                 tile_size=ts, overlap=ov, keep_empty_fraction=frac, rng_seed=7)
    want = gold["runs"][name]
    labels, images = _snapshot(oi, ol)
    assert labels == want["labels"]
    assert images == want["images"]


def test_label_tables_random_against_numpy(cuda_dev):
    """20k labels on a 16384^2 map, 416/100: the device tables equal a direct numpy evaluation of the rule."""
    from oriented_object_detection_b200 import train
    rng = np.random.default_rng(3)
    H = W = 16384
    n = 20000
    c = rng.uniform(-50, H + 50, (n, 2)); wh = rng.uniform(8, 150, (n, 2)); th = rng.uniform(-1, 2.5, n)
    v1 = np.stack([wh[:, 0] / 2 * np.cos(th), wh[:, 0] / 2 * np.sin(th)], 1); v2 = np.stack([-wh[:, 1] / 2 * np.sin(th), wh[:, 1] / 2 * np.cos(th)], 1)
    q = np.concatenate([c + v1 + v2, c + v1 - v2, c - v1 - v2, c - v1 + v2], 1)
    labels = np.concatenate([rng.integers(0, 12, (n, 1)).astype(np.float64), q], 1)
    tables = train.tile_label_tables(labels, H, W, 416, 100, 0.1)
    rows, cols, span = train.full_tile_grid(H, W, 416, 100)
    assert (rows, cols, span) == (51, 51, 2)
    checked = 0
    for tile_id in list(tables.keys())[::37] + [0, rows * cols - 1]:
        y, x = (tile_id // cols) * 316, (tile_id % cols) * 316
        mx, my = (labels[:, 1] + labels[:, 7]) / 2, (labels[:, 2] + labels[:, 8]) / 2
        cand = (mx >= x) & (mx < x + 416) & (my >= y) & (my < y + 416)
        xs, ys = labels[:, 1::2], labels[:, 2::2]
        ax = np.maximum(0, np.minimum(xs.max(1), x + 416) - np.maximum(xs.min(1), x))
        ay = np.maximum(0, np.minimum(ys.max(1), y + 416) - np.maximum(ys.min(1), y))
        cov = ax * ay / np.maximum(1e-6, (xs.max(1) - xs.min(1)) * (ys.max(1) - ys.min(1)))
        keep = cand & (cov >= 0.1)
        want = labels[keep].copy()
        want[:, 1::2] = np.clip(want[:, 1::2] - x, 0, 416) / 416
        want[:, 2::2] = np.clip(want[:, 2::2] - y, 0, 416) / 416
        got = tables.get(tile_id, np.zeros((0, 9)))
        assert np.array_equal(got, want), tile_id
        checked += 1
    assert checked > 20
