"""csrc/dtedge_otsu.cuh compiled for the HOST (its functions are __host__ __device__): the per-tile arithmetic of the
Otsu binarisation (normalisation constants, acc8 map, Otsu scan, integer threshold on S) against the oracle, without a GPU."""
import os
import struct
import subprocess

import numpy as np
import pytest

from oracle import pixel as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("otsu") / "otsu_host")
    subprocess.run(["nvcc", "-O2", "-Wno-deprecated-gpu-targets", "-o", exe,
                    os.path.join(ROOT, "tests", "host_harness", "otsu_host.cu")], check=True)
    return exe


def _run(exe, s_list):
    blob = struct.pack("q", len(s_list))
    for S in s_list:
        flat = np.ascontiguousarray(S.reshape(-1), dtype=np.uint32)
        blob += struct.pack("q", flat.size) + flat.tobytes()
    out = subprocess.run([exe], input=blob, capture_output=True, check=True).stdout
    return np.frombuffer(out, dtype=np.uint32).reshape(-1, 6)


def _crops(pixel_golden):
    rng = np.random.default_rng(3)
    crops = [pixel_golden[k] for k in pixel_golden.files if k.startswith("in_")]
    crops += [rng.integers(0, 256, (40, 56, 3), dtype=np.uint8), np.full((16, 16, 3), 5, np.uint8)]
    two = np.full((32, 32, 3), 230, np.uint8)
    two[:, 16:] = 20
    crops.append(two)
    crops.append(np.repeat(np.arange(0, 240, 2, dtype=np.uint8)[None, :, None], 30, 0).repeat(3, 2))
    return crops


def test_otsu_tile_arithmetic_equals_oracle(harness, pixel_golden):
    crops = _crops(pixel_golden)
    stages = [P.dt_edge_stages(np.ascontiguousarray(c), bin_method="otsu") for c in crops]
    got = _run(harness, [st["S"] for st in stages])
    for st, g in zip(stages, got):
        S = st["S"].astype(np.int64)
        s_thr, thr8 = int(g[0]), int(g[1])
        assert thr8 == int(st["hi"])
        edges = (S >= s_thr) if s_thr != 0xFFFFFFFF else np.zeros(S.shape, bool)
        assert np.array_equal(edges, st["edges"])
        # both cv2.normalize constant pairs, bit for bit
        acc = st["acc"]
        smin, smax = float(acc.min()), float(acc.max())
        inv = 1.0 / (smax - smin) if (smax - smin) > np.finfo(np.float64).eps else 0.0
        for (a, b), (ks, kh) in (((0.0, 255.0), (2, 3)), ((0.0, 1.0), (4, 5))):
            scale = (b - a) * inv
            shift = a - smin * scale
            assert int(g[ks]) == int(np.float32(scale).view(np.uint32))
            assert int(g[kh]) == int(np.float32(shift).view(np.uint32))


def test_otsu_threshold_is_the_smallest_S_above(harness):
    """S >= s_thr <=> acc8(S) > thr for EVERY S of the range (not only the values present in the tile)."""
    rng = np.random.default_rng(9)
    tiles = [rng.integers(0, 5000, 4096).astype(np.uint32), (rng.integers(0, 300, 3000) ** 2).astype(np.uint32),
             np.concatenate([rng.integers(0, 50, 3000), rng.integers(100000, 400000, 500)]).astype(np.uint32)]
    got = _run(harness, tiles)
    for S, g in zip(tiles, got):
        a8 = P.acc8_from_acc(P.acc_from_S(S.astype(np.int64)))
        thr = P.otsu_threshold_u8(a8)
        assert int(g[1]) == thr
        assert np.array_equal(S.astype(np.int64) >= int(g[0]), a8 > thr)
        dense = np.arange(int(S.min()), int(S.max()) + 1, dtype=np.int64)
        fs = np.array([g[2]], np.uint32).view(np.float32)[0]
        fh = np.array([g[3]], np.uint32).view(np.float32)[0]
        v = P._round_sum_to_f32(np.sqrt(dense.astype(np.float32)).astype(np.float64) * np.float64(fs), np.float64(fh))
        d8 = np.trunc(v).clip(0, 255).astype(np.int64)
        assert np.array_equal(dense >= int(g[0]), d8 > thr)
