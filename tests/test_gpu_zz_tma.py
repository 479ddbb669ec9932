"""k_grad_fast<true>: the BGR patch of every block arrives by ONE cp.async.bulk.tensor.2d (TMA) when the map's row pitch is
a multiple of 16 bytes.  Bit-exact final bytes against the pixel oracle (oracle/pixel.py, itself pinned on the lifted
reference) on maps whose width qualifies, with every kind of block: interior, tile borders (REFLECT_101 mirror passes),
map borders (the TMA unit's zero fill at negative / past-the-end coordinates), ragged and tiny tiles."""
import numpy as np
import pytest
import torch

from oracle import pixel as P

pytestmark = pytest.mark.gpu


def _check(img, plan, dev, ops):
    m = torch.from_numpy(img).to(dev)
    assert (img.shape[1] * 3) % 16 == 0 and m.data_ptr() % 16 == 0          # the TMA path is the one that runs
    t4 = ops.dtedge_build(m, plan).cpu().numpy()
    bad = 0
    for t in plan.tiles:
        y0, x0, h, w, off = (int(t[k]) for k in ("y0", "x0", "h", "w", "px_off"))
        crop = np.ascontiguousarray(img[y0:y0 + h, x0:x0 + w])
        bad += int((t4[4 * off:4 * (off + h * w)].reshape(h, w, 4) != P.build_multich(crop, 4)).sum())
    return bad


def test_tma_gradient_bit_exact_on_aligned_maps(cuda_dev):
    from oriented_object_detection_b200 import ops, synth
    for (H, W, ts, ov, seed) in ((700, 656, 416, 100, 3), (420, 512, 128, 30, 4), (333, 1024, 416, 100, 5)):
        img = synth.synthetic_map_numpy(H, W, seed=seed)
        plan = ops.make_plan(H, W, ts, ov, device=cuda_dev)
        assert _check(img, plan, cuda_dev, ops) == 0, (H, W, ts)


def test_tma_gradient_on_tiny_and_edge_tiles(cuda_dev):
    from oriented_object_detection_b200 import ops, synth
    H, W = 200, 208
    img = synth.synthetic_map_numpy(H, W, seed=9)
    tiles = [(0, 0, 1, 1), (0, 0, 3, 7), (0, 200, 5, 8), (195, 0, 5, 9), (192, 200, 8, 8), (10, 10, 9, 100), (0, 0, 200, 208),
             (100, 37, 70, 33), (199, 207, 1, 1), (0, 150, 64, 58), (136, 0, 64, 40)]
    plan = ops.plan_from_tiles(H, W, tiles, device=cuda_dev)
    assert _check(img, plan, cuda_dev, ops) == 0


def test_stacked_bands_of_two_maps_equal_separate_builds(cuda_dev):
    """bench.py's batching: the row bands of several maps stacked in one buffer, one plan over the stack - every tile's
    bytes equal the build of its own map."""
    from oriented_object_detection_b200 import ops, synth
    H, W, ts, ov = 900, 1040, 416, 100
    maps = [torch.from_numpy(synth.synthetic_map_numpy(H, W, seed=20 + m)).to(cuda_dev) for m in range(2)]
    plan = ops.make_plan(H, W, ts, ov, device=cuda_dev)
    singles = [ops.dtedge_build(m, plan).clone() for m in maps]
    t = plan.tiles
    stack = torch.cat(maps, 0).contiguous()
    y = np.concatenate([t["y0"], t["y0"] + H])
    rep = lambda a: np.tile(np.asarray(a), 2)
    plan2 = ops.plan_from_arrays(2 * H, W, y, rep(t["x0"]), rep(t["h"]), rep(t["w"]), device=cuda_dev, tile_size=ts, overlap=ov)
    both = ops.dtedge_build(stack, plan2)
    assert torch.equal(both, torch.cat(singles))
    host = stack.cpu().pin_memory()
    out, _ = ops.build_tiles_from_host(host, plan2, 4, n_chunks=5)
    torch.cuda.synchronize()
    assert torch.equal(out, both)
