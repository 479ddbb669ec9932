"""The PyTorch restatement of yolo11-obb.yaml (oriented_object_detection_b200/yolo11_obb.py): published parameter
counts, head layout (anchor order, DFL expectation, dist2rbox, angle range), 4-channel input, BatchNorm calibration,
and that the oracle's decode restatement consumes its output.  CPU only."""
import math

import numpy as np
import pytest
import torch

from oracle import decode as D


@pytest.fixture(scope="module")
def Y(built_lib):
    from oriented_object_detection_b200 import yolo11_obb
    return yolo11_obb


@pytest.mark.parametrize("scale,published", [("n", 2_624_080), ("s", 9_458_752), ("m", 20_114_688),
                                             ("l", 25_372_160), ("x", 56_966_176)])
def test_detect_variant_parameter_counts_equal_the_published_ones(Y, scale, published):
    """`YOLO11<scale> summary: ... parameters` as Ultralytics prints it (nc = 80, Detect head = everything but the angle branch)."""
    m = Y.YOLO11OBB(scale, nc=80)
    total = sum(p.numel() for p in m.parameters())
    angle = sum(p.numel() for p in m.head.cv4.parameters())
    assert total - angle == published


def test_obb_parameter_count_of_the_n_scale(Y):
    assert sum(p.numel() for p in Y.YOLO11OBB("n", nc=80).parameters()) == 2_695_747      # `yolo11n-obb.yaml` summary (nc: 80)


@pytest.mark.parametrize("size,anchors", [(416, 52 * 52 + 26 * 26 + 13 * 13), (128, 16 * 16 + 8 * 8 + 4 * 4)])
def test_head_layout(Y, size, anchors):
    torch.manual_seed(1)
    m = Y.YOLO11OBB("n", nc=12, ch=3).eval()
    # box branch: every side picks bin 5 of the DFL; angle logit 0 -> theta = pi/4; lt == rb -> centre = anchor
    for box, ang in zip(m.head.cv2, m.head.cv4):
        box[-1].weight.data.zero_()
        box[-1].bias.data[:] = torch.tensor([0.0] * 5 + [60.0] + [0.0] * 10).repeat(4)
        ang[-1].weight.data.zero_()
        ang[-1].bias.data.zero_()
    with torch.no_grad():
        o = m(torch.rand(2, 3, size, size))
    assert o.shape == (2, 4 + 12 + 1, anchors) and o.dtype == torch.float32
    at = 0
    for s in (8, 16, 32):
        g = size // s
        ys, xs = np.divmod(np.arange(g * g), g)
        blk = o[0, :, at:at + g * g].numpy()
        assert np.allclose(blk[0], (xs + 0.5) * s, atol=1e-3) and np.allclose(blk[1], (ys + 0.5) * s, atol=1e-3)
        assert np.allclose(blk[2:4], 10.0 * s, atol=1e-3)                       # w = h = (l + r) * stride = 2 * 5 * s
        assert np.allclose(blk[16], math.pi / 4, atol=1e-6)
        at += g * g
    assert float(o[:, 4:16].min()) >= 0.0 and float(o[:, 4:16].max()) <= 1.0


def test_dist2rbox_offsets_rotate_with_theta(Y):
    m = Y.YOLO11OBB("n", nc=3).eval()
    for box, ang in zip(m.head.cv2, m.head.cv4):
        box[-1].weight.data.zero_()
        b = torch.zeros(4, 16)
        b[0, 2] = b[1, 2] = 60.0            # l = t = 2
        b[2, 6] = 60.0                      # r = 6
        b[3, 2] = 60.0                      # b = 2
        box[-1].bias.data[:] = b.reshape(-1)
        ang[-1].weight.data.zero_()
        ang[-1].bias.data[:] = math.log(3.0)          # sigmoid = 0.75 -> theta = pi/2
    with torch.no_grad():
        o = m(torch.rand(1, 3, 128, 128))[0].numpy()
    # xf = (r - l) / 2 = 2, yf = 0; theta = pi/2: (xf cos - yf sin, xf sin + yf cos) = (0, 2)
    assert np.allclose(o[0, 0], 0.5 * 8, atol=1e-3) and np.allclose(o[1, 0], (0.5 + 2.0) * 8, atol=1e-3)
    assert np.allclose(o[2, 0], 8.0 * 8, atol=1e-3) and np.allclose(o[3, 0], 4.0 * 8, atol=1e-3)


def test_four_channel_input_and_rect_shapes(Y):
    m = Y.YOLO11OBB("n", nc=12, ch=4).eval()
    assert m.b0.conv.weight.shape[1] == 4
    with torch.no_grad():
        o = m(torch.rand(1, 4, 288, 416))            # a rect-letterboxed ragged tile (175 x 263 at 416)
    assert o.shape == (1, 17, 36 * 52 + 18 * 26 + 9 * 13)


def _smooth_batch(b, ch, size, seed):
    """Image-like random batch: noise at every octave, so every pyramid level of the network sees variance."""
    gen = torch.Generator().manual_seed(seed)
    out = 0.3 * torch.rand(b, ch, size, size, generator=gen)
    k = 2
    while size // k >= 2:
        out += torch.nn.functional.interpolate(torch.rand(b, ch, size // k, size // k, generator=gen), size=(size, size),
                                               mode="bilinear", align_corners=False)
        k *= 2
    out -= out.amin((1, 2, 3), keepdim=True)
    return out / out.amax((1, 2, 3), keepdim=True)


def test_random_init_factory_gives_input_dependent_detections(Y):
    net = Y.random_init_yolo11_obb("n", nc=12, ch=3, imgsz=128, seed=0)
    again = Y.random_init_yolo11_obb("n", nc=12, ch=3, imgsz=128, seed=0)
    first = _smooth_batch(16, 3, 128, 1)
    x = _smooth_batch(4, 3, 128, 5)
    with torch.no_grad():
        net(first), again(first)                                          # the first batch sets the BatchNorm statistics
        o, o2 = net(x), again(x)
    assert net.calibrated and torch.equal(o, o2)                          # seeded: reproducible
    best = o[:, 4:16].max(1)[0]
    frac = float((best > 0.25).float().mean())
    assert 0.002 < frac < 0.4, frac                                       # some anchors pass conf 0.25, most do not
    assert float(o[:, 16].min()) >= -math.pi / 4 - 1e-6 and float(o[:, 16].max()) <= 3 * math.pi / 4 + 1e-6
    assert float((o[0] - o[1]).abs().max()) > 1e-3                        # outputs depend on the input
    fresh = Y.YOLO11OBB("n", nc=12, cls_bias=-2.7).eval()                 # no calibration: a dead network
    with torch.no_grad():
        f = fresh(x)
    assert float(f[:, 4:16].std()) < 1e-3
    # the decode restatement (Ultralytics' predictor tail) takes the head as it is
    total = 0
    for k in range(o.shape[0]):
        boxes, cls, conf = D.decode_tile(o[k].numpy(), 128, 128, (128, 128), 0.25, 0.7, 300)
        assert len(conf) == len(cls) == len(boxes) and np.isfinite(np.asarray(boxes)).all()
        assert all(conf[i] >= conf[i + 1] for i in range(len(conf) - 1))
        total += len(conf)
    assert total > 0


def test_calibration_on_map_tiles_transfers_to_other_tiles_of_the_map(Y):
    from oriented_object_detection_b200 import synth
    img = synth.synthetic_map_numpy(832, 1248, seed=5)
    tiles = np.stack([img[y:y + 416, x:x + 416] for y in (0, 416) for x in (0, 416, 832)])
    x = torch.from_numpy(tiles[..., ::-1].copy()).permute(0, 3, 1, 2).float() / 255
    net = Y.random_init_yolo11_obb("n", nc=12, ch=3, imgsz=416, calib=x[:4], seed=0)
    with torch.no_grad():
        o = net(x[4:])
    frac = float((o[:, 4:16].max(1)[0] > 0.25).float().mean())
    assert 0.001 < frac < 0.3, frac
