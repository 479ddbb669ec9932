"""YOLO11-OBB in plain PyTorch, random-init (BASELINE config 1: "random-init yolo11n-obb").

The reference builds its networks with ``YOLO("yolo11x-obb.pt")`` / ``YOLO("best128.pt")`` (Train_OBB.py:792,
Detect_OBB.py:26): the architecture is Ultralytics' ``yolo11-obb.yaml`` and lives in the third-party package
(ultralytics 8.3.196, requirements.txt:3), which is neither in the reference tree nor installable offline, and the
checkpoints are Google-Drive downloads.  The scope contract keeps the CNN forward in PyTorch, so this module restates
that architecture from its published definition (backbone Conv / C3k2 / SPPF / C2PSA, upsample-concat neck, decoupled
Detect head with DFL + the OBB angle branch; scales n / s / m / l / x) with random weights, and returns what the
Ultralytics head returns at inference: ``[B, 4 + nc + 1, A]`` = (cx, cy, w, h in input pixels, class probabilities,
theta) with anchors at cell centres of strides 8 / 16 / 32.  It is plumbing around the hot
path - the kernels under test are the ones before (tiler, DT-Edge, letterbox) and after it (decode, remap, NMS, fusion).

Pinned on what is published: the parameter counts Ultralytics prints for YOLO11 n / s / m / l / x (2,624,080 /
9,458,752 / 20,114,688 / 25,372,160 / 56,966,176 with nc = 80, Detect head) and for yolo11n-obb.yaml (2,695,747, nc = 80)
are reproduced exactly by ``tests/test_yolo11_obb.py``.  Numerical parity with an Ultralytics checkpoint is
**unpinned** (no weights offline).
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import torch
import torch.nn as nn

SCALES = {"n": (0.50, 0.25, 1024), "s": (0.50, 0.50, 1024), "m": (0.50, 1.00, 512),
          "l": (1.00, 1.00, 512), "x": (1.00, 1.50, 512)}          # depth, width, max_channels


def _make_divisible(x: float, divisor: int = 8) -> int:
    return int(math.ceil(x / divisor) * divisor)


def _autopad(k: int, p: Optional[int] = None) -> int:
    return k // 2 if p is None else p


class Conv(nn.Module):
    """Conv2d(bias=False) + BatchNorm2d(eps=1e-3, momentum=0.03) + SiLU."""

    def __init__(self, c1: int, c2: int, k: int = 1, s: int = 1, p: Optional[int] = None, g: int = 1, act: bool = True):
        super().__init__()
        self.conv = nn.Conv2d(c1, c2, k, s, _autopad(k, p), groups=g, bias=False)
        self.bn = nn.BatchNorm2d(c2, eps=1e-3, momentum=0.03)
        self.act = nn.SiLU(inplace=True) if act else nn.Identity()

    def forward(self, x):
        return self.act(self.bn(self.conv(x)))


class DWConv(Conv):
    def __init__(self, c1: int, c2: int, k: int = 1, s: int = 1, act: bool = True):
        super().__init__(c1, c2, k, s, g=math.gcd(c1, c2), act=act)


class Bottleneck(nn.Module):
    def __init__(self, c1: int, c2: int, shortcut: bool = True, g: int = 1, k: Sequence[int] = (3, 3), e: float = 0.5):
        super().__init__()
        c_ = int(c2 * e)
        self.cv1 = Conv(c1, c_, k[0], 1)
        self.cv2 = Conv(c_, c2, k[1], 1, g=g)
        self.add = shortcut and c1 == c2

    def forward(self, x):
        return x + self.cv2(self.cv1(x)) if self.add else self.cv2(self.cv1(x))


class C3k(nn.Module):
    """CSP bottleneck with three convolutions and n k x k bottlenecks."""

    def __init__(self, c1: int, c2: int, n: int = 1, shortcut: bool = True, g: int = 1, e: float = 0.5, k: int = 3):
        super().__init__()
        c_ = int(c2 * e)
        self.cv1 = Conv(c1, c_, 1, 1)
        self.cv2 = Conv(c1, c_, 1, 1)
        self.cv3 = Conv(2 * c_, c2, 1)
        self.m = nn.Sequential(*(Bottleneck(c_, c_, shortcut, g, k=(k, k), e=1.0) for _ in range(n)))

    def forward(self, x):
        return self.cv3(torch.cat((self.m(self.cv1(x)), self.cv2(x)), 1))


class C3k2(nn.Module):
    """C2f layout (split, n blocks chained, concat of every stage) with C3k or Bottleneck blocks."""

    def __init__(self, c1: int, c2: int, n: int = 1, c3k: bool = False, e: float = 0.5, g: int = 1, shortcut: bool = True):
        super().__init__()
        self.c = int(c2 * e)
        self.cv1 = Conv(c1, 2 * self.c, 1, 1)
        self.cv2 = Conv((2 + n) * self.c, c2, 1)
        self.m = nn.ModuleList(C3k(self.c, self.c, 2, shortcut, g) if c3k else Bottleneck(self.c, self.c, shortcut, g)
                               for _ in range(n))

    def forward(self, x):
        y = list(self.cv1(x).chunk(2, 1))
        y.extend(m(y[-1]) for m in self.m)
        return self.cv2(torch.cat(y, 1))


class SPPF(nn.Module):
    def __init__(self, c1: int, c2: int, k: int = 5):
        super().__init__()
        c_ = c1 // 2
        self.cv1 = Conv(c1, c_, 1, 1)
        self.cv2 = Conv(c_ * 4, c2, 1, 1)
        self.m = nn.MaxPool2d(kernel_size=k, stride=1, padding=k // 2)

    def forward(self, x):
        y = [self.cv1(x)]
        y.extend(self.m(y[-1]) for _ in range(3))
        return self.cv2(torch.cat(y, 1))


class Attention(nn.Module):
    """Multi-head self-attention over the positions of a feature map, with a depth-wise positional term."""

    def __init__(self, dim: int, num_heads: int = 8, attn_ratio: float = 0.5):
        super().__init__()
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        self.key_dim = int(self.head_dim * attn_ratio)
        self.scale = self.key_dim ** -0.5
        nh_kd = self.key_dim * num_heads
        self.qkv = Conv(dim, dim + nh_kd * 2, 1, act=False)
        self.proj = Conv(dim, dim, 1, act=False)
        self.pe = Conv(dim, dim, 3, 1, g=dim, act=False)

    def forward(self, x):
        B, C, H, W = x.shape
        N = H * W
        qkv = self.qkv(x)
        q, k, v = qkv.view(B, self.num_heads, self.key_dim * 2 + self.head_dim, N).split(
            [self.key_dim, self.key_dim, self.head_dim], dim=2)
        attn = (q.transpose(-2, -1) @ k) * self.scale
        attn = attn.softmax(dim=-1)
        x = (v @ attn.transpose(-2, -1)).view(B, C, H, W) + self.pe(v.reshape(B, C, H, W))
        return self.proj(x)


class PSABlock(nn.Module):
    def __init__(self, c: int, attn_ratio: float = 0.5, num_heads: int = 4, shortcut: bool = True):
        super().__init__()
        self.attn = Attention(c, attn_ratio=attn_ratio, num_heads=num_heads)
        self.ffn = nn.Sequential(Conv(c, c * 2, 1), Conv(c * 2, c, 1, act=False))
        self.add = shortcut

    def forward(self, x):
        x = x + self.attn(x) if self.add else self.attn(x)
        return x + self.ffn(x) if self.add else self.ffn(x)


class C2PSA(nn.Module):
    def __init__(self, c1: int, c2: int, n: int = 1, e: float = 0.5):
        super().__init__()
        assert c1 == c2
        self.c = int(c1 * e)
        self.cv1 = Conv(c1, 2 * self.c, 1, 1)
        self.cv2 = Conv(2 * self.c, c1, 1)
        self.m = nn.Sequential(*(PSABlock(self.c, attn_ratio=0.5, num_heads=self.c // 64) for _ in range(n)))

    def forward(self, x):
        a, b = self.cv1(x).split((self.c, self.c), dim=1)
        return self.cv2(torch.cat((a, self.m(b)), 1))


class OBBHead(nn.Module):
    """Detect head (box branch with DFL over reg_max = 16 bins, class branch with depth-wise stems) + the OBB
    angle branch (ne = 1).  ``forward`` returns the inference tensor [B, 4 + nc + 1, A]."""

    reg_max = 16

    def __init__(self, nc: int, ch: Sequence[int], ne: int = 1, strides: Sequence[int] = (8, 16, 32)):
        super().__init__()
        self.nc, self.ne, self.strides = nc, ne, tuple(strides)
        c2 = max(16, ch[0] // 4, self.reg_max * 4)
        c3 = max(ch[0], min(nc, 100))
        c4 = max(ch[0] // 4, ne)
        self.cv2 = nn.ModuleList(nn.Sequential(Conv(x, c2, 3), Conv(c2, c2, 3), nn.Conv2d(c2, 4 * self.reg_max, 1)) for x in ch)
        self.cv3 = nn.ModuleList(nn.Sequential(nn.Sequential(DWConv(x, x, 3), Conv(x, c3, 1)),
                                               nn.Sequential(DWConv(c3, c3, 3), Conv(c3, c3, 1)),
                                               nn.Conv2d(c3, nc, 1)) for x in ch)
        self.cv4 = nn.ModuleList(nn.Sequential(Conv(x, c4, 3), Conv(c4, c4, 3), nn.Conv2d(c4, ne, 1)) for x in ch)
        # the DFL projection: a frozen 1x1 convolution over the bins with weights 0..15 (counted as parameters upstream)
        self.dfl = nn.Parameter(torch.arange(self.reg_max, dtype=torch.float32), requires_grad=False)

    def bias_init(self, imgsz: int = 640, cls_bias: Optional[float] = None) -> None:
        """Ultralytics' Detect.bias_init: box bias 1.0, class bias log(5 / nc / (imgsz / stride)^2) (or ``cls_bias``)."""
        for a, b, s in zip(self.cv2, self.cv3, self.strides):
            a[-1].bias.data[:] = 1.0
            b[-1].bias.data[:self.nc] = math.log(5 / self.nc / (imgsz / s) ** 2) if cls_bias is None else cls_bias

    def forward(self, feats: List[torch.Tensor]) -> torch.Tensor:
        outs = []
        for f, box, cls, ang, s in zip(feats, self.cv2, self.cv3, self.cv4, self.strides):
            b, _, gh, gw = f.shape
            ys, xs = torch.meshgrid(torch.arange(gh, device=f.device, dtype=torch.float32),
                                    torch.arange(gw, device=f.device, dtype=torch.float32), indexing="ij")
            ax, ay = xs.reshape(1, -1) + 0.5, ys.reshape(1, -1) + 0.5               # anchors, grid units
            d = box(f).float().view(b, 4, self.reg_max, gh * gw).softmax(2)
            ltrb = (d * self.dfl.view(1, 1, -1, 1)).sum(2)                          # DFL: expectation over the bins
            lt, rb = ltrb[:, 0:2], ltrb[:, 2:4]
            theta = (ang(f).float().view(b, self.ne, gh * gw).sigmoid() - 0.25) * math.pi
            cos, sin = torch.cos(theta[:, 0]), torch.sin(theta[:, 0])
            xf, yf = (rb[:, 0] - lt[:, 0]) / 2, (rb[:, 1] - lt[:, 1]) / 2            # dist2rbox
            cx = (xf * cos - yf * sin + ax) * s
            cy = (xf * sin + yf * cos + ay) * s
            wh = (lt + rb) * s
            outs.append(torch.cat([cx.unsqueeze(1), cy.unsqueeze(1), wh, cls(f).float().view(b, self.nc, gh * gw).sigmoid(),
                                   theta], dim=1))
        return torch.cat(outs, dim=2)


class YOLO11OBB(nn.Module):
    """``yolo11-obb.yaml`` at one of the scales n / s / m / l / x; ``ch`` = 3, or 4 for the [R,G,B,DT-Edge] input
    (the dataset yaml's ``channels: 4`` switches the first convolution, datasets/GeoMap/data4ch.yaml:15)."""

    def __init__(self, scale: str = "n", nc: int = 12, ch: int = 3, imgsz: int = 640, cls_bias: Optional[float] = None,
                 cls_gain: float = 1.0):
        super().__init__()
        depth, width, max_ch = SCALES[scale]
        c3k_all = scale in "mlx"                      # parse_model forces c3k = True for the larger scales

        def C(c):
            return _make_divisible(min(c, max_ch) * width, 8)

        def R(n):
            return max(round(n * depth), 1) if n > 1 else n

        self.nc = nc
        c64, c128, c256, c512, c1024 = C(64), C(128), C(256), C(512), C(1024)
        self.b0 = Conv(ch, c64, 3, 2)                                   # 0  P1/2
        self.b1 = Conv(c64, c128, 3, 2)                                 # 1  P2/4
        self.b2 = C3k2(c128, c256, R(2), c3k_all, 0.25)                 # 2
        self.b3 = Conv(c256, c256, 3, 2)                                # 3  P3/8
        self.b4 = C3k2(c256, c512, R(2), c3k_all, 0.25)                 # 4
        self.b5 = Conv(c512, c512, 3, 2)                                # 5  P4/16
        self.b6 = C3k2(c512, c512, R(2), True)                          # 6
        self.b7 = Conv(c512, c1024, 3, 2)                               # 7  P5/32
        self.b8 = C3k2(c1024, c1024, R(2), True)                        # 8
        self.b9 = SPPF(c1024, c1024, 5)                                 # 9
        self.b10 = C2PSA(c1024, c1024, R(2))                            # 10
        self.up = nn.Upsample(scale_factor=2, mode="nearest")
        self.h13 = C3k2(c1024 + c512, c512, R(2), c3k_all)              # 13
        self.h16 = C3k2(c512 + c512, c256, R(2), c3k_all)               # 16 (P3/8)
        self.h17 = Conv(c256, c256, 3, 2)
        self.h19 = C3k2(c256 + c512, c512, R(2), c3k_all)               # 19 (P4/16)
        self.h20 = Conv(c512, c512, 3, 2)
        self.h22 = C3k2(c512 + c1024, c1024, R(2), True)                # 22 (P5/32)
        self.head = OBBHead(nc, (c256, c512, c1024))
        self.head.bias_init(imgsz, cls_bias)
        if cls_gain != 1.0:       # random-init class logits are nearly constant: spread them so a harness gets detections
            with torch.no_grad():
                for b in self.head.cv3:
                    b[-1].weight.mul_(cls_gain)

    def features(self, x: torch.Tensor) -> List[torch.Tensor]:
        x = self.b1(self.b0(x))
        p3 = self.b4(self.b3(self.b2(x)))
        p4 = self.b6(self.b5(p3))
        p5 = self.b10(self.b9(self.b8(self.b7(p4))))
        n4 = self.h13(torch.cat((self.up(p5), p4), 1))
        n3 = self.h16(torch.cat((self.up(n4), p3), 1))
        o4 = self.h19(torch.cat((self.h17(n3), n4), 1))
        o5 = self.h22(torch.cat((self.h20(o4), p5), 1))
        return [n3, o4, o5]

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.head(self.features(x))


@torch.no_grad()
def calibrate_batchnorm(net: nn.Module, x: torch.Tensor, passes: int = 1) -> nn.Module:
    """Random-init weights with FRESH BatchNorm statistics (mean 0, var 1) shrink the activations layer after layer
    until the head sees ~1e-6 and every anchor gets the same output.  A harness that wants a random-init network to
    behave like a network (outputs that depend on the input) first sets the running statistics from a sample batch:
    cumulative averages over ``passes`` forward passes in train mode, no gradient, then eval mode."""
    saved = {}
    for mod in net.modules():
        if isinstance(mod, nn.BatchNorm2d):
            mod.reset_running_stats()
            saved[mod] = mod.momentum
            mod.momentum = None                     # cumulative moving average
    net.train()
    for _ in range(passes):
        net(x)
    for mod, mom in saved.items():
        mod.momentum = mom
    return net.eval()


class SelfCalibrating(nn.Module):
    """Wraps a random-init network: the FIRST batch it sees sets the BatchNorm statistics (``calibrate_batchnorm`` on up
    to ``max_batch`` samples of it), every call after that is a plain eval-mode forward.  The statistics have to come
    from the distribution the network will see - calibrated on one kind of image and run on another, a random-init
    network saturates (every anchor fires or none does)."""

    def __init__(self, net: nn.Module, max_batch: int = 32):
        super().__init__()
        self.net = net.eval()
        self.max_batch = max_batch
        self.calibrated = False

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not self.calibrated:
            calibrate_batchnorm(self.net, x[:self.max_batch])
            self.calibrated = True
        return self.net(x)


def random_init_yolo11_obb(scale: str = "n", nc: int = 12, ch: int = 3, imgsz: int = 416, calib: Optional[torch.Tensor] = None,
                           cls_bias: Optional[float] = -2.7, seed: int = 0) -> nn.Module:
    """The offline stand-in for ``YOLO("best*.pt")``: the real architecture, seeded random weights, BatchNorm statistics
    from ``calib`` ([B, ch, H, W] in [0, 1]) - or, if None, from the first batch the network is called on
    (:class:`SelfCalibrating`) - and a class bias under which roughly 1-5 % of the anchors pass ``conf = 0.25``
    (Ultralytics' own bias init, ``cls_bias=None``, gives probabilities ~1e-4 and no detections at all).  The detections
    mean nothing; the data path around the network is the real one."""
    with torch.random.fork_rng(devices=[]):
        torch.manual_seed(seed)
        net = YOLO11OBB(scale, nc, ch, imgsz, cls_bias=cls_bias)
    if calib is None:
        return SelfCalibrating(net)
    return calibrate_batchnorm(net, calib.to(next(net.parameters()).device))
