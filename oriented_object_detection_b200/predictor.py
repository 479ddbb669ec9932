"""Device-resident batched predictor around the (PyTorch) OBB network.

The reference calls ``model(tile, conf=...)`` once per tile (Detect_OBB.py:76-85, :225-231): a host
round trip, a batch of one and Ultralytics' Python post-processing per call.  ``TilePredictor`` keeps
the reference's model protocol available (it is what ``detect.detect_symbols`` calls through
``predict_tiles``) but works on the packed tile batch in device memory:

    tiles of one shape -> gm_letterbox_tiles (LetterBox, BGR->RGB, CHW, /255; one launch)
                       -> net(x)  (PyTorch; [B, C, H, W] -> raw head [B, 4 + nc + 1, A])
                       -> gm_decode_tiles (confidence filter, probiou fast-NMS, un-letterbox, corners)

The CNN itself stays in PyTorch (scope contract): any ``nn.Module`` with that input /
output convention works - an exported Ultralytics ``model.model`` when the package and a checkpoint are
available, or ``StandInOBBNet`` (random weights, same head layout) for offline runs, since the
reference's best*.pt checkpoints are not obtainable without network access.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np
import torch
import torch.nn as nn

from . import ops


class StandInOBBNet(nn.Module):
    """A small random-init CNN with the YOLO-OBB head layout: strides 8/16/32, anchors at cell centres,
    output [B, 4 + nc + 1, A] = (cx, cy, w, h in input pixels, class probabilities, theta).  NOT the
    YOLOv11 architecture - a stand-in so the batched path can run end to end without a checkpoint."""

    def __init__(self, channels: int = 3, n_classes: int = 12, width: int = 16, hot_bias: float = -2.5,
                 cls_gain: float = 200.0):
        super().__init__()
        self.nc = n_classes

        def block(ci, co, s):
            return nn.Sequential(nn.Conv2d(ci, co, 3, s, 1, bias=False), nn.BatchNorm2d(co), nn.SiLU())

        self.stem = nn.Sequential(block(channels, width, 2), block(width, 2 * width, 2), block(2 * width, 4 * width, 2))
        self.down4 = block(4 * width, 8 * width, 2)
        self.down5 = block(8 * width, 16 * width, 2)
        self.heads = nn.ModuleList([nn.Conv2d(c, 4 + n_classes + 1, 1) for c in (4 * width, 8 * width, 16 * width)])
        for h in self.heads:                       # class logits spread and biased so that ~1 % of the anchors pass conf 0.25
            nn.init.constant_(h.bias[4:4 + n_classes], hot_bias)
            with torch.no_grad():
                h.weight[4:4 + n_classes] *= cls_gain

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        p3 = self.stem(x)
        p4 = self.down4(p3)
        p5 = self.down5(p4)
        outs = []
        for f, head, s in zip((p3, p4, p5), self.heads, (8, 16, 32)):
            o = head(f)
            b, _, gh, gw = o.shape
            ys, xs = torch.meshgrid(torch.arange(gh, device=o.device), torch.arange(gw, device=o.device), indexing="ij")
            ax = (xs.reshape(1, -1) + 0.5) * s
            ay = (ys.reshape(1, -1) + 0.5) * s
            o = o.reshape(b, 4 + self.nc + 1, gh * gw)
            lt, rb = torch.nn.functional.softplus(o[:, 0:2]), torch.nn.functional.softplus(o[:, 2:4])
            theta = (torch.sigmoid(o[:, 4 + self.nc]) - 0.25) * np.pi
            xf, yf = (rb[:, 0] - lt[:, 0]) / 2, (rb[:, 1] - lt[:, 1]) / 2
            cx = (xf * torch.cos(theta) - yf * torch.sin(theta)) * s + ax
            cy = (xf * torch.sin(theta) + yf * torch.cos(theta)) * s + ay
            wh = (lt + rb) * s
            outs.append(torch.cat([cx.unsqueeze(1), cy.unsqueeze(1), wh, torch.sigmoid(o[:, 4:4 + self.nc]),
                                   theta.unsqueeze(1)], dim=1))
        return torch.cat(outs, dim=2)


class TilePredictor:
    """``predict_tiles(packed, plan, channels, conf)`` for ``detect.detect_symbols``; also callable on one host
    crop with the reference's ``model(ndarray, conf=)`` protocol."""

    def __init__(self, net: nn.Module, imgsz: int, iou: float = 0.7, max_det: int = 300, batch: int = 256,
                 stride: int = 32, auto: bool = True, device=None):
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.net = net.to(self.device).eval()
        self.imgsz, self.iou, self.max_det, self.batch, self.stride, self.auto = imgsz, iou, max_det, batch, stride, auto

    @torch.no_grad()
    def predict_tiles(self, packed: torch.Tensor, plan: ops.TilePlan, channels: int, conf: float):
        """-> (tile-local corners float32 [n,8], cls int32 [n], conf float32 [n], tile_id int32 [n]); tile_id is
        non-decreasing and detections of a tile are confidence-descending, like the reference's per-tile lists."""
        groups: Dict[Tuple[int, int], List[int]] = {}
        for ti, t in enumerate(plan.tiles):
            groups.setdefault((int(t["h"]), int(t["w"])), []).append(ti)
        parts = []
        for (h, w), idx in groups.items():
            for b0 in range(0, len(idx), self.batch):
                sub = idx[b0:b0 + self.batch]
                sub_t = torch.tensor(sub, dtype=torch.int64)
                x = ops.letterbox_tiles(packed, plan, sub_t, channels, self.imgsz, self.stride, self.auto)
                head = self.net(x).float().contiguous()
                sub_plan = ops.plan_from_tiles(plan.H, plan.W, [(int(plan.tiles["y0"][i]), int(plan.tiles["x0"][i]), h, w)
                                                                for i in sub], device=self.device)
                boxes, cls, cf, count = ops.decode_tiles(head, sub_plan, (x.shape[2], x.shape[3]), conf, self.iou, self.max_det)
                b, c, f, tid = ops.compact_decoded(boxes, cls, cf, count, self.max_det)
                parts.append((b, c, f, sub_t.to(self.device)[tid.to(torch.int64)].to(torch.int32)))
        if not parts:
            z = torch.zeros(0, device=self.device)
            return z.reshape(0, 8), z.to(torch.int32), z, z.to(torch.int32)
        b = torch.cat([p[0] for p in parts]); c = torch.cat([p[1] for p in parts])
        f = torch.cat([p[2] for p in parts]); t = torch.cat([p[3] for p in parts])
        order = torch.sort(t, stable=True)[1]           # groups were visited by shape: restore tile order
        return b[order], c[order], f[order], t[order]

    def __call__(self, img: np.ndarray, conf: float = 0.25):
        """Reference protocol on one crop: returns [result] with ``result.obb`` items (.xyxyxyxy/.cls/.conf)."""
        crop = np.ascontiguousarray(img)
        h, w, ch = crop.shape
        plan = ops.plan_from_tiles(h, w, [(0, 0, h, w)], device=self.device)
        packed = torch.from_numpy(crop.reshape(-1)).to(self.device)
        b, c, f, _ = self.predict_tiles(packed, plan, ch, conf)

        class _Det:
            def __init__(self, bb, cc, ff):
                self.xyxyxyxy, self.cls, self.conf = bb.reshape(1, 4, 2), cc.reshape(1).float(), ff.reshape(1)

        class _Res:
            pass

        r = _Res()
        r.obb = [_Det(b[i], c[i], f[i]) for i in range(b.shape[0])]
        return [r]
