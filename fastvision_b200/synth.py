"""Seeded synthetic inputs for the detection hot path.

Shared by tests, bench.py and the CPU-baseline leg so the oracle and the kernels see the same
tensors.  Pure torch-CPU generation (then moved by the caller); no oracle import.

Shapes follow the reference: head tensors ``[B, A, H, W, 5+C]`` per level as produced by
detection/head/yolov3head.py:63, labels ``[T, 6] = [batch_idx, cls, xc, yc, w, h]`` normalised and
grouped by image as datasets/detection_dataloader.py:98-103 (collate_fn) produces them.
"""
import math
from dataclasses import dataclass, field
from typing import List

import torch

SEED_BASE = 20220504  # the reference's --seed default (demos/*/run.py)


@dataclass
class YoloConfig:
    name: str
    img: int
    num_classes: int
    anchors_px: List[List[float]]          # 9 anchors, (w, h) pixels, level order stride 32 -> 16 -> 8
    strides: List[int] = field(default_factory=lambda: [32, 16, 8])
    anchors_per_level: int = 3
    labels_per_img: float = 7.3
    max_labels: int = 40

    @property
    def levels(self):
        return len(self.strides)

    @property
    def k(self):
        return 5 + self.num_classes

    @property
    def feat(self):
        return [self.img // s for s in self.strides]

    @property
    def cells(self):
        return sum(self.anchors_per_level * f * f for f in self.feat)

    def anchors_levels(self):
        """List of [A,1,1,2] pixel-unit tensors, the layout detection/models/yolov3.py:11-17 builds."""
        a = torch.tensor(self.anchors_px, dtype=torch.float32).view(self.levels, self.anchors_per_level, 1, 1, 2)
        return [a[i] for i in range(self.levels)]


# demos/yolov3_u/train.py:60-62 (COCO anchors, px)
COCO416 = YoloConfig("yolov3-416-coco", 416, 80,
                     [[116, 90], [156, 198], [373, 326], [30, 61], [62, 45], [59, 119], [10, 13], [16, 30], [33, 23]])
# demos/yolov3_huaweiShip/train.py:60-62 (anchors / 2, px) and data/v1.yaml:5 (10 classes)
SHIP608 = YoloConfig("yolov3-608-ship", 608, 10,
                     [[231.0, 101.0], [170.0, 215.5], [261.0, 182.0], [56.5, 135.0], [115.0, 67.0], [98.5, 198.5],
                      [7.5, 12.0], [23.0, 31.5], [39.5, 67.0]], labels_per_img=3.0)

CONFIGS = {c.name: c for c in (COCO416, SHIP608)}


def make_generator(config_id: int, rank: int = 0) -> torch.Generator:
    g = torch.Generator()
    g.manual_seed(SEED_BASE + 1000 * config_id + rank)
    return g


def make_labels(cfg: YoloConfig, batch: int, gen: torch.Generator) -> torch.Tensor:
    """[T,6] rows [batch_idx, cls, xc, yc, w, h], normalised, sorted by batch_idx."""
    lam = torch.full((batch,), float(cfg.labels_per_img))
    cnt = torch.poisson(lam, generator=gen).clamp_(1, cfg.max_labels).long()
    total = int(cnt.sum())
    bidx = torch.repeat_interleave(torch.arange(batch), cnt).float()
    cls = torch.randint(0, cfg.num_classes, (total,), generator=gen).float()
    ctr = torch.rand(total, 2, generator=gen) * 0.9 + 0.05
    lo, hi = math.log(0.03), math.log(0.6)
    wh = torch.exp(torch.rand(total, 2, generator=gen) * (hi - lo) + lo)
    # clip inside the image: shrink to fit around the centre
    wh = torch.minimum(wh, 2 * torch.minimum(ctr, 1 - ctr))
    return torch.cat([bidx[:, None], cls[:, None], ctr, wh], dim=1).contiguous()


def make_heads(cfg: YoloConfig, batch: int, labels: torch.Tensor, gen: torch.Generator,
               stress: bool = False, plant: bool = True) -> List[torch.Tensor]:
    """Three raw head tensors [B,A,H,W,K] fp32 with planted detections around every label."""
    heads = []
    k = cfg.k
    anchors = torch.tensor(cfg.anchors_px, dtype=torch.float32).view(cfg.levels, cfg.anchors_per_level, 2)
    for lvl, (f, s) in enumerate(zip(cfg.feat, cfg.strides)):
        a = cfg.anchors_per_level
        t = torch.empty(batch, a, f, f, k, dtype=torch.float32)
        t[..., 0:2].copy_(torch.randn(batch, a, f, f, 2, generator=gen))
        t[..., 2:4].copy_(torch.randn(batch, a, f, f, 2, generator=gen) * 0.5)
        if stress:
            t[..., 4].copy_(torch.randn(batch, a, f, f, generator=gen))
        else:
            t[..., 4].copy_(torch.randn(batch, a, f, f, generator=gen) * 2.5 - 5.0)
        t[..., 5:].copy_(torch.randn(batch, a, f, f, k - 5, generator=gen) * 1.5 - 3.0)
        if plant and labels.numel():
            _plant(t, labels, anchors[lvl], f, s, cfg.img, gen)
        heads.append(t)
    return heads


def _plant(t, labels, anchors_px, f, stride, img, gen):
    """Overwrite the responsible cell and its 4-neighbours for each (label, matching anchor)."""
    bidx = labels[:, 0].long()
    cls = labels[:, 1].long()
    box_px = labels[:, 2:6] * img                                  # cx, cy, w, h in pixels
    ratio = box_px[:, None, 2:4] / anchors_px[None]                # [T,A,2]
    ok = torch.maximum(ratio, 1 / ratio).max(2)[0] < 4             # same test as loss/yolov3_loss.py:98-99
    t_idx, a_idx = torch.nonzero(ok, as_tuple=True)
    if t_idx.numel() == 0:
        return
    nbr = torch.tensor([[0, 0], [1, 0], [-1, 0], [0, 1], [0, -1]])
    m = t_idx.numel()
    t_idx = t_idx.repeat_interleave(5)
    a_idx = a_idx.repeat_interleave(5)
    d = nbr.repeat(m, 1)
    bx = box_px[t_idx] * (1 + 0.05 * torch.randn(t_idx.numel(), 4, generator=gen))
    g0 = torch.floor(box_px[t_idx, 0:2] / stride).long()
    g = g0 + d
    inside = (g[:, 0] >= 0) & (g[:, 0] < f) & (g[:, 1] >= 0) & (g[:, 1] < f)
    t_idx, a_idx, bx, g = t_idx[inside], a_idx[inside], bx[inside], g[inside]
    frac = (bx[:, 0:2] / stride - g.float()).clamp(0.02, 0.98)
    txy = torch.log(frac / (1 - frac))
    twh = torch.log(bx[:, 2:4].clamp_min(1.0) / anchors_px[a_idx])
    n = t_idx.numel()
    obj = torch.randn(n, generator=gen) + 2.0
    cl = torch.randn(n, generator=gen) + 3.0
    b = bidx[t_idx]
    t[b, a_idx, g[:, 1], g[:, 0], 0:2] = txy
    t[b, a_idx, g[:, 1], g[:, 0], 2:4] = twh
    t[b, a_idx, g[:, 1], g[:, 0], 4] = obj
    t[b, a_idx, g[:, 1], g[:, 0], 5 + cls[t_idx]] = cl


def make_rpn_inputs(batch: int, fh: int, fw: int, num_anchors: int, gen: torch.Generator):
    """cls[B,H,W,A,2] ~ N(0,1), reg[B,H,W,A,4] ~ N(0,0.25)."""
    cls = torch.randn(batch, fh, fw, num_anchors, 2, generator=gen)
    reg = torch.randn(batch, fh, fw, num_anchors, 4, generator=gen) * 0.25
    return cls, reg


def labels_to_pixel_targets(labels: torch.Tensor, img_idx: int, img_h: int, img_w: int) -> torch.Tensor:
    """Targets of one image as [N,5] = [cls, x1, y1, x2, y2] pixels, the form utils/fit.py:98-99 feeds the mAP."""
    t = labels[labels[:, 0] == img_idx, 1:].clone()
    half = t[:, 3:5] / 2
    xyxy = torch.cat([t[:, 1:3] - half, t[:, 1:3] + half], dim=1)
    scale = torch.tensor([img_w, img_h, img_w, img_h], dtype=t.dtype)
    t[:, 1:] = xyxy * scale
    return t
