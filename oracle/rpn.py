"""Oracle: Faster R-CNN RPN proposal filter.  TEST INFRASTRUCTURE ONLY.

Restates demos/faster_rcnn/models/rpn.py: ``make_anchors_xywh`` :160-166 (integer cell xy, base
anchors in feature units), ``dxdydwdh2xywh`` :111-119 (h uses exp(dw) -- sic, :117),
``filter_proposals`` :168-208 (softmax fg score, xyxy, clamp to [0,W-1]/[0,H-1], per image
topk -> nms(thr) -> first post_n -> back to xywh).  ``base_anchors`` here are already divided by
the stride (rpn.py:87).  Anchor shapes: demos/faster_rcnn/utils/anchor_generator.py:4-14.
"""
import math

import numpy as np
import torch

from .nms import _nms


def get_base_anchor(scales, ratios):
    """demos/faster_rcnn/utils/anchor_generator.py:4-14 -> float32 [len(ratios)*len(scales), 2] (w,h) px."""
    out = []
    for r in ratios:
        for s in scales:
            w = math.sqrt(s ** 2 / r)
            out.append((w, s ** 2 / w))
    return np.array(out, dtype=np.float32).reshape([-1, 2])


def make_anchors_xywh(base_anchors_feat, fh, fw):
    """rpn.py:160-166 -> [1,fh,fw,A,4] = (x=col, y=row, w, h)."""
    a = base_anchors_feat.shape[0]
    wh = base_anchors_feat.repeat(1, fh, fw, 1, 1)
    ys, xs = torch.meshgrid(torch.arange(fh), torch.arange(fw), indexing="ij")
    xy = torch.stack([xs, ys], dim=-1).unsqueeze(0).unsqueeze(3).expand(1, fh, fw, a, 2)
    return torch.cat([xy.to(wh.dtype), wh], dim=4)


def filter_proposals(cls, reg, base_anchors_feat, pre_n=2000, post_n=2000, thr=0.7, backend="numpy"):
    """rpn.py:168-208.  cls[B,H,W,A,2], reg[B,H,W,A,4] -> list of [k_i,4] xywh (feature units)."""
    bsz, fh, fw, na, _ = reg.shape
    anc = make_anchors_xywh(base_anchors_feat, fh, fw)
    xywh = reg.clone()
    xywh[..., 0] = reg[..., 0] * anc[..., 2] + anc[..., 0]
    xywh[..., 1] = reg[..., 1] * anc[..., 3] + anc[..., 1]
    xywh[..., 2] = torch.exp(reg[..., 2]) * anc[..., 2]
    xywh[..., 3] = torch.exp(reg[..., 2]) * anc[..., 3]          # :117 exp(dw), not exp(dh)
    score = torch.softmax(cls.clone(), dim=4)[..., 1]
    prop = torch.cat([score[..., None], xywh], dim=4).detach().view(bsz, -1, 5)
    half_w, half_h = prop[..., 3] / 2, prop[..., 4] / 2
    x1 = (prop[..., 1] - half_w).clamp(min=0, max=fw - 1)
    y1 = (prop[..., 2] - half_h).clamp(min=0, max=fh - 1)
    x2 = (prop[..., 1] + half_w).clamp(min=0, max=fw - 1)
    y2 = (prop[..., 2] + half_h).clamp(min=0, max=fh - 1)
    xyxy = torch.stack([x1, y1, x2, y2], dim=-1)
    out = []
    for b in range(bsz):
        k = min(pre_n, xyxy.size(1))
        _, top = prop[b, :, 0].topk(k, dim=-1)
        bx, sc = xyxy[b, top], prop[b, top, 0]
        keep = _nms(bx, sc, thr, backend)
        keep = keep[:min(post_n, keep.size(0))]
        kb = bx[keep]
        out.append(torch.stack([(kb[:, 0] + kb[:, 2]) / 2, (kb[:, 1] + kb[:, 3]) / 2,
                                kb[:, 2] - kb[:, 0], kb[:, 3] - kb[:, 1]], dim=1))
    return out
