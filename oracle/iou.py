"""Oracle: the reference's IoU family (bug-compatible).  TEST INFRASTRUCTURE ONLY.

Follows the *torch* branches of detection/tools/IOU.py (the ones loss/iou_loss.py,
loss/yolov3_loss.py and metrics/map.py reach).  Quirks kept on purpose:
  * element-wise ``xyxy_iou`` puts eps inside the height factor of both areas (IOU.py:74-75);
    the pairwise one does not (IOU.py:143-144);
  * ``GIOU`` returns shape [n] and uses ``iou - (C-U)/C`` (IOU.py:239) but ``GIOU_batch``
    returns ``iou + (C-U)/C`` (IOU.py:290);
  * ``DIOU``/``DIOU_batch`` ADD rho^2/c^2 (IOU.py:341,393), hence CIOU = iou + rho^2/c^2 - alpha*v.
``variant='demo'`` gives demos/yolov3_u/utils/iou.py:334-341,385-393 instead
(``iou - rho'^2/c^2`` with un-halved centre sums).
All arithmetic fp32, eps folded to fp32 when it meets a tensor, same association order.
"""
import math

import torch

from .boxes import xywh2xyxy


def _prep(b1, b2, mode):
    if mode == "xywh":
        return xywh2xyxy(b1), xywh2xyxy(b2)
    if mode == "xyxy":
        return b1, b2
    raise Exception("mode must be xyxy or xywh or wh")


def _cols(b):
    return b[:, 0], b[:, 1], b[:, 2], b[:, 3]


def _inter(ax1, ay1, ax2, ay2, bx1, by1, bx2, by2):
    iw = (torch.minimum(ax2, bx2) - torch.maximum(ax1, bx1)).clamp(0)
    ih = (torch.minimum(ay2, by2) - torch.maximum(ay1, by1)).clamp(0)
    return iw * ih


def xyxy_iou(a, b, eps=1e-7):
    """IOU.py:73-85 (torch branch) -> [n,1]."""
    ax1, ay1, ax2, ay2 = _cols(a)
    bx1, by1, bx2, by2 = _cols(b)
    area_a = (ax2 - ax1) * (ay2 - ay1 + eps)
    area_b = (bx2 - bx1) * (by2 - by1 + eps)
    inter = _inter(ax1, ay1, ax2, ay2, bx1, by1, bx2, by2)
    union = area_a + area_b - inter + eps
    return (inter / union).reshape([-1, 1])


def xyxy_iou_batch(a, b, eps=1e-7):
    """IOU.py:142-154 (torch branch) -> [N,M]."""
    ax1, ay1, ax2, ay2 = (c[:, None] for c in _cols(a))
    bx1, by1, bx2, by2 = _cols(b)
    area_a = (ax2 - ax1) * (ay2 - ay1)
    area_b = (bx2 - bx1) * (by2 - by1)
    inter = _inter(ax1, ay1, ax2, ay2, bx1, by1, bx2, by2)
    union = area_a + area_b - inter + eps
    return inter / union


def xywh_iou(a, b, eps=1e-7):
    """IOU.py:27-38."""
    return xyxy_iou(xywh2xyxy(a), xywh2xyxy(b), eps)


def xywh_iou_batch(a, b, eps=1e-7):
    """IOU.py:40-51."""
    return xyxy_iou_batch(xywh2xyxy(a), xywh2xyxy(b), eps)


def wh_iou(a, b, eps=1e-7):
    """IOU.py:108-120 -> [n,1]."""
    inter = torch.minimum(a[:, 0], b[:, 0]) * torch.minimum(a[:, 1], b[:, 1])
    union = a[:, 0] * a[:, 1] + b[:, 0] * b[:, 1] - inter + eps
    return (inter / union).reshape([-1, 1])


def wh_iou_batch(a, b, eps=1e-7):
    """IOU.py:177-189 -> [N,M]."""
    inter = torch.minimum(a[:, None, 0], b[:, 0]) * torch.minimum(a[:, None, 1], b[:, 1])
    union = (a[:, 0] * a[:, 1])[:, None] + b[:, 0] * b[:, 1] - inter + eps
    return inter / union


def cal_iou(b1, b2, mode="xyxy", eps=1e-7):
    """IOU.py:7-15."""
    if mode == "xyxy":
        return xyxy_iou(b1, b2, eps)
    if mode == "xywh":
        return xywh_iou(b1, b2, eps)
    if mode == "wh":
        return wh_iou(b1, b2, eps)
    raise Exception("mode must be xyxy or xywh or wh")


def cal_iou_batch(b1, b2, mode="xyxy", eps=1e-7):
    """IOU.py:17-25."""
    if mode == "xyxy":
        return xyxy_iou_batch(b1, b2, eps)
    if mode == "xywh":
        return xywh_iou_batch(b1, b2, eps)
    if mode == "wh":
        return wh_iou_batch(b1, b2, eps)
    raise Exception("mode must be xyxy or xywh or wh")


def _giou_terms(a_cols, b_cols, eps):
    ax1, ay1, ax2, ay2 = a_cols
    bx1, by1, bx2, by2 = b_cols
    area_a = (ax2 - ax1) * (ay2 - ay1)
    area_b = (bx2 - bx1) * (by2 - by1)
    inter = _inter(ax1, ay1, ax2, ay2, bx1, by1, bx2, by2)
    union = area_a + area_b - inter + eps
    iou = inter / union
    cw = torch.maximum(ax2, bx2) - torch.minimum(ax1, bx1)
    ch = torch.maximum(ay2, by2) - torch.minimum(ay1, by1)
    convex = cw * ch + eps
    return iou, union, convex


def GIOU(b1, b2, mode="xyxy", eps=1e-7):
    """IOU.py:220-239 -> shape [n] (not [n,1]); iou - (C-U)/C."""
    a, b = _prep(b1, b2, mode)
    iou, union, convex = _giou_terms(_cols(a), _cols(b), eps)
    return iou - (convex - union) / convex


def GIOU_batch(b1, b2, mode="xyxy", eps=1e-7):
    """IOU.py:270-290 -> [N,M]; iou PLUS (C-U)/C (sic)."""
    a, b = _prep(b1, b2, mode)
    iou, union, convex = _giou_terms(tuple(c[:, None] for c in _cols(a)), _cols(b), eps)
    return iou + (convex - union) / convex


def _diou_penalty(a_cols, b_cols, eps, variant):
    ax1, ay1, ax2, ay2 = a_cols
    bx1, by1, bx2, by2 = b_cols
    cw = torch.maximum(ax2, bx2) - torch.minimum(ax1, bx1)
    ch = torch.maximum(ay2, by2) - torch.minimum(ay1, by1)
    c2 = cw ** 2 + ch ** 2 + eps
    if variant == "demo":  # demos/yolov3_u/utils/iou.py:334-341 -- centre sums not halved
        rho2 = ((ax1 + ax2) - (bx1 + bx2)) ** 2 + ((ay1 + ay2) - (by1 + by2)) ** 2
    else:
        rho2 = ((ax1 + ax2) * 0.5 - (bx1 + bx2) * 0.5) ** 2 + ((ay1 + ay2) * 0.5 - (by1 + by2) * 0.5) ** 2
    return rho2 / c2


def DIOU(b1, b2, mode="xyxy", eps=1e-7, variant="lib"):
    """IOU.py:307,324-341 -> [n,1]; lib: iou + rho^2/c^2; demo: iou - rho'^2/c^2."""
    a, b = _prep(b1, b2, mode)
    iou = xyxy_iou(a, b, eps)
    pen = _diou_penalty(_cols(a), _cols(b), eps, variant).view(-1, 1)
    return iou - pen if variant == "demo" else iou + pen


def DIOU_batch(b1, b2, mode="xyxy", eps=1e-7, variant="lib"):
    """IOU.py:358,375-393 -> [N,M]."""
    a, b = _prep(b1, b2, mode)
    iou = xyxy_iou_batch(a, b, eps)
    pen = _diou_penalty(tuple(c[:, None] for c in _cols(a)), _cols(b), eps, variant)
    return iou - pen if variant == "demo" else iou + pen


def CIOU(b1, b2, mode="xyxy", eps=1e-7, variant="lib"):
    """IOU.py:410-438 -> [n,1]; DIOU - alpha*v, alpha = v/((v-iou)+(1+eps))."""
    a, b = _prep(b1, b2, mode)
    iou = xyxy_iou(a, b, eps)
    diou = DIOU(a, b, "xyxy", eps, variant)
    w1, h1 = a[:, 2] - a[:, 0], a[:, 3] - a[:, 1]
    w2, h2 = b[:, 2] - b[:, 0], b[:, 3] - b[:, 1]
    v = (4 / math.pi ** 2) * torch.pow(torch.atan(w2 / (h2 + eps)) - torch.atan(w1 / (h1 + eps)), 2)
    v = v.view(-1, 1)
    with torch.no_grad():                                  # IOU.py:436-437: alpha is a constant for autograd
        alpha = v / (v - iou + (1 + eps))
    return diou - alpha * v


def CIOU_batch(b1, b2, mode="xyxy", eps=1e-7, variant="lib"):
    """IOU.py:455-480 -> [N,M]."""
    a, b = _prep(b1, b2, mode)
    iou = xyxy_iou_batch(a, b, eps)
    diou = DIOU_batch(a, b, "xyxy", eps, variant)
    w1, h1 = a[:, 2] - a[:, 0], a[:, 3] - a[:, 1]
    w2, h2 = b[:, 2] - b[:, 0], b[:, 3] - b[:, 1]
    v = (4 / math.pi ** 2) * torch.pow(torch.atan(w1 / (h1 + eps))[:, None] - torch.atan(w2 / (h2 + eps)), 2)
    with torch.no_grad():                                  # IOU.py:478-479
        alpha = v / (v - iou + (1 + eps))
    return diou - alpha * v


_KIND = {"iou": cal_iou, "giou": GIOU, "diou": DIOU, "ciou": CIOU}


def iou_loss(kind, y_pre, y_true, weights=None, mode="xyxy", reduction="mean"):
    """loss/iou_loss.py:5-107 -- loss = 1 - kind(...); * weights; mean or sum.

    (GIOU's [n] result times [n,1] weights broadcasts to [n,n] in the reference; kept.)
    """
    val = _KIND[kind](y_pre, y_true, mode=mode)
    loss = 1 - val
    if weights is None:
        weights = torch.ones_like(loss)
    loss = loss * weights
    return torch.mean(loss) if reduction == "mean" else torch.sum(loss)
