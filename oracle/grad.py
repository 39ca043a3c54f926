"""Oracle gradients: torch autograd through the CPU restatements.  TEST INFRASTRUCTURE ONLY.

The reference trains with ``loss.backward()`` (utils/fit.py:57-63), i.e. the gradients are whatever torch autograd
derives from loss/yolov3_loss.py:29-72, loss/iou_loss.py:5-107 and loss/classification_loss.py:36-65.  The oracle
restatements use the same differentiable torch ops in the same order (alpha under no_grad, targets_conf not
detached, index_put with duplicates), so autograd through them is the gradient oracle; it is pinned against gradients
recorded from the reference itself (tests/golden/grads_small.npz, made by oracle/make_golden.py:gold_grads).
"""
import torch

from . import iou as _iou
from . import loss as _loss


def yolov3_loss_grad(heads, labels, anchors_per_level, strides, ratio_box=0.05, ratio_conf=1.0, ratio_cls=0.5,
                     upstream=1.0):
    """-> (loss[1], [d loss*upstream / d head_l])."""
    hs = [h.detach().clone().requires_grad_(True) for h in heads]
    out = _loss.yolov3_loss(hs, labels, anchors_per_level, strides, ratio_box, ratio_conf, ratio_cls)
    (out * upstream).sum().backward()
    return out.detach(), [h.grad for h in hs]


def iou_loss_grad(kind, y_pre, y_true, weights=None, mode="xyxy", reduction="mean", upstream=1.0):
    a = y_pre.detach().clone().requires_grad_(True)
    b = y_true.detach().clone().requires_grad_(True)
    out = _iou.iou_loss(kind, a, b, weights, mode, reduction)
    (out * upstream).backward()
    return out.detach(), a.grad, b.grad


def bce_loss_grad(y_pre, y_true, already_sigmoid=False, weights=None, reduction="mean", upstream=1.0):
    x = y_pre.detach().clone().requires_grad_(True)
    out = _loss.bi_cross_entropy(x, y_true, already_sigmoid, weights, reduction)
    (out * upstream).backward()
    return out.detach(), x.grad
