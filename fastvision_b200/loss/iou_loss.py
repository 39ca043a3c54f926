"""IoU-family losses -- drop-in for loss/iou_loss.py:5-107: ``1 - kind(y_pre, y_true)`` [* weights], mean or sum.

One fused IoU+reduction kernel plus a fixed-order fp64 finish (forward only; the backward is a next-row item).
"""
import torch
import torch.nn as nn

from .. import _lib


class _IoUFamilyLoss(nn.Module):
    kind = "iou"

    def __init__(self, reduction='mean'):
        super().__init__()
        self.reduction = reduction

    def forward(self, y_pre, y_true, weights=None, mode='xyxy'):
        if mode not in _lib.BOX_MODES or (self.kind != "iou" and mode == "wh"):
            raise Exception('mode must be xyxy or xywh or wh')
        y_pre = _lib.require_cuda(y_pre, "y_pre")
        y_true = _lib.require_cuda(y_true, "y_true")
        n = y_pre.size(0)
        if weights is not None:
            weights = _lib.require_cuda(weights, "weights")
            if weights.numel() != n:
                raise ValueError("weights must have one entry per box")
        out = torch.empty((), dtype=torch.float32, device=y_pre.device)
        lib = _lib.load()
        ws = _lib.workspace(lib.fvb_reduce_workspace_bytes(n), y_pre.device, "reduce")
        red = _lib.REDUCTIONS["mean" if self.reduction == 'mean' else "sum"]
        with torch.cuda.device(y_pre.device):
            _lib.check(lib.fvb_iou_loss_f32(_lib.dptr(y_pre), _lib.dptr(y_true), _lib.dptr(weights), n,
                                            _lib.BOX_MODES[mode], _lib.IOU_KINDS[self.kind], 0, 1e-7, red,
                                            _lib.dptr(out), _lib.dptr(ws), _lib.stream()), "iou_loss")
        return out


class IOULoss(_IoUFamilyLoss):
    """loss/iou_loss.py:5-29."""
    kind = "iou"


class GIOULoss(_IoUFamilyLoss):
    """loss/iou_loss.py:31-55 (GIOU returns [n]; weights [n,1] broadcast to [n,n] as in the reference)."""
    kind = "giou"


class DIOULoss(_IoUFamilyLoss):
    """loss/iou_loss.py:57-81."""
    kind = "diou"


class CIOULoss(_IoUFamilyLoss):
    """loss/iou_loss.py:83-107."""
    kind = "ciou"
