"""IoU-family losses -- drop-in for loss/iou_loss.py:5-107: ``1 - kind(y_pre, y_true)`` [* weights], mean or sum.

One fused IoU+reduction kernel plus a fixed-order fp64 finish; when an input requires grad the call goes through
``_IoULossFn`` whose backward is one element-wise kernel (``fvb_iou_loss_backward_f32``, torch's tie / clamp conventions).
"""
import torch
import torch.nn as nn

from .. import _lib


def _forward(y_pre, y_true, weights, mode, kind, reduction):
    n = y_pre.size(0)
    out = torch.empty((), dtype=torch.float32, device=y_pre.device)
    lib = _lib.load()
    ws = _lib.workspace(lib.fvb_reduce_workspace_bytes(n), y_pre.device, "reduce")
    with torch.cuda.device(y_pre.device):
        _lib.check(lib.fvb_iou_loss_f32(_lib.dptr(y_pre), _lib.dptr(y_true), _lib.dptr(weights), n,
                                        _lib.BOX_MODES[mode], _lib.IOU_KINDS[kind], 0, 1e-7, _lib.REDUCTIONS[reduction],
                                        _lib.dptr(out), _lib.dptr(ws), _lib.stream()), "iou_loss")
    return out


class _IoULossFn(torch.autograd.Function):

    @staticmethod
    def forward(fctx, y_pre, y_true, weights, mode, kind, reduction):
        fctx.save_for_backward(y_pre, y_true, weights)
        fctx.cfg = (mode, kind, reduction)
        return _forward(y_pre, y_true, weights, mode, kind, reduction)

    @staticmethod
    def backward(fctx, grad_out):
        y_pre, y_true, weights = fctx.saved_tensors
        mode, kind, reduction = fctx.cfg
        g_pre = torch.empty_like(y_pre) if fctx.needs_input_grad[0] else None
        g_true = torch.empty_like(y_true) if fctx.needs_input_grad[1] else None
        if g_pre is None and g_true is None:
            return None, None, None, None, None, None
        grad_out = _lib.require_cuda(grad_out.detach().reshape(-1)[:1], "grad_out")
        lib = _lib.load()
        ws = _lib.workspace(256, y_pre.device, "iou_loss_bwd")
        with torch.cuda.device(y_pre.device):
            _lib.check(lib.fvb_iou_loss_backward_f32(_lib.dptr(y_pre), _lib.dptr(y_true), _lib.dptr(weights), y_pre.size(0),
                                                     _lib.BOX_MODES[mode], _lib.IOU_KINDS[kind], 0, 1e-7,
                                                     _lib.REDUCTIONS[reduction], _lib.dptr(grad_out), _lib.dptr(g_pre),
                                                     _lib.dptr(g_true), _lib.dptr(ws), _lib.stream()), "iou_loss_backward")
        return g_pre, g_true, None, None, None, None


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
        red = "mean" if self.reduction == 'mean' else "sum"
        if torch.is_grad_enabled() and (y_pre.requires_grad or y_true.requires_grad):
            return _IoULossFn.apply(y_pre, y_true, None if weights is None else weights.detach(), mode, self.kind, red)
        return _forward(y_pre, y_true, weights, mode, self.kind, red)


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
