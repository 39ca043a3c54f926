"""Binary cross-entropy -- drop-in for BiCrossEntropyLoss, loss/classification_loss.py:36-65 (forward + backward w.r.t. y_pre)."""
import torch
import torch.nn as nn

from .. import _lib


def _forward(y_pre, tidx, tval, weights, rows, classes, sig, red):
    out = torch.empty((), dtype=torch.float32, device=y_pre.device)
    lib = _lib.load()
    ws = _lib.workspace(lib.fvb_reduce_workspace_bytes(y_pre.numel()), y_pre.device, "reduce")
    with torch.cuda.device(y_pre.device):
        _lib.check(lib.fvb_bce_loss_f32(_lib.dptr(y_pre), rows, classes, _lib.dptr(tidx), _lib.dptr(tval), sig,
                                        _lib.dptr(weights), _lib.REDUCTIONS[red], _lib.dptr(out), _lib.dptr(ws),
                                        _lib.stream()), "bce_loss")
    return out


class _BCEFn(torch.autograd.Function):

    @staticmethod
    def forward(fctx, y_pre, tidx, tval, weights, rows, classes, sig, red):
        fctx.save_for_backward(y_pre, tidx, tval, weights)
        fctx.cfg = (rows, classes, sig, red)
        return _forward(y_pre, tidx, tval, weights, rows, classes, sig, red)

    @staticmethod
    def backward(fctx, grad_out):
        y_pre, tidx, tval, weights = fctx.saved_tensors
        rows, classes, sig, red = fctx.cfg
        g = torch.empty_like(y_pre)
        grad_out = _lib.require_cuda(grad_out.detach().reshape(-1)[:1], "grad_out")
        lib = _lib.load()
        with torch.cuda.device(y_pre.device):
            _lib.check(lib.fvb_bce_loss_backward_f32(_lib.dptr(y_pre), rows, classes, _lib.dptr(tidx), _lib.dptr(tval), sig,
                                                     _lib.dptr(weights), _lib.REDUCTIONS[red], _lib.dptr(grad_out), _lib.dptr(g),
                                                     _lib.stream()), "bce_loss_backward")
        return g, None, None, None, None, None, None, None


class BiCrossEntropyLoss(nn.Module):

    def __init__(self, reduction='mean'):
        super(BiCrossEntropyLoss, self).__init__()
        self.reduction = reduction

    def forward(self, y_pre, y_true, already_sigmoid=False, weights=None):
        y_pre = _lib.require_cuda(y_pre, "y_pre")
        classes = y_pre.size(-1)
        rows = y_pre.numel() // classes
        tidx = tval = None
        if classes > 1:                                   # one_hot(y_true, C)  (classification_loss.py:44-46)
            tidx = _lib.require_cuda(y_true, "y_true", dtype=None).reshape(-1).long().contiguous()
            if tidx.numel() != rows:
                raise ValueError("y_true must hold one class index per row")
        else:                                             # C == 1: y_true itself is the target (:47-48)
            tval = _lib.require_cuda(y_true, "y_true", dtype=None).reshape(-1).float().contiguous()
            if tval.numel() != rows:
                raise ValueError("y_true must hold one target per row")
        if weights is not None:
            weights = _lib.require_cuda(weights, "weights")
            if weights.numel() != y_pre.numel():
                raise ValueError("weights must match the flattened prediction")
        red = "mean" if self.reduction == 'mean' else "sum"
        sig = 1 if already_sigmoid else 0
        if torch.is_grad_enabled() and y_pre.requires_grad:
            return _BCEFn.apply(y_pre, tidx, tval, None if weights is None else weights.detach(), rows, classes, sig, red)
        return _forward(y_pre, tidx, tval, weights, rows, classes, sig, red)
