"""The demos' training loss ``ComputeLoss`` -- drop-in for demos/yolov3_huaweiShip/utils/lossv3.py:7-125 (flavour "ship":
returns (loss_box, loss_cls, loss_conf)) and demos/yolov3_u/utils/lossv3.py:7-119 (flavour "u": returns the scalar
2*xy + wh + cls + conf), forward and backward.

``forward(predict_layers, target_all, model)`` keeps the reference's signature: ``predict_layers`` are the raw conv
outputs [B, A*K, H, W], ``target_all`` is [T,6] = [batch_idx, cls, x, y, w, h] (normalised), ``model.anchors`` (or
``model.module.anchors``) the per-level [A,2] anchors in feature units.  The reference's Python loop over images
(pairwise IoU of every predicted box with the image's targets, lossv3.py:102-113) and its ~60 ATen launches per level
become 4 launches for all levels; the backward is 3 launches.

Deviations (documented in include/fvb200.h): the reference's debug ``print`` of the components (yolov3_u) is dropped;
with ``strict=True`` (default) an image without targets raises IndexError like lossv3.py:107 does (one host sync);
``strict=False`` treats it as an image without ignore region and never synchronises.
"""
import torch
import torch.nn as nn

from .. import _lib


class _Ctx:
    def __init__(self, heads, anchors):
        b = int(heads[0].size(0))
        na = int(anchors[0].reshape(-1, 2).size(0))
        ak = int(heads[0].size(1))
        if ak % na:
            raise ValueError("head channels %d not divisible by %d anchors" % (ak, na))
        self.key = (b, ak, tuple((int(h.size(2)), int(h.size(3))) for h in heads), heads[0].device)
        self.geom = _lib.make_geom(b, ak // na, [int(h.size(2)) for h in heads], [int(h.size(3)) for h in heads],
                                   [1.0] * len(heads), [a.detach().float().cpu().reshape(-1, 2) for a in anchors], "nchw")
        self.device = heads[0].device
        self.levels = len(heads)
        lib = _lib.load()
        self.mask_bytes = int(lib.fvb_demo_loss_mask_bytes(self.geom))


class _DemoLossFn(torch.autograd.Function):

    @staticmethod
    def forward(fctx, module, labels, ctx, *heads):
        out, partials, mask = module._run(list(heads), labels, ctx, want_mask=True)
        fctx.module, fctx.labels, fctx.ctx, fctx.partials, fctx.mask = module, labels, ctx, partials, mask
        fctx.save_for_backward(*heads)
        return out

    @staticmethod
    def backward(fctx, grad_out):
        grads = fctx.module.backward_heads(list(fctx.saved_tensors), fctx.labels, grad_out, fctx.partials, fctx.mask, ctx=fctx.ctx)
        return (None, None, None) + tuple(grads)


class ComputeLoss(nn.Module):
    flavour = "ship"

    def __init__(self, strict=True):
        super(ComputeLoss, self).__init__()
        self.strict = strict
        self._ctx = None
        self.partials = None      # [L,6] f64 {S_box|S_xy, S_wh, S_cls, S_conf, n_valid, T}: what data-parallel ranks all-reduce

    def get_model(self, model):
        return model.module if hasattr(model, 'module') else model

    def _context(self, heads, anchors):
        key = (int(heads[0].size(0)), int(heads[0].size(1)), tuple((int(h.size(2)), int(h.size(3))) for h in heads), heads[0].device)
        if self._ctx is None or self._ctx.key != key:
            self._ctx = _Ctx(heads, anchors)
        return self._ctx

    def _run(self, heads, labels, ctx, want_mask=False):
        lib = _lib.load()
        dev = ctx.device
        t = labels.size(0)
        out = torch.empty(3, dtype=torch.float32, device=dev)
        partials = torch.empty(ctx.levels, 6, dtype=torch.float64, device=dev)
        mask = torch.empty(ctx.mask_bytes, dtype=torch.int8, device=dev) if want_mask else None
        ws = _lib.workspace(lib.fvb_demo_loss_workspace_bytes(ctx.geom, t), dev, "demo_loss")
        with torch.cuda.device(dev):
            _lib.check(lib.fvb_demo_loss_f32(ctx.geom, _lib.head_ptrs(heads), _lib.dptr(labels), t,
                                             _lib.DEMO_LOSS_FLAVOURS[self.flavour], _lib.dptr(partials), _lib.dptr(out),
                                             _lib.dptr(mask), _lib.dptr(ws), _lib.stream()), "demo_loss")
        self.partials = partials
        return out, partials, mask

    def forward(self, predict_layers, target_all, model):
        heads = [_lib.require_cuda(h, "predict_layers[%d]" % i) for i, h in enumerate(predict_layers)]
        labels = _lib.require_cuda(target_all, "target_all").view(-1, 6)
        anchors = self.get_model(model).anchors
        ctx = self._context(heads, anchors)
        if self.strict:
            counts = torch.bincount(labels[:, 0].long(), minlength=ctx.geom.batch) if labels.size(0) else \
                torch.zeros(ctx.geom.batch, dtype=torch.long, device=labels.device)
            if int(counts.min()) == 0:      # lossv3.py:107: torch.max over an empty [n,0] IoU matrix
                raise IndexError("max(): Expected reduction dim 1 to have non-zero size.")
        if torch.is_grad_enabled() and any(h.requires_grad for h in heads):
            out = _DemoLossFn.apply(self, labels.detach(), ctx, *heads)
        else:
            out = self._run(heads, labels, ctx)[0]
        return self._shape(out)

    def _shape(self, out):
        return out[0:1], out[1:2], out[2:3]          # three Tensor[1], lossv3.py:125

    def combine(self, partials, ctx=None):
        """Outputs from (all-reduced) per-level partials."""
        ctx = ctx or self._ctx
        out = torch.empty(3, dtype=torch.float32, device=partials.device)
        lib = _lib.load()
        with torch.cuda.device(partials.device):
            _lib.check(lib.fvb_demo_loss_combine_f32(ctx.geom, _lib.dptr(partials), _lib.DEMO_LOSS_FLAVOURS[self.flavour],
                                                     _lib.dptr(out), _lib.stream()), "demo_loss_combine")
        return self._shape(out)

    def backward_heads(self, predict_layers, target_all, grad_out, partials, mask, ctx=None, grads=None, anchors=None):
        """Gradients w.r.t. the conv outputs; ``grad_out``: Tensor[3] (ship) / Tensor[1] (u) upstream gradients or None."""
        heads = [_lib.require_cuda(h.detach(), "predict_layers[%d]" % i) for i, h in enumerate(predict_layers)]
        labels = _lib.require_cuda(target_all, "target_all").view(-1, 6)
        ctx = ctx or self._context(heads, anchors)
        if grads is None:
            grads = [torch.empty_like(h) for h in heads]
        if grad_out is not None:
            grad_out = _lib.require_cuda(grad_out.detach().reshape(-1), "grad_out")
        lib = _lib.load()
        ws = _lib.workspace(lib.fvb_demo_loss_workspace_bytes(ctx.geom, labels.size(0)), ctx.device, "demo_loss_bwd")
        with torch.cuda.device(ctx.device):
            _lib.check(lib.fvb_demo_loss_backward_f32(ctx.geom, _lib.head_ptrs(heads), _lib.dptr(labels), labels.size(0),
                                                      _lib.DEMO_LOSS_FLAVOURS[self.flavour], _lib.dptr(partials),
                                                      _lib.dptr(mask), _lib.dptr(grad_out), _lib.head_ptrs(grads),
                                                      _lib.dptr(ws), _lib.stream()), "demo_loss_backward")
        return grads


class ComputeLossU(ComputeLoss):
    """demos/yolov3_u/utils/lossv3.py:7-119 -> Tensor[1] = 2*loss_xy + loss_wh + loss_cls + loss_conf."""
    flavour = "u"

    def _shape(self, out):
        return out[0:1]
