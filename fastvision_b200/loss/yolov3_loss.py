"""YOLOv3 loss -- drop-in for Yolov3Loss, loss/yolov3_loss.py:8-124, forward and backward.

``forward(y_pred, y_true)`` returns Tensor[1] exactly like the reference; the ~200 ATen launches and
>=6 host syncs collapse into 3-4 small kernels.  ``partials`` ([L,4] f64: S_cls, S_box, S_conf, M per
level) is what a data-parallel run all-reduces; ``combine`` turns reduced partials into the
scalar with the global-batch normalisers.

When a head tensor requires grad the call goes through ``_Yolov3LossFn``: ``loss.backward()`` (utils/fit.py:57-63) then runs
``fvb_yolov3_loss_backward_f32`` -- one streaming kernel that writes the whole gradient of the three head tensors plus one
small kernel for the matched rows -- instead of autograd's ~400 backward launches through the reference graph.
"""
import torch
import torch.nn as nn

from .. import _lib
from ..detection.models.yolov3 import DecodeContext


class _Yolov3LossFn(torch.autograd.Function):
    """autograd bridge: forward = fvb_yolov3_loss_f32, backward = fvb_yolov3_loss_backward_f32."""

    @staticmethod
    def forward(fctx, module, labels, geom_ctx, conf_bce0, out, partials, batch_global, *heads):
        saved = torch.empty(int(_lib.load().fvb_yolov3_saved_conf_floats(geom_ctx.geom)), dtype=torch.float32, device=geom_ctx.device)
        loss = module._forward_impl(list(heads), labels, geom_ctx, conf_bce0, out, partials, saved_conf=saved)
        fctx.module, fctx.geom_ctx, fctx.labels, fctx.saved_conf = module, geom_ctx, labels, saved
        fctx.partials, fctx.batch_global = module.partials, batch_global
        fctx.save_for_backward(*heads)
        return loss

    @staticmethod
    def backward(fctx, grad_out):
        heads = fctx.saved_tensors
        grads = fctx.module.backward_heads(list(heads), fctx.labels, grad_out, fctx.partials, fctx.batch_global, ctx=fctx.geom_ctx,
                                           saved_conf=fctx.saved_conf)
        return (None, None, None, None, None, None, None) + tuple(grads)


class Yolov3Loss(nn.Module):

    def __init__(self, model, iou_negative_thres, ratio_box, ratio_conf, ratio_cls):
        super(Yolov3Loss, self).__init__()
        if isinstance(model, torch.nn.DataParallel):
            model = model.module
        self.anchor_levels = model.anchors_per_level
        self.backbone_stride_levels = model.backbone_strides_per_level
        self.levels = len(self.backbone_stride_levels)
        self.iou_negative_thres = iou_negative_thres       # stored, never used -- as in the reference
        self.ratio_box = ratio_box
        self.ratio_conf = ratio_conf
        self.ratio_cls = ratio_cls
        self._ctx = None
        self.partials = None
        self._ws = _lib.Workspaces()      # owned: a captured graph / side stream of this module keeps using its pointers

    def _context(self, y_pred):
        key = (y_pred[0].size(0), y_pred[0].size(1), y_pred[0].size(4), tuple(int(h.size(2)) for h in y_pred),
               tuple(int(h.size(3)) for h in y_pred), y_pred[0].device)
        if self._ctx is None or self._ctx.key != key:
            self._ctx = DecodeContext(y_pred, self.anchor_levels, self.backbone_stride_levels)
        return self._ctx

    def match(self, y_pred, y_true, ctx, conf_bce0_precise=False):
        """First part of the two-part inference form (fvb_yolov3_loss_match_f32): target assignment and matched-row terms from
        the RAW heads and the labels, enqueued on the current stream -- a caller that also decodes these heads (ValStep) runs
        it beside the decode.  ``finish`` completes the loss once the decode's objectness partials exist."""
        heads = [_lib.require_cuda(h.detach(), "y_pred[%d]" % i) for i, h in enumerate(y_pred)]
        labels = _lib.require_cuda(y_true, "y_true").view(-1, 6)
        lib = _lib.load()
        ws = self._ws.get("yolov3_loss", lib.fvb_yolov3_loss_workspace_bytes(ctx.geom, labels.size(0)), ctx.device)
        with torch.cuda.device(ctx.device):
            _lib.check(lib.fvb_yolov3_loss_match_dense_f32(ctx.geom, _lib.head_ptrs(heads), _lib.dptr(labels), labels.size(0),
                                                           1 if conf_bce0_precise else 0, _lib.dptr(ws), _lib.stream()),
                       "yolov3_loss_match")

    def finish(self, num_labels, ctx, conf_bce0, out=None, partials=None):
        """Second part (fvb_yolov3_loss_finish_f32), after ``match`` and the decode on the same stream order."""
        dev = ctx.device
        if out is None:
            out = torch.empty(1, dtype=torch.float32, device=dev)
        if partials is None:
            partials = torch.empty(ctx.geom.levels, 4, dtype=torch.float64, device=dev)
        lib = _lib.load()
        ws = self._ws.get("yolov3_loss", lib.fvb_yolov3_loss_workspace_bytes(ctx.geom, int(num_labels)), dev)
        with torch.cuda.device(dev):
            _lib.check(lib.fvb_yolov3_loss_finish_f32(ctx.geom, int(num_labels), float(self.ratio_box), float(self.ratio_conf),
                                                      float(self.ratio_cls), _lib.dptr(conf_bce0), _lib.dptr(partials),
                                                      _lib.dptr(out), _lib.dptr(ws), _lib.stream()), "yolov3_loss_finish")
        self.partials = partials
        return out

    def forward(self, y_pred, y_true, conf_bce0=None, ctx=None, out=None, partials=None, conf_bce0_precise=False):
        """y_pred: list of raw [B,A,H,W,K]; y_true [T,6] = [batch_idx, cls, xc, yc, w, h] -> Tensor[1].

        ``conf_bce0``: partials written by ``yolov3_decode(..., want_bce0=True)`` over the same heads; when
        given the loss does not touch the dense objectness channel again (``conf_bce0_precise``: that decode ran with
        ``precise=True``).
        """
        heads = [_lib.require_cuda(h, "y_pred[%d]" % i) for i, h in enumerate(y_pred)]
        labels = _lib.require_cuda(y_true, "y_true").view(-1, 6)
        ctx = ctx or self._context(heads)
        if torch.is_grad_enabled() and any(h.requires_grad for h in heads):
            return _Yolov3LossFn.apply(self, labels.detach(), ctx, conf_bce0, out, partials, None, *heads)
        return self._forward_impl(heads, labels, ctx, conf_bce0, out, partials, conf_bce0_precise=conf_bce0_precise)

    def _forward_impl(self, heads, labels, ctx, conf_bce0, out, partials, saved_conf=None, conf_bce0_precise=False):
        dev = ctx.device
        t = labels.size(0)
        if out is None:
            out = torch.empty(1, dtype=torch.float32, device=dev)
        if partials is None:
            partials = torch.empty(ctx.geom.levels, 4, dtype=torch.float64, device=dev)
        lib = _lib.load()
        ws = self._ws.get("yolov3_loss", lib.fvb_yolov3_loss_workspace_bytes(ctx.geom, t), dev)
        with torch.cuda.device(dev):
            if conf_bce0 is not None and conf_bce0_precise and saved_conf is None:
                _lib.check(lib.fvb_yolov3_loss_dense_f32(ctx.geom, _lib.head_ptrs(heads), _lib.dptr(labels), t,
                                                         float(self.ratio_box), float(self.ratio_conf), float(self.ratio_cls),
                                                         _lib.dptr(conf_bce0), 1, _lib.dptr(partials), _lib.dptr(out),
                                                         _lib.dptr(ws), _lib.stream()), "yolov3_loss")
                self.partials = partials
                return out
            _lib.check(lib.fvb_yolov3_loss_train_f32(ctx.geom, _lib.head_ptrs(heads), _lib.dptr(labels), t,
                                                     float(self.ratio_box), float(self.ratio_conf), float(self.ratio_cls),
                                                     _lib.dptr(conf_bce0), _lib.dptr(partials), _lib.dptr(out),
                                                     _lib.dptr(saved_conf), _lib.dptr(ws), _lib.stream()), "yolov3_loss")
        self.partials = partials
        return out

    def backward_heads(self, y_pred, y_true, grad_out=None, partials=None, batch_global=None, ctx=None, grads=None,
                       saved_conf=None):
        """Gradients of ``forward`` w.r.t. the raw head tensors (list of [B,A,H,W,K], written completely).

        ``partials``: the forward's [L,4] partials (all-reduced under data parallelism, together with ``batch_global``);
        ``grad_out``: the upstream gradient (Tensor[1] on the device) or None for 1;
        ``saved_conf``: the compact objectness logits the training forward wrote (None: strided re-read of the heads).
        """
        heads = [_lib.require_cuda(h.detach(), "y_pred[%d]" % i) for i, h in enumerate(y_pred)]
        labels = _lib.require_cuda(y_true, "y_true").view(-1, 6)
        ctx = ctx or self._context(heads)
        dev = ctx.device
        partials = self.partials if partials is None else partials
        if partials is None:
            raise RuntimeError("backward_heads needs the forward's partials (call forward first)")
        if grads is None:
            grads = [torch.empty_like(h) for h in heads]
        if grad_out is not None:
            grad_out = _lib.require_cuda(grad_out.detach().reshape(-1)[:1], "grad_out")
        lib = _lib.load()
        ws = self._ws.get("yolov3_loss_bwd", lib.fvb_yolov3_loss_backward_workspace_bytes(ctx.geom, labels.size(0)), dev)
        bg = int(batch_global) if batch_global else int(heads[0].size(0))
        with torch.cuda.device(dev):
            _lib.check(lib.fvb_yolov3_loss_backward_f32(ctx.geom, _lib.head_ptrs(heads), _lib.dptr(labels), labels.size(0),
                                                        float(self.ratio_box), float(self.ratio_conf), float(self.ratio_cls),
                                                        bg, _lib.dptr(partials), _lib.dptr(saved_conf), _lib.dptr(grad_out),
                                                        _lib.head_ptrs(grads),
                                                        _lib.dptr(ws), _lib.stream()), "yolov3_loss_backward")
        return grads

    def combine(self, partials, batch_global, ctx=None, out=None):
        """Scalar loss from (all-reduced) per-level partials with GLOBAL-batch normalisers."""
        ctx = ctx or self._ctx
        if out is None:
            out = torch.empty(1, dtype=torch.float32, device=partials.device)
        lib = _lib.load()
        with torch.cuda.device(partials.device):
            _lib.check(lib.fvb_yolov3_loss_combine_f32(ctx.geom, int(batch_global), _lib.dptr(partials),
                                                       float(self.ratio_box), float(self.ratio_conf),
                                                       float(self.ratio_cls), _lib.dptr(out), _lib.stream()), "loss_combine")
        return out

    def build_target(self, y_pred, y_true):
        """loss/yolov3_loss.py:75-124 -> (gt_locations, gt_categories, gt_xywh, matched_anchors), matches in (t,a) order."""
        heads = [_lib.require_cuda(h, "y_pred[%d]" % i) for i, h in enumerate(y_pred)]
        labels = _lib.require_cuda(y_true, "y_true").view(-1, 6)
        ctx = self._context(heads)
        dev = ctx.device
        t, a = labels.size(0), ctx.num_anchors
        lib = _lib.load()
        locs, cats, xywhs, anchs = [], [], [], []
        for lvl in range(ctx.geom.levels):
            n = t * a
            b = torch.empty(n, dtype=torch.int64, device=dev)
            gxy = torch.empty(n, 2, dtype=torch.int64, device=dev)
            ai = torch.empty(n, dtype=torch.int64, device=dev)
            cls = torch.empty(n, dtype=torch.int64, device=dev)
            xywh = torch.empty(n, 4, dtype=torch.float32, device=dev)
            anc = torch.empty(n, 2, dtype=torch.float32, device=dev)
            match = torch.empty(max(n, 1), dtype=torch.uint8, device=dev)
            count = torch.zeros(1, dtype=torch.int32, device=dev)
            with torch.cuda.device(dev):
                _lib.check(lib.fvb_yolov3_build_target_f32(ctx.geom, lvl, _lib.dptr(labels), t, 1, _lib.dptr(b),
                                                           _lib.dptr(gxy), _lib.dptr(ai), _lib.dptr(cls), _lib.dptr(xywh),
                                                           _lib.dptr(anc), _lib.dptr(match), _lib.dptr(count),
                                                           _lib.stream()), "build_target")
            m = int(count[0])  # ragged result: the reference syncs at the same point (boolean-mask index, :105)
            locs.append((b[:m], gxy[:m], ai[:m]))
            cats.append(cls[:m])
            xywhs.append(xywh[:m])
            anchs.append(anc[:m])
        return locs, cats, xywhs, anchs
