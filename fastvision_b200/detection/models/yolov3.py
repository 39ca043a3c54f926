"""YOLOv3 wrapper with the decode block on the GPU kernel -- drop-in for detection/models/yolov3.py.

Backbone / neck / head are dense-conv stacks (cuDNN's job, out of scope): the user
passes the same callables the reference takes.  Only ``forward``'s decode block
(detection/models/yolov3.py:33-53) is replaced: ~30 ATen launches + 3 H2D copies become one kernel.
"""
import torch
import torch.nn as nn

from ... import _lib


class DecodeContext:
    """Geometry + reusable buffers for decoding one head-stack shape.

    Owns (per shape) the ``results`` tensor, the NMS candidate bitmap and the objectness-BCE partials
    so that a steady-state step allocates nothing.
    """

    def __init__(self, head_out, anchors_per_level, strides, layout="bahwk"):
        h0 = head_out[0]
        self.layout = layout
        if layout == "nchw":
            # the conv output [B, A*K, H, W] itself (channel a*K + k): the head's permute().contiguous() copy
            # (detection/head/yolov3head.py:61-63) is folded into the decode kernel's loads
            a0 = anchors_per_level[0]
            self.num_anchors = int(a0.numel() // 2) if isinstance(a0, torch.Tensor) else len(a0)
            if h0.dim() != 4 or h0.size(1) % self.num_anchors:
                raise ValueError("nchw head tensors must be [B, A*K, H, W] with A=%d; got %s" % (self.num_anchors, tuple(h0.shape)))
            self.batch, self.k = h0.size(0), h0.size(1) // self.num_anchors
            self.heights = [int(h.size(2)) for h in head_out]
            self.widths = [int(h.size(3)) for h in head_out]
            for h in head_out:
                if h.dim() != 4 or h.size(0) != self.batch or h.size(1) != self.num_anchors * self.k:
                    raise ValueError("head tensors must all be [B,A*K,H,W]; got %s" % [tuple(x.shape) for x in head_out])
        elif layout == "bahwk":
            self.batch, self.num_anchors, _, _, self.k = h0.shape
            self.heights = [int(h.size(2)) for h in head_out]
            self.widths = [int(h.size(3)) for h in head_out]
            for h in head_out:
                if h.dim() != 5 or h.size(0) != self.batch or h.size(1) != self.num_anchors or h.size(4) != self.k:
                    raise ValueError("head tensors must all be [B,A,H,W,K]; got %s" % [tuple(x.shape) for x in head_out])
        else:
            raise ValueError("layout must be 'bahwk' or 'nchw'")
        self.geom = _lib.make_geom(self.batch, self.k, self.heights, self.widths, strides, anchors_per_level, layout)
        lib = _lib.load()
        self.rows = lib.fvb_yolo_rows_per_image(self.geom)
        if self.rows < 0:
            _lib.check(-1, "geometry")
        self.words = lib.fvb_yolo_bitmap_words(self.geom)
        self._sched = None
        self.device = h0.device
        self.key = (self.batch, self.num_anchors, self.k, tuple(self.heights), tuple(self.widths), self.device)
        if layout != "bahwk":
            self.key = self.key + (layout,)
        self._bitmap = None
        self._bce0 = None
        self._rec = None
        self._tile_sync = None
        self.tiles_per_image = lib.fvb_yolo_decode_tiles_per_image(self.geom)

    def tile_sync(self):
        """[B+1] u32 hand-shake between ``fvb_yolo_decode_sync_f32`` and ``fvb_yolo_nms_after_decode_f32``: finished tiles per
        image + the consumer grid's exit counter.  Zero once; every decode + NMS pair leaves it zero."""
        if self._tile_sync is None:
            self._tile_sync = torch.zeros(self.batch + 1, dtype=torch.int32, device=self.device)
        return self._tile_sync

    def bitmap(self):
        if self._bitmap is None:  # zero once; the NMS kernel clears what it consumes
            self._bitmap = torch.zeros(self.batch, self.words, dtype=torch.int32, device=self.device)
        return self._bitmap

    def records(self):
        """[B,N,8] NMS candidate records (32 B per row; only candidate rows are ever written)."""
        if self._rec is None:
            self._rec = torch.empty(self.batch, self.rows, 8, dtype=torch.float32, device=self.device)
        return self._rec

    def sched(self):
        """Tile queue of the persistent decode kernel: zero once, every launch leaves it zero (one per context)."""
        if self._sched is None:
            self._sched = torch.zeros(_lib.load().fvb_yolo_decode_workspace_bytes(), dtype=torch.uint8, device=self.device)
        return self._sched

    def bce0(self):
        if self._bce0 is None:
            n = _lib.load().fvb_yolo_decode_partials(self.geom)
            if n < 0:
                _lib.check(-3, "decode_partials")
            self._bce0 = torch.empty(n, dtype=torch.float64, device=self.device)
        return self._bce0


def yolov3_decode(head_out, anchors_per_level, strides, form="v3", precise=False, ctx=None, out=None,
                  conf_thres=None, want_bce0=False, layout="bahwk", row_order="ayx", tile_sync=None):
    """Decode raw heads (list of [B,A,H,W,K]) into ``results`` [B,N,K]  (yolov3.py:36-51).

    ``layout="nchw"`` takes the conv outputs [B,A*K,H,W] directly (demos/yolov3_huaweiShip/customize_service.py:437).
    ``row_order="yxa"`` returns each level's rows in (y, x, a) order -- the order of the demos' single-image
    ``postProcess`` (demos/yolov3_huaweiShip/inference.py:107, view [B,H,W,A,K]); that is a permuted view-copy of
    the kernel's (a, y, x) result and drops the fused NMS side outputs' row numbering, so use it only for that API.

    With ``conf_thres`` the kernel also fills ``ctx.bitmap()`` / ``ctx.records()`` (NMS candidates and their
    score / class / box records); with ``want_bce0`` it fills ``ctx.bce0()`` (zero-target objectness BCE
    partials for Yolov3Loss).  ``tile_sync`` (``ctx.tile_sync()``): publish per-image progress for an NMS launch that
    follows DIRECTLY on the same stream with the same tensor (``non_max_suppression_batched(..., tile_sync=...)``).
    """
    heads = [_lib.require_cuda(h, "head_out[%d]" % i) for i, h in enumerate(head_out)]
    if ctx is None:
        ctx = DecodeContext(heads, anchors_per_level, strides, layout)
    if out is None:
        out = torch.empty(ctx.batch, ctx.rows, ctx.k, dtype=torch.float32, device=ctx.device)
    lib = _lib.load()
    bitmap = ctx.bitmap() if conf_thres is not None else None
    rec = ctx.records() if conf_thres is not None else None
    bce0 = ctx.bce0() if want_bce0 else None
    with torch.cuda.device(ctx.device):
        _lib.check(lib.fvb_yolo_decode_sync_f32(ctx.geom, _lib.head_ptrs(heads), _lib.DECODE_FORMS[form], 1 if precise else 0,
                                                _lib.dptr(out), float(conf_thres if conf_thres is not None else 0.0),
                                                _lib.dptr(bitmap), _lib.dptr(rec), _lib.dptr(bce0), _lib.dptr(tile_sync),
                                                _lib.dptr(ctx.sched()), _lib.stream()),
                   "yolo_decode")
    if row_order == "yxa":
        parts, r0 = [], 0
        for h, w in zip(ctx.heights, ctx.widths):
            n = ctx.num_anchors * h * w
            parts.append(out[:, r0:r0 + n].view(ctx.batch, ctx.num_anchors, h * w, ctx.k).permute(0, 2, 1, 3).reshape(ctx.batch, n, ctx.k))
            r0 += n
        return torch.cat(parts, 1)
    if row_order != "ayx":
        raise ValueError("row_order must be 'ayx' or 'yxa'")
    return out


class Yolov3(nn.Module):
    """Same constructor and ``forward(images, val=False)`` contract as detection/models/yolov3.py:6-54."""

    def __init__(self, backbone, neck, head, anchors, num_anchors_per_level, in_channels=3, num_classes=80, training=False):
        super(Yolov3, self).__init__()
        self.training = training
        anchors = anchors.view(-1, 2)
        self.anchors_per_level = []
        start = 0
        for n in num_anchors_per_level:                      # yolov3.py:11-17: [A,1,1,2] pixel anchors per level
            self.anchors_per_level.append(anchors[start:start + n].view(n, 1, 1, 2))
            start += n
        self.num_classes = num_classes
        self.backbone = backbone(in_channels=in_channels, including_top=False)
        self.backbone_strides_per_level = self.backbone.backbone_strides_per_level()
        self.backbone_channels_per_level = self.backbone.backbone_channels_per_level()
        self.neck = neck(feature_channels=self.backbone_channels_per_level)
        self.head = head(feature_channels=self.backbone_channels_per_level, num_levels=len(self.backbone_channels_per_level),
                         num_anchors_per_level=num_anchors_per_level, num_classes=num_classes)
        self._ctx = None
        # False: forward returns (head_out, None) in eval mode -- for callers that decode the raw heads themselves
        # (fastvision_b200.utils.Fit._val runs the fused decode + NMS + loss step and would otherwise decode twice)
        self.decode_in_forward = True

    def decode(self, head_out):
        ctx = self._ctx
        probe = DecodeContext(head_out, self.anchors_per_level, self.backbone_strides_per_level) if ctx is None else None
        if ctx is None or ctx.key != (head_out[0].size(0), head_out[0].size(1), head_out[0].size(4),
                                      tuple(int(h.size(2)) for h in head_out), tuple(int(h.size(3)) for h in head_out),
                                      head_out[0].device):
            ctx = probe or DecodeContext(head_out, self.anchors_per_level, self.backbone_strides_per_level)
            self._ctx = ctx
        return yolov3_decode(head_out, self.anchors_per_level, self.backbone_strides_per_level, ctx=ctx)

    def forward(self, images, val=False):
        head_out = self.head(self.neck(self.backbone(images)))
        if self.training == False or val == True:            # noqa: E712  (yolov3.py:34)
            return (head_out, self.decode(head_out) if self.decode_in_forward else None)
        return head_out


def yolov3(backbone=None, neck=None, head=None, anchors=None, num_anchors_per_level=None, in_channels=3, num_classes=80,
           training=False):
    """Factory with the reference's signature (yolov3.py:57-69).  The conv stacks are not part of this
    package, so unlike the reference there are no defaults to fall back to."""
    if backbone is None or neck is None or head is None:
        raise ValueError("fastvision_b200 ships the decode path only: pass backbone, neck and head callables "
                         "(e.g. fastvision's darknet53 / yolov3neck / yolov3head)")
    return Yolov3(backbone=backbone, neck=neck, head=head, anchors=anchors, num_anchors_per_level=num_anchors_per_level,
                  in_channels=in_channels, num_classes=num_classes, training=training)
