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

    def __init__(self, head_out, anchors_per_level, strides):
        h0 = head_out[0]
        self.batch, self.num_anchors, _, _, self.k = h0.shape
        self.heights = [int(h.size(2)) for h in head_out]
        self.widths = [int(h.size(3)) for h in head_out]
        for h in head_out:
            if h.dim() != 5 or h.size(0) != self.batch or h.size(1) != self.num_anchors or h.size(4) != self.k:
                raise ValueError("head tensors must all be [B,A,H,W,K]; got %s" % [tuple(x.shape) for x in head_out])
        self.geom = _lib.make_geom(self.batch, self.k, self.heights, self.widths, strides, anchors_per_level)
        lib = _lib.load()
        self.rows = lib.fvb_yolo_rows_per_image(self.geom)
        if self.rows < 0:
            _lib.check(-1, "geometry")
        self.words = lib.fvb_yolo_bitmap_words(self.geom)
        self._sched = None
        self.device = h0.device
        self.key = (self.batch, self.num_anchors, self.k, tuple(self.heights), tuple(self.widths), self.device)
        self._bitmap = None
        self._bce0 = None
        self._rec = None

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
                  conf_thres=None, want_bce0=False):
    """Decode raw heads (list of [B,A,H,W,K]) into ``results`` [B,N,K]  (yolov3.py:36-51).

    With ``conf_thres`` the kernel also fills ``ctx.bitmap()`` / ``ctx.records()`` (NMS candidates and their
    score / class / box records); with ``want_bce0`` it fills ``ctx.bce0()`` (zero-target objectness BCE
    partials for Yolov3Loss).
    """
    heads = [_lib.require_cuda(h, "head_out[%d]" % i) for i, h in enumerate(head_out)]
    if ctx is None:
        ctx = DecodeContext(heads, anchors_per_level, strides)
    if out is None:
        out = torch.empty(ctx.batch, ctx.rows, ctx.k, dtype=torch.float32, device=ctx.device)
    lib = _lib.load()
    bitmap = ctx.bitmap() if conf_thres is not None else None
    rec = ctx.records() if conf_thres is not None else None
    bce0 = ctx.bce0() if want_bce0 else None
    with torch.cuda.device(ctx.device):
        _lib.check(lib.fvb_yolo_decode_f32(ctx.geom, _lib.head_ptrs(heads), _lib.DECODE_FORMS[form], 1 if precise else 0,
                                           _lib.dptr(out), float(conf_thres if conf_thres is not None else 0.0),
                                           _lib.dptr(bitmap), _lib.dptr(rec), _lib.dptr(bce0), _lib.dptr(ctx.sched()),
                                           _lib.stream()),
                   "yolo_decode")
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
            return (head_out, self.decode(head_out))
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
