"""Drop-in for ``fastvision.detection.tools`` (detection/tools/__init__.py:1-5): same names, CUDA kernels behind."""
from .box import xywh2xyxy, xyxy2xywh, xyxy2xywhn
from .grid import grid, offset
from .iou import (cal_iou, cal_iou_batch, xyxy_iou, xywh_iou, wh_iou, xyxy_iou_batch, xywh_iou_batch, wh_iou_batch,
                  GIOU, GIOU_batch, DIOU, DIOU_batch, CIOU, CIOU_batch)
from .nms import (non_max_suppression, non_max_suppression_batched, non_max_suppression_demo, non_max_suppression_batch,
                  non_max_suppression_frcnn, nms)
from .rpn import filter_proposals, filter_proposals_batched, make_anchors_xywh, get_base_anchor
from .anchor import KMeans, AnchorGenerator

__all__ = [
    "xywh2xyxy", "xyxy2xywh", "xyxy2xywhn", "grid", "offset",
    "cal_iou", "cal_iou_batch", "xyxy_iou", "xywh_iou", "wh_iou", "xyxy_iou_batch", "xywh_iou_batch", "wh_iou_batch",
    "GIOU", "GIOU_batch", "DIOU", "DIOU_batch", "CIOU", "CIOU_batch",
    "non_max_suppression", "non_max_suppression_batched", "non_max_suppression_demo", "non_max_suppression_batch",
    "non_max_suppression_frcnn", "nms",
    "filter_proposals", "filter_proposals_batched", "make_anchors_xywh", "get_base_anchor",
    "KMeans", "AnchorGenerator",
]
