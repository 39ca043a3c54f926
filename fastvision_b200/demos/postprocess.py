"""The demos' single-image ``postProcess`` -- drop-in for demos/yolov3_u/inference.py:55-121 (YOLOv5 decode form) and
demos/yolov3_huaweiShip/inference.py:90-141 (YOLOv3 form): decode of the raw conv outputs, un-letterbox / clamp /
min-size filter / xyxy, class-aware NMS -- three launches plus the row reorder instead of ~120."""
import torch

from .. import _lib
from ..detection.models.yolov3 import yolov3_decode
from ..detection.tools.nms import non_max_suppression_demo


def post_process(predict_layers, strides, anchors, conf_thres, iou_thres, resize_ratio, padding_left, padding_top, ori_width,
                 ori_height, form):
    heads = [_lib.require_cuda(h, "predict_layers[%d]" % i) for i, h in enumerate(predict_layers)]
    if heads[0].size(0) != 1:
        raise ValueError("postProcess handles one image (the demos feed batch 1; rows of several images would be mixed)")
    # anchors arrive in feature units (anchor_fn: px / stride); the decode kernel takes pixels + strides
    anchors_px = [a.detach().float().cpu().reshape(-1, 2) * float(s) for a, s in zip(anchors, strides)]
    rows = yolov3_decode(heads, anchors_px, [float(s) for s in strides], form=form, precise=True, layout="nchw", row_order="yxa")
    rows = rows.reshape(-1, rows.size(-1)).contiguous()
    lib = _lib.load()
    with torch.cuda.device(rows.device):
        _lib.check(lib.fvb_demo_boxes_postprocess_f32(_lib.dptr(rows), rows.size(0), rows.size(1), float(padding_left),
                                                      float(padding_top), float(resize_ratio), float(ori_width),
                                                      float(ori_height), 5.0, _lib.stream()), "demo_boxes_postprocess")
    results = non_max_suppression_demo(rows, conf_thres=conf_thres, iou_thres=iou_thres, max_det=300)
    return results[:, 4:5], results[:, 5:6], results[:, :4]        # scores, categories, boxes (inference.py:115-119)
