"""Drop-in for the hot-path part of demos/yolov3_huaweiShip/inference.py: ``postProcess`` (:90-141, YOLOv3 decode form :113-114)."""
from ..postprocess import post_process


def postProcess(predict_layers, strides, anchors, conf_thres, iou_thres, resize_ratio, padding_left, padding_top, ori_width, ori_height):
    return post_process(predict_layers, strides, anchors, conf_thres, iou_thres, resize_ratio, padding_left, padding_top,
                        ori_width, ori_height, form="v3")
