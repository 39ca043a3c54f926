"""Oracle: the demos' single-image ``postProcess``.  TEST INFRASTRUCTURE ONLY.

Follows demos/yolov3_u/inference.py:55-121 (form "v5") and demos/yolov3_huaweiShip/inference.py:90-141 (form "v3") with
torch fp32 ops on CPU in the reference's order; the NMS is oracle.nms.nms_demo (demos/yolov3_u/utils/nms.py:5-53).
"""
import torch

from .boxes import xywh2xyxy
from .demo_loss import demo_grid_xy
from .nms import nms_demo


def post_process(predict_layers, strides, anchors, conf_thres, iou_thres, resize_ratio, padding_left, padding_top, ori_width,
                 ori_height, form="v5", max_det=300):
    ori_predict = []
    for layer_idx in range(len(predict_layers)):
        predict = predict_layers[layer_idx].clone()
        anchor = anchors[layer_idx]
        num_anchors = anchor.size(0)
        stride = strides[layer_idx]
        bs, c, h, w = predict.size()                                                              # :69
        num_classes = c // num_anchors - 5
        predict = predict.permute(0, 2, 3, 1).view(bs, h, w, num_anchors, -1).clone()              # :72
        grid_xy = demo_grid_xy(h, w).repeat(1, 1, 1, 1).unsqueeze(3).to(predict)                   # :74
        anchor_wh = anchor.repeat(1, 1, 1, 1, 1)                                                   # :75
        if form == "v5":
            predict[..., 0:2] = (torch.sigmoid(predict[..., 0:2]) * 2 - 0.5 + grid_xy) * stride   # :86
            predict[..., 2:4] = (torch.sigmoid(predict[..., 2:4]) * 2) ** 2 * anchor_wh * stride   # :87
        else:
            predict[..., 0:2] = (torch.sigmoid(predict[..., 0:2]) + grid_xy) * stride             # huaweiShip :113
            predict[..., 2:4] = (torch.exp(predict[..., 2:4]) * anchor_wh) * stride               # :114
        predict[..., 4:5] = torch.sigmoid(predict[..., 4:5])                                       # :88
        predict[..., 5:] = torch.sigmoid(predict[..., 5:])                                         # :89
        predict = predict.reshape(-1, num_classes + 5)
        predict[:, 0] = (predict[:, 0] - padding_left) / resize_ratio                              # :92-95
        predict[:, 1] = (predict[:, 1] - padding_top) / resize_ratio
        predict[:, 2] = predict[:, 2] / resize_ratio
        predict[:, 3] = predict[:, 3] / resize_ratio
        predict[:, 0] = predict[:, 0].clamp(0, ori_width - 1)                                      # :97-100
        predict[:, 1] = predict[:, 1].clamp(0, ori_height - 1)
        predict[:, 2] = predict[:, 2].clamp(0, ori_width)
        predict[:, 3] = predict[:, 3].clamp(0, ori_height)
        keep = (predict[..., 2] > 5) & (predict[..., 3] > 5)                                       # :102
        predict = predict[keep, :]
        predict[:, 0:4] = xywh2xyxy(predict[:, 0:4])                                               # :105
        predict[:, 0] = predict[:, 0].clamp(0, ori_width - 1)                                      # :106-109
        predict[:, 1] = predict[:, 1].clamp(0, ori_height - 1)
        predict[:, 2] = predict[:, 2].clamp(0, ori_width - 1)
        predict[:, 3] = predict[:, 3].clamp(0, ori_height - 1)
        ori_predict.append(predict)
    ori_predict = torch.cat(ori_predict, dim=0)
    results = nms_demo(ori_predict, conf_thres, iou_thres, max_det)                                # :113
    return results[:, 4:5], results[:, 5:6], results[:, :4], ori_predict
