"""Oracle: NMS arithmetic and the reference's NMS front-ends.  TEST INFRASTRUCTURE ONLY.

``nms_greedy`` restates the published algorithm of ``torchvision.ops.nms`` (third-party; the
reference pins torchvision==0.11.2+cu113, demos/yolov3_huaweiShip/requirements.txt:28-30; its
source is not in the reference tree).  Semantics (probed on torchvision 0.26.0 CPU):
stable descending sort of scores; area=(x2-x1)*(y2-y1); walk the order, an unsuppressed i is
kept and suppresses every later j with inter/(area_i+area_j-inter) > thr (strict, fp32 IEEE
divide); NaN never suppresses; returns kept original indices, score-descending.
Call sites in the reference: detection/tools/NMS.py:18; demos/yolov3_u/utils/nms.py:47,92;
demos/faster_rcnn/utils/nms.py:33,78; demos/faster_rcnn/models/rpn.py:198.
Pinned by tests/golden/nms_*.npz (outputs of torchvision 0.26.0 CPU) and live when it imports.

Front-ends: ``nms_lib`` = detection/tools/NMS.py:5-23; ``nms_demo`` =
demos/yolov3_u/utils/nms.py:5-53; ``nms_demo_batch`` = same file :55-98;
``nms_frcnn`` = demos/faster_rcnn/utils/nms.py:5-39.
"""
import numpy as np
import torch

from .boxes import xywh2xyxy


def nms_greedy(boxes, scores, iou_thr, return_iou_margin=False):
    """boxes [n,4] xyxy fp32, scores [n] fp32 -> int64 kept indices (score desc).

    With ``return_iou_margin`` also returns min |iou - thr| over every pair the walk evaluated
    (used by tests to excuse pairs within 1e-6 of the threshold).
    """
    b = boxes.detach().cpu().numpy().astype(np.float32, copy=False)
    s = scores.detach().cpu().numpy().astype(np.float32, copy=False)
    n = b.shape[0]
    order = np.argsort(-s, kind="stable")
    x1, y1, x2, y2 = b[:, 0], b[:, 1], b[:, 2], b[:, 3]
    area = (x2 - x1) * (y2 - y1)
    dead = np.zeros(n, dtype=bool)
    keep = []
    margin = np.inf
    thr = float(iou_thr)  # the CPU op compares the fp32 ratio, promoted, against the double threshold
    with np.errstate(invalid="ignore", divide="ignore"):
        for pos in range(n):
            i = order[pos]
            if dead[i]:
                continue
            keep.append(i)
            rest = order[pos + 1:]
            if rest.size == 0:
                continue
            w = np.maximum(np.float32(0), np.minimum(x2[i], x2[rest]) - np.maximum(x1[i], x1[rest]))
            h = np.maximum(np.float32(0), np.minimum(y2[i], y2[rest]) - np.maximum(y1[i], y1[rest]))
            inter = w * h
            ovr = inter / (area[i] + area[rest] - inter)
            if return_iou_margin:
                live = ~dead[rest]
                if live.any():
                    d = np.abs(ovr[live].astype(np.float64) - thr)
                    d = d[~np.isnan(d)]
                    if d.size:
                        margin = min(margin, float(d.min()))
            dead[rest[ovr.astype(np.float64) > thr]] = True
    keep = torch.from_numpy(np.asarray(keep, dtype=np.int64))
    return (keep, margin) if return_iou_margin else keep


def _nms(boxes, scores, thr, backend):
    if backend == "torchvision":
        import torchvision
        return torchvision.ops.nms(boxes, scores, thr)
    return nms_greedy(boxes, scores, thr)


def nms_lib(prediction, conf_thres=0.25, iou_thres=0.45, max_det=300, backend="numpy"):
    """detection/tools/NMS.py:5-23 -> (scores[k,1], categories[k,1] int64, boxes[k,4] xyxy)."""
    cand = prediction[prediction[..., 4] > conf_thres]          # :7-8 (copy; a [B,N,K] input flattens)
    if cand.size(0) == 0:                                       # :10-11 three empty CPU float tensors
        return torch.Tensor().view(-1, 1), torch.Tensor().view(-1, 1), torch.Tensor().view(-1, 4)
    cls_scores = cand[:, 5:] * cand[:, 4:5]                     # :13
    boxes = xywh2xyxy(cand[:, :4])                              # :14
    scores, cats = torch.max(cls_scores, dim=1)                 # :16 (first max on ties)
    keep = _nms(boxes, scores, iou_thres, backend)              # :18 class-agnostic
    keep = keep[:max_det]
    return scores[keep].view(-1, 1), cats[keep].view(-1, 1), boxes[keep].view(-1, 4)


def nms_demo(prediction, conf_thres=0.25, iou_thres=0.45, max_det=300, backend="numpy",
             max_wh=4096, max_nms=30000):
    """demos/yolov3_u/utils/nms.py:5-53 -> [k,6] = [x1,y1,x2,y2,obj,cat]; boxes arrive as xyxy."""
    p = prediction[prediction[:, 4] > conf_thres]                # :18
    if len(p) == 0:
        return torch.zeros((0, 6), device=prediction.device)
    cls_scores = p[:, 5:] * p[:, 4:5]                            # :24
    _, cats = cls_scores.max(1, keepdim=True)                    # :25
    keep = p[:, 4] > conf_thres                                  # :29 (second filter, again on obj)
    p = torch.cat([p[keep, :5], cats[keep].view(-1, 1).to(p.dtype)], dim=1)  # :30-36
    if p.size(0) > max_nms:                                      # :39-41
        p = p[p[:, 4].argsort(descending=True)[:max_nms]]
    gap = p[:, 5:6] * max_wh                                     # :44 fp32
    k = _nms(p[:, :4] + gap, p[:, 4], iou_thres, backend)        # :45-47 ranked by obj
    return p[k[:max_det]]


def nms_demo_batch(prediction_batch, conf_thres=0.25, iou_thres=0.45, max_det=300, backend="numpy",
                   max_wh=4096, max_nms=30000):
    """demos/yolov3_u/utils/nms.py:55-98 -> list of CPU [k,6] = [x1,y1,x2,y2,score,cat]."""
    out = [torch.zeros((0, 6))] * len(prediction_batch)
    for i, pred in enumerate(prediction_batch):
        p = pred[pred[:, 4] > conf_thres]
        cls_scores = p[:, 5:] * p[:, 4:5]
        boxes = xywh2xyxy(p[:, :4])
        scores, cats = cls_scores.max(1, keepdim=True)
        p = torch.cat((boxes, scores, cats.float()), 1)[scores.view(-1) > conf_thres]
        if p.size(0) > max_nms:
            p = p[p[:, 4].argsort(descending=True)[:max_nms]]
        gap = p[:, 5:6] * max_wh
        k = _nms(p[:, :4] + gap, p[:, 4], iou_thres, backend)
        out[i] = p[k[:max_det]].detach().cpu()
    return out


def nms_frcnn(prediction, conf_thres=0.25, iou_thres=0.45, max_det=300, backend="numpy", max_wh=4096, max_nms=30000):
    """demos/faster_rcnn/utils/nms.py:5-39: rows [x1,y1,x2,y2,cat,score] -> kept rows (same 6 columns), score-descending.

    Filter ``score > conf_thres`` (:18); more than ``max_nms`` rows keep the top by COLUMN 0 (sic, :24); class gap
    ``cat * max_wh`` added in fp32 (:30-31); nms ranked by score (:33); first ``max_det`` (:34-35).
    """
    p = prediction[prediction[:, 5] > conf_thres]
    if len(p) == 0:
        return torch.zeros((0, 6), device=prediction.device)
    if p.size(0) > max_nms:
        p = p[p[:, 0].argsort(descending=True)[:max_nms]]
    boxes = p[..., :4] + p[..., 4:5] * max_wh
    k = _nms(boxes, p[..., 5], iou_thres, backend)
    return p[k[:max_det]]
