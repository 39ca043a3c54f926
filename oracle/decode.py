"""Oracle: YOLOv3 head decode.  TEST INFRASTRUCTURE ONLY.

Restates detection/models/yolov3.py:33-53 (the module itself cannot be imported: line 4
imports ``offset`` which does not exist; per SURVEY F1/F2 it must have been
``grid(h, w, mode='yx', dtype='numpy')`` -> [H,W,2] with [...,0]=column (x), [...,1]=row (y)).
Anchors are per level ``[A,1,1,2]`` in pixels (yolov3.py:11-17), strides [32,16,8]
(classfication/models/darknet53.py:106-107).  Row order inside a level: a*H*W + y*W + x;
levels concatenated in the order given.

``form='v5'`` is the demos' YOLOv5-style variant (demos/yolov3_u/inference.py:86-89):
xy = (sigmoid*2 - 0.5 + g)*s, wh = (sigmoid*2)^2 * anchor_px.
"""
import torch

from .boxes import grid


def decode(head_out, anchors_per_level, strides, form="v3"):
    results = []
    for lvl, raw in enumerate(head_out):
        bs, na, h, w, k = raw.shape
        cell = torch.from_numpy(grid(h, w, mode="yx", dtype="numpy").copy()).to(raw)  # [H,W,2] (x,y)
        cell = cell.expand_as(raw[..., 0:2])
        anc = anchors_per_level[lvl].to(raw).view(na, 1, 1, 2).expand_as(raw[..., 2:4])
        if form == "v3":
            xy = (raw[..., 0:2].sigmoid() + cell) * strides[lvl]
            wh = torch.exp(raw[..., 2:4]) * anc
        elif form == "v5":
            xy = (raw[..., 0:2].sigmoid() * 2 - 0.5 + cell) * strides[lvl]
            wh = (raw[..., 2:4].sigmoid() * 2) ** 2 * anc
        else:
            raise ValueError(form)
        dec = torch.cat((xy, wh, raw[..., 4:].sigmoid()), -1)
        results.append(dec.view(bs, -1, k))
    return torch.cat(results, 1)


def decode_demo_nchw(predict_layers, anchors_feat_levels, strides, order="yxa"):
    """The demos' decode of the raw NCHW conv outputs, batch kept (no un-letterbox): list of [B, A*K, H, W] ->
    [B, N, K].  ``order="yxa"`` restates demos/yolov3_huaweiShip/inference.py:107-117 (permute(0,2,3,1).view(bs,h,w,A,K),
    rows (y,x,a)); ``order="ayx"`` restates customize_service.py:437-447 (view(bs,A,K,h,w).permute(0,1,3,4,2), rows
    (a,y,x)).  Anchors are in FEATURE units and re-multiplied by the stride: wh = (exp(t) * a_feat) * stride.
    """
    outs = []
    for lvl, predict in enumerate(predict_layers):
        anchor = anchors_feat_levels[lvl].view(-1, 2)
        na = anchor.size(0)
        stride = strides[lvl]
        bs, c, h, w = predict.shape
        ys, xs = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
        cell = torch.stack([xs, ys], -1).to(predict)                                       # [H,W,2] = (x, y)
        if order == "yxa":
            p = predict.permute(0, 2, 3, 1).reshape(bs, h, w, na, -1).clone()
            g = cell.view(1, h, w, 1, 2)
            a = anchor.view(1, 1, 1, na, 2)
        else:
            p = predict.view(bs, na, c // na, h, w).permute(0, 1, 3, 4, 2).contiguous().clone()
            g = cell.view(1, 1, h, w, 2)
            a = anchor.view(1, na, 1, 1, 2)
        p[..., 0:2] = (torch.sigmoid(p[..., 0:2]) + g) * stride
        p[..., 2:4] = (torch.exp(p[..., 2:4]) * a) * stride
        p[..., 4:] = torch.sigmoid(p[..., 4:])
        outs.append(p.reshape(bs, -1, c // na))
    return torch.cat(outs, 1)
