"""Oracle: the demos' training loss ``ComputeLoss``.  TEST INFRASTRUCTURE ONLY.

Follows demos/yolov3_huaweiShip/utils/lossv3.py:19-125 (flavour "ship": CIoU box term, returns
(loss_box, loss_cls, loss_conf)) and demos/yolov3_u/utils/lossv3.py:17-119 (flavour "u": BCE-with-logits xy + MSE wh,
returns 2*xy + wh + cls + conf) with torch fp32 ops on CPU in the reference's order.  Both use the demos' own IoU file
(demos/<x>/utils/iou.py == oracle.iou with variant="demo").  Heads are the raw conv outputs [B, A*K, H, W]; anchors are
per level [A,2] in FEATURE units (model.anchors, demos/yolov3_huaweiShip/models/yolov3.py:150).
``partials=True`` also returns, per level, the sums / counts a data-parallel run all-reduces.
"""
import torch
import torch.nn.functional as F

from .iou import wh_iou_batch, xywh_iou_batch, CIOU


def demo_grid_xy(height, width):
    """demos/yolov3_huaweiShip/utils/box.py:36-47, mode='xy' -> [H,W,2] holding (x, y)."""
    ys = torch.arange(0, height)
    xs = torch.arange(0, width)
    ox, oy = torch.meshgrid(xs, ys, indexing="ij")
    return torch.stack([ox, oy]).permute(1, 2, 0).permute(1, 0, 2)


def compute_loss(predict_layers, target_all, anchor_layers, flavour="ship", partials=False):
    z = torch.zeros(1).to(predict_layers[0])
    loss_box, loss_cls, loss_conf, loss_xy, loss_wh = z.clone(), z.clone(), z.clone(), z.clone(), z.clone()
    parts = []
    for layer_idx in range(len(predict_layers)):
        anchor = anchor_layers[layer_idx].to(predict_layers[0])                                   # :38
        na = anchor.size(0)
        bs, _, fh, fw = predict_layers[layer_idx].size()                                          # :42
        predict = predict_layers[layer_idx].permute(0, 2, 3, 1).view(bs, fh, fw, na, -1)          # :43
        target = target_all.clone()                                                                # :46
        target[:, 2:] = target[:, 2:] * torch.tensor([fw, fh, fw, fh]).to(predict)                 # :47
        iou_ta = wh_iou_batch(target[:, 4:], anchor)                                               # :52
        _, best = torch.max(iou_ta, dim=1)                                                         # :53
        anc_t = anchor[best, :]                                                                    # :54
        gxy = torch.floor(target[:, 2:4])                                                          # :57
        off = target[:, 2:4] - gxy.float()                                                         # :58
        target = torch.cat([target, gxy, off, best.unsqueeze(1), anc_t], dim=1).float()            # :62
        p_xy = torch.sigmoid(predict[..., 0:2])                                                    # :65
        p_wh = torch.exp(predict[..., 2:4]) * anchor.repeat(1, 1, 1, 1, 1)                         # :67
        grid_xy = demo_grid_xy(fh, fw).repeat(1, 1, 1, 1).unsqueeze(3).to(p_xy)                    # :68
        p_xywh = torch.cat([p_xy.float() + grid_xy.float(), p_wh.float()], dim=4)                  # :69
        ib, iy, ix, ia = target[:, 0].long(), target[:, 7].long(), target[:, 6].long(), target[:, 10].long()
        s_a = s_b = torch.zeros((), dtype=torch.float64)
        if flavour == "ship":
            ciou = CIOU(p_xywh[ib, iy, ix, ia, :], target[:, 2:6], mode="xywh", eps=1e-7, variant="demo")   # :85-87
            loss_box += (1 - ciou).mean()                                                          # :88
            s_a = (1 - ciou).double().sum()
        else:
            pxy = predict[ib, iy, ix, ia, 0:2]                                                     # yolov3_u/utils/lossv3.py:71
            lxy = F.binary_cross_entropy_with_logits(pxy, target[:, 8:10])                         # :73
            loss_xy += lxy
            pwh = predict[ib, iy, ix, ia, 2:4]                                                     # :76
            twh = torch.log((target[:, 4:6] / target[:, 11:13]) + 1e-14)                           # :77
            loss_wh += F.mse_loss(pwh, twh)                                                        # :78
            s_a = F.binary_cross_entropy_with_logits(pxy, target[:, 8:10], reduction="sum").double()
            s_b = F.mse_loss(pwh, twh, reduction="sum").double()
        p_cls = predict[ib, iy, ix, ia, 5:]                                                        # :92
        t_cls = torch.zeros_like(p_cls)
        t_cls[range(len(t_cls)), target[:, 1].long()] = 1                                          # :94
        loss_cls += F.binary_cross_entropy_with_logits(p_cls, t_cls)                               # :95
        s_cls = F.binary_cross_entropy_with_logits(p_cls, t_cls, reduction="sum").double()
        masks = []
        for gt_idx in range(bs):                                                                   # :102-113
            pb = p_xywh[gt_idx, ...]
            tb = target[target[:, 0] == gt_idx][:, 2:6]
            iou_tp = xywh_iou_batch(pb.reshape(-1, 4), tb)
            mx, _ = torch.max(iou_tp, dim=1)              # raises IndexError for an image without targets, as the reference
            m = torch.zeros((iou_tp.size(0), 1)).to(predict)
            m[mx > 0.5] = -1
            masks.append(m.view(fh, fw, na, 1).unsqueeze(0))
        mask = torch.cat(masks, 0)
        mask[ib, iy, ix, ia, ...] = 1                                                              # :115
        p_conf = predict[..., 4:5][mask != -1]                                                     # :118
        t_conf = mask[mask != -1]
        loss_conf += F.binary_cross_entropy_with_logits(p_conf, t_conf)                            # :120
        s_conf = F.binary_cross_entropy_with_logits(p_conf, t_conf, reduction="sum").double()
        parts.append([float(torch.as_tensor(v).detach()) for v in (s_a, s_b, s_cls, s_conf)] + [float(t_conf.numel())])
    if flavour == "ship":
        out = (loss_box, loss_cls, loss_conf)                                                      # :125
    else:
        loss_xy = loss_xy * 2.0                                                                    # yolov3_u/utils/lossv3.py:111
        out = loss_xy + loss_wh + loss_cls + loss_conf                                             # :117
    return (out, parts) if partials else out
