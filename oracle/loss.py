"""Oracle: BCE, target assignment and the YOLOv3 loss.  TEST INFRASTRUCTURE ONLY.

Follows loss/classification_loss.py:36-65 (BiCrossEntropyLoss), datasets/common/id_2_onehot.py:10-15,
loss/yolov3_loss.py:75-124 (build_target) and :29-72 (forward) with torch fp32 ops on CPU in the
reference's order.  The only deviation from the text of the reference is the int() cast of the clamp
bounds (yolov3_loss.py:116-117 crash on torch>=2) -- semantics unchanged.
``loss_partials`` additionally returns the per-level sums/counts the multi-GPU all-reduce carries
(SURVEY 8e): for each level [S_cls, S_box, S_conf, M].
"""
import torch

from .iou import CIOU, cal_iou


def one_hot(y, num_classes):
    """datasets/common/id_2_onehot.py:10-15 (torch branch)."""
    idx = y.view(-1, 1).long()
    return torch.zeros((idx.size(0), num_classes)).to(y).scatter_(1, idx, 1)


def bce_terms(p, target):
    """classification_loss.py:55 -- -t*log(p+1e-8) - (1-t)*log(1-p+1e-8), p already a probability."""
    return -target * torch.log(p + 1e-8) - (1 - target) * torch.log(1 - p + 1e-8)


def bi_cross_entropy(y_pre, y_true, already_sigmoid=False, weights=None, reduction="mean"):
    """classification_loss.py:42-65.  C==1 uses y_true itself as the target (:47-48)."""
    c = y_pre.size(-1)
    if c > 1:
        target = one_hot(y_true, c).float().view(-1, 1)
    else:
        target = y_true.float().view(-1, 1)
    flat = y_pre.view(-1, 1)
    p = flat if already_sigmoid else flat.sigmoid()
    loss = torch.sum(bce_terms(p, target), dim=1)
    if weights is None:
        weights = torch.ones_like(loss)
    loss = loss * weights
    if reduction == "mean":
        return torch.sum(loss) / flat.numel()
    return torch.sum(loss)


def build_target(y_pred, y_true, anchors_per_level, strides):
    """yolov3_loss.py:75-124.

    y_true[T,6] = [batch_idx, cls, xc, yc, w, h] normalised.  Per level returns
    ((b[M], grid_xy[M,2], a[M]) int64, cls[M] int64, xywh[M,4], anchors[M,2]); matches in (t,a) row-major order.
    """
    locs, cats, xywh, anchs = [], [], [], []
    for lvl, pre in enumerate(y_pred):
        anchors = anchors_per_level[lvl].to(pre).reshape(-1, 2) / strides[lvl]       # :88-89 feature units
        na = anchors.size(0)
        h, w = pre.size(2), pre.size(3)
        whwh = torch.tensor([w, h, w, h]).to(pre)                                     # :92
        tgt = y_true.clone()
        tgt[:, 2:] = y_true[:, 2:] * whwh                                             # :94-95
        ratio = tgt[:, None, 4:] / anchors                                            # :98  [T,A,2]
        ok = torch.max(ratio, 1 / ratio).max(2)[0] < 4                                # :99  [T,A]
        t_idx, a_idx = torch.nonzero(ok, as_tuple=True)                               # row-major == :101-105
        m = tgt[t_idx]
        b = m[:, 0].long()
        c = m[:, 1].long()
        xy = m[:, 2:4]
        gxy = torch.floor(xy).long()                                                  # :113
        off = xy - gxy.float()                                                        # :114 (before the clamp)
        gxy[:, 0] = gxy[:, 0].clamp(0, w - 1)                                         # :116
        gxy[:, 1] = gxy[:, 1].clamp(0, h - 1)                                         # :117
        locs.append((b, gxy, a_idx))
        cats.append(c)
        xywh.append(torch.cat([off, m[:, 4:6]], dim=1))
        anchs.append(anchors[a_idx])
    return locs, cats, xywh, anchs


def _level_terms(pre, loc, cat, txywh, anc):
    """One iteration of the level loop, yolov3_loss.py:39-64 -> (loss_cls_l, loss_box_l, loss_conf_l, sums)."""
    b, gxy, a = loc
    rows = pre[b, a, gxy[:, 1], gxy[:, 0]]                                            # :44
    tconf = torch.zeros_like(pre[..., 4:5])                                           # :48
    l_cls = l_box = None
    s_cls = s_box = torch.zeros((), dtype=torch.float64)
    m = b.size(0)
    if m:
        p_cls = rows[..., 5:].sigmoid()                                               # :50
        l_cls = bi_cross_entropy(p_cls, cat, already_sigmoid=True)                    # :52
        p_xy = rows[..., 0:2].sigmoid()                                               # :54
        p_wh = torch.exp(rows[..., 2:4]) * anc                                        # :55
        p_xywh = torch.cat([p_xy, p_wh], dim=1)
        ciou = CIOU(p_xywh, txywh, mode="xywh")
        l_box = torch.mean(1 - ciou)                                                  # :58 CIOULoss(mean)
        iou = cal_iou(p_xywh, txywh, mode="xywh")                                     # :60
        tconf[b, a, gxy[:, 1], gxy[:, 0]] = iou                                       # :61 duplicates: last wins on CPU
        s_cls = bce_terms(p_cls.reshape(-1, 1), one_hot(cat, p_cls.size(-1)).float().view(-1, 1)).double().sum()
        s_box = (1 - ciou).double().sum()
    p_conf = pre[..., 4:5].sigmoid()                                                  # :63
    l_conf = bi_cross_entropy(p_conf.view(-1, 1), tconf.view(-1, 1), already_sigmoid=True)  # :64
    s_conf = bce_terms(p_conf.view(-1, 1), tconf.view(-1, 1)).double().sum()
    return l_cls, l_box, l_conf, (s_cls, s_box, s_conf, m)


def yolov3_loss(y_pred, y_true, anchors_per_level, strides, ratio_box=0.05, ratio_conf=1.0, ratio_cls=0.5,
                return_partials=False):
    """yolov3_loss.py:29-72 -> Tensor[1] = (r_box*sum_l box_l + r_conf*sum_l conf_l + r_cls*sum_l cls_l) * B."""
    locs, cats, xywh, anchs = build_target(y_pred, y_true, anchors_per_level, strides)
    z = torch.zeros(1).to(y_pred[0])
    loss_cls, loss_box, loss_conf = z.clone(), z.clone(), z.clone()
    partials = []
    for lvl, pre in enumerate(y_pred):
        l_cls, l_box, l_conf, sums = _level_terms(pre, locs[lvl], cats[lvl], xywh[lvl], anchs[lvl])
        if l_cls is not None:
            loss_cls += l_cls
            loss_box += l_box
        loss_conf += l_conf
        partials.append([float(torch.as_tensor(v).detach()) for v in sums])
    loss_box *= ratio_box
    loss_conf *= ratio_conf
    loss_cls *= ratio_cls
    bs = y_pred[0].size(0)
    out = (loss_box + loss_conf + loss_cls) * bs
    return (out, partials) if return_partials else out
