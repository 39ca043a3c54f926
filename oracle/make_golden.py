"""Generate tests/golden/*.npz from the REAL reference.  TEST INFRASTRUCTURE ONLY (container-only).

Run in the build container (where /root/reference is mounted):

    python -m oracle.make_golden

Imports the unmodified reference through ``oracle/ref_shim.py`` and records, for seeded inputs,
the outputs of the reference's own functions on the hot path plus the
installed torchvision CPU ``nms`` (the third-party arithmetic the reference calls).  The fixtures
pin the oracle (tests/test_oracle_golden.py, CPU) and the CUDA kernels (tests/test_*_gpu.py).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from fastvision_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def _np(x):
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


def rand_boxes_xyxy(n, gen, scale=100.0, degenerate=True):
    xy = torch.rand(n, 2, generator=gen) * scale
    wh = torch.rand(n, 2, generator=gen) * scale * 0.5 + 0.5
    b = torch.cat([xy, xy + wh], dim=1)
    if degenerate and n >= 8:
        b[1] = b[0]                      # coincident pair
        b[2, 2] = b[2, 0]                # zero width
        b[3, 3] = b[3, 1]                # zero height
        b[4] = torch.tensor([0.0, 0.0, 0.0, 0.0])
    return b


def gold_boxes_iou(ns):
    g = torch.Generator().manual_seed(1)
    t = ns.tools
    n, m = 64, 37
    a = rand_boxes_xyxy(n, g)
    b = rand_boxes_xyxy(n, g, degenerate=False)
    b[:16] = a[:16] + torch.randn(16, 4, generator=g) * 3.0     # real overlaps
    b[5] = a[5]
    c = rand_boxes_xyxy(m, g)
    c[:10] = a[:10] + torch.randn(10, 4, generator=g) * 2.0
    aw, bw, cw = t.xyxy2xywh(a), t.xyxy2xywh(b), t.xyxy2xywh(c)
    out = dict(a=_np(a), b=_np(b), c=_np(c), a_xywh=_np(aw), b_xywh=_np(bw), c_xywh=_np(cw))
    out["xywh2xyxy_a"] = _np(t.xywh2xyxy(aw))
    out["xyxy2xywhn_a"] = _np(t.xyxy2xywhn(a, 80, 120))
    for name, fn in [("iou", t.cal_iou), ("giou", t.GIOU), ("diou", t.DIOU), ("ciou", t.CIOU)]:
        out["ew_%s_xyxy" % name] = _np(fn(a, b, mode="xyxy"))
        out["ew_%s_xywh" % name] = _np(fn(aw, bw, mode="xywh"))
    out["ew_iou_wh"] = _np(t.cal_iou(aw[:, 2:], bw[:, 2:], mode="wh"))
    for name, fn in [("iou", t.cal_iou_batch), ("giou", t.GIOU_batch), ("diou", t.DIOU_batch), ("ciou", t.CIOU_batch)]:
        out["pw_%s_xyxy" % name] = _np(fn(a, c, mode="xyxy"))
        out["pw_%s_xywh" % name] = _np(fn(aw, cw, mode="xywh"))
    out["pw_iou_wh"] = _np(t.cal_iou_batch(aw[:, 2:], cw[:, 2:], mode="wh"))
    # demo variant (demos/yolov3_u/utils/iou.py)
    d = ns.load_demo("yolov3_u", "iou")
    out["demo_ew_diou_xyxy"] = _np(d.DIOU(a, b, mode="xyxy"))
    out["demo_ew_ciou_xywh"] = _np(d.CIOU(aw, bw, mode="xywh"))
    out["demo_pw_ciou_xyxy"] = _np(d.CIOU_batch(a, c, mode="xyxy"))
    # losses (loss/iou_loss.py), with and without weights
    w = torch.rand(n, 1, generator=g)
    out["w"] = _np(w)
    for name, cls in [("iou", ns.loss.IOULoss), ("giou", ns.loss.GIOULoss), ("diou", ns.loss.DIOULoss), ("ciou", ns.loss.CIOULoss)]:
        out["loss_%s_mean" % name] = _np(cls("mean")(a, b))
        out["loss_%s_sum_xywh" % name] = _np(cls("sum")(aw, bw, mode="xywh"))
        out["loss_%s_mean_w" % name] = _np(cls("mean")(a, b, weights=w))
    # BCE (loss/classification_loss.py)
    logits = torch.randn(33, 7, generator=g) * 4
    logits[0, 0] = 40.0
    logits[1, 1] = -40.0
    idx = torch.randint(0, 7, (33,), generator=g)
    bce = ns.loss.BiCrossEntropyLoss("mean")
    out["bce_logits"], out["bce_idx"] = _np(logits), _np(idx)
    out["bce_mean"] = _np(bce(logits, idx))
    out["bce_mean_sig"] = _np(bce(logits.sigmoid(), idx, already_sigmoid=True))
    out["bce_sum"] = _np(ns.loss.BiCrossEntropyLoss("sum")(logits, idx))
    one = torch.randn(50, 1, generator=g) * 3
    tgt = torch.rand(50, 1, generator=g)
    out["bce1_logits"], out["bce1_tgt"] = _np(one), _np(tgt)
    out["bce1_mean"] = _np(bce(one, tgt))
    # grid
    for mode in ("xy", "yx"):
        out["grid_torch_%s" % mode] = _np(t.grid(3, 5, mode=mode, dtype="torch"))
        out["grid_numpy_%s" % mode] = _np(t.grid(3, 5, mode=mode, dtype="numpy"))
    np.savez_compressed(os.path.join(OUT, "iou_family.npz"), **out)


SMALL = synth.YoloConfig("tiny", 64, 4, [[40, 30], [50, 60], [30, 50], [20, 24], [16, 10], [12, 22], [4, 6], [8, 5], [7, 9]],
                         labels_per_img=4.0, max_labels=9)


def small_case(seed, batch=3, stress=False):
    g = torch.Generator().manual_seed(seed)
    labels = synth.make_labels(SMALL, batch, g)
    heads = synth.make_heads(SMALL, batch, labels, g, stress=stress)
    return labels, heads


def gold_decode_nms_loss(ns):
    out = {}
    labels, heads = small_case(7)
    # edge labels: x == 1.0 (grid clamp, yolov3_loss.py:116), duplicate cell, tiny box matching no anchor
    extra = torch.tensor([[0, 1, 1.0, 0.5, 0.3, 0.4],
                          [1, 2, 0.26, 0.26, 0.3, 0.35],
                          [1, 3, 0.27, 0.27, 0.32, 0.33],
                          [2, 0, 0.5, 0.5, 0.001, 0.001]], dtype=torch.float32)
    labels = torch.cat([labels, extra], 0)
    anc = SMALL.anchors_levels()
    res = ns.decode(heads, anc, SMALL.strides, SMALL.num_classes)
    out["labels"] = _np(labels)
    for i, h in enumerate(heads):
        out["head%d" % i] = _np(h)
    out["decoded"] = _np(res)

    class Model:
        anchors_per_level = anc
        backbone_strides_per_level = SMALL.strides

    lossf = ns.Yolov3Loss(Model(), 0.5, 0.05, 1.0, 0.5)
    out["loss"] = _np(lossf(heads, labels))
    locs, cats, xywh, anchs = lossf.build_target(heads, labels)
    for l in range(3):
        out["bt_b%d" % l] = _np(locs[l][0])
        out["bt_gxy%d" % l] = _np(locs[l][1])
        out["bt_a%d" % l] = _np(locs[l][2])
        out["bt_cls%d" % l] = _np(cats[l])
        out["bt_xywh%d" % l] = _np(xywh[l])
        out["bt_anc%d" % l] = _np(anchs[l])
    out["loss_nolabels"] = _np(lossf(heads, labels[:0]))
    # library NMS per image (detection/tools/NMS.py), low threshold so clusters really suppress
    for i in range(res.size(0)):
        for tag, (ct, it, md) in {"a": (0.25, 0.45, 300), "b": (0.05, 0.3, 20)}.items():
            s, c, b = ns.tools.non_max_suppression(res[i], ct, it, md)
            out["nms_%s_s%d" % (tag, i)] = _np(s)
            out["nms_%s_c%d" % (tag, i)] = _np(c)
            out["nms_%s_b%d" % (tag, i)] = _np(b)
    # demo class-aware NMS (demos/yolov3_u/utils/nms.py): wants xyxy boxes
    demo = ns.load_demo("yolov3_u", "nms")
    for i in range(res.size(0)):
        p = res[i].clone()
        p[:, :4] = ns.tools.xywh2xyxy(p[:, :4])
        out["demo_in%d" % i] = _np(p)
        out["demo_nms%d" % i] = _np(demo.non_max_suppression(p.clone(), 0.1, 0.3, 50))
    bl = demo.non_max_suppression_batch([res[i].clone() for i in range(res.size(0))], 0.1, 0.3, 50)
    for i, o in enumerate(bl):
        out["demo_batch%d" % i] = _np(o)
    # stress logits: many candidates
    labels2, heads2 = small_case(11, batch=2, stress=True)
    res2 = ns.decode(heads2, anc, SMALL.strides, SMALL.num_classes)
    out["s_labels"] = _np(labels2)
    for i, h in enumerate(heads2):
        out["s_head%d" % i] = _np(h)
    out["s_loss"] = _np(lossf(heads2, labels2))
    for i in range(2):
        s, c, b = ns.tools.non_max_suppression(res2[i], 0.25, 0.45, 300)
        out["s_nms_s%d" % i], out["s_nms_c%d" % i], out["s_nms_b%d" % i] = _np(s), _np(c), _np(b)
    np.savez_compressed(os.path.join(OUT, "yolo_small.npz"), **out)


def gold_tv_nms():
    """torchvision.ops.nms CPU (0.26.0) on clustered boxes, exact ties and degenerate boxes."""
    import torchvision
    out = {"torchvision_version": np.array(torchvision.__version__)}
    g = torch.Generator().manual_seed(3)
    cases = {}
    ctr = torch.rand(12, 2, generator=g) * 200
    pick = torch.randint(0, 12, (300,), generator=g)
    xy = ctr[pick] + torch.randn(300, 2, generator=g) * 6
    wh = 30 + torch.rand(300, 2, generator=g) * 20
    boxes = torch.cat([xy - wh / 2, xy + wh / 2], 1)
    cases["cluster"] = (boxes, torch.rand(300, generator=g), 0.45)
    sc = torch.rand(300, generator=g)
    sc[10:40] = sc[10]                   # 30-way score tie: lower index first
    cases["ties"] = (boxes, sc, 0.5)
    deg = boxes.clone()
    deg[::7, 2] = deg[::7, 0]            # zero-area boxes: NaN-free but area 0
    deg[5] = 0.0
    deg[6] = 0.0                         # 0/0 -> NaN never suppresses
    cases["degenerate"] = (deg, torch.rand(300, generator=g), 0.3)
    big = torch.cat([boxes + 4096.0 * torch.randint(0, 5, (300, 1), generator=g).float()], 1)
    cases["gap"] = (big, torch.rand(300, generator=g), 0.45)
    cases["single"] = (boxes[:1], sc[:1], 0.5)
    cases["rpn_like"] = (boxes.clamp(0, 49) / 4, torch.rand(300, generator=g), 0.7)
    for k, (b, s, thr) in cases.items():
        out[k + "_boxes"], out[k + "_scores"], out[k + "_thr"] = _np(b), _np(s), np.float64(thr)
        out[k + "_keep"] = _np(torchvision.ops.nms(b, s, thr))
    np.savez_compressed(os.path.join(OUT, "tv_nms.npz"), **out)


def gold_map(ns):
    out = {}
    g = torch.Generator().manual_seed(5)
    thr = np.linspace(0.5, 0.95, 10)
    est = ns.metrics.CalculateMAP(thr)
    n_img = 12
    out["n_img"] = np.int64(n_img)
    for i in range(n_img):
        nt = int(torch.randint(0, 6, (1,), generator=g)) if i != 3 else 0
        tb = rand_boxes_xyxy(nt, g, degenerate=False)
        tc = torch.randint(0, 3, (nt, 1), generator=g).float()
        y_true = torch.cat([tc, tb], 1).view(-1, 5)
        preds = []
        for j in range(nt):
            reps = int(torch.randint(0, 4, (1,), generator=g))
            for _ in range(reps):
                jit = torch.randn(4, generator=g) * float(torch.rand(1, generator=g)) * 6
                c = tc[j] if float(torch.rand(1, generator=g)) < 0.8 else (tc[j] + 1) % 3
                preds.append(torch.cat([c, torch.rand(1, generator=g), tb[j] + jit]))
        nf = int(torch.randint(0, 4, (1,), generator=g)) if i != 5 else 0
        fb = rand_boxes_xyxy(nf, g, degenerate=False)
        for j in range(nf):
            preds.append(torch.cat([torch.randint(0, 4, (1,), generator=g).float(), torch.rand(1, generator=g), fb[j]]))
        y_pred = torch.stack(preds).view(-1, 6) if preds else torch.zeros(0, 6)
        if i == 7 and nt:                      # exact duplicate detections -> IoU ties between preds
            y_pred = torch.cat([y_pred, y_pred[:2]], 0)
        before = len(est.correct_all_images)
        est.process_one(y_pred, y_true)
        out["pred%d" % i], out["true%d" % i] = _np(y_pred), _np(y_true)
        out["correct%d" % i] = est.correct_all_images[-1] if len(est.correct_all_images) > before else np.zeros((0, 12))
    m_iou, m_cls, ids = est.fetch()
    out["map_each_iou"], out["map_each_cls"], out["cls_ids"] = m_iou, m_cls, np.asarray(ids)
    out["thr"] = thr
    np.savez_compressed(os.path.join(OUT, "map_small.npz"), **out)


def gold_rpn(ns):
    rpn_mod = ns.load_rpn()
    import math
    base = []
    for r in [1, 0.5, 2]:
        for s in [128, 256, 512]:
            w = math.sqrt(s ** 2 / r)
            base.append((w, s ** 2 / w))
    base = torch.tensor(np.array(base, dtype=np.float32))
    out = {"base_anchors_px": _np(base)}
    g = torch.Generator().manual_seed(9)
    for tag, (b, fh, fw, pre, post, thr) in {"a": (2, 6, 5, 2000, 2000, 0.7), "b": (2, 12, 10, 300, 40, 0.5)}.items():
        rpn = rpn_mod.RPN(training=False, base_anchors=base, backbone_stride=16, in_channels=8,
                          rpn_pre_nms_top_n=pre, rpn_post_nms_top_n=post, rpn_nms_thresh=thr)
        cls, reg = synth.make_rpn_inputs(b, fh, fw, 9, g)
        anc = rpn.make_anchors_xywh(fh, fw, "cpu")
        props = rpn.filter_proposals(cls, reg, anc, fh, fw)
        out[tag + "_cls"], out[tag + "_reg"] = _np(cls), _np(reg)
        out[tag + "_cfg"] = np.array([pre, post, thr], dtype=np.float64)
        for i, p in enumerate(props):
            out["%s_prop%d" % (tag, i)] = _np(p)
    np.savez_compressed(os.path.join(OUT, "rpn_small.npz"), **out)


def gold_frcnn_nms(ns):
    """demos/faster_rcnn/utils/nms.py:5-39 on clustered class-tagged detections -> tests/golden/frcnn_nms.npz."""
    ref = ns.load_demo("faster_rcnn", "nms")
    g = torch.Generator().manual_seed(11)
    out = {}
    for tag, (n, ncls, thr, iou, md) in {"a": (400, 5, 0.25, 0.45, 300), "b": (60, 2, 0.5, 0.3, 10), "c": (32, 3, 0.99, 0.45, 300)}.items():
        ctr = torch.rand(n // 4, 2, generator=g) * 300 + 50
        ctr = ctr.repeat_interleave(4, 0) + torch.randn(n, 2, generator=g) * 6
        wh = torch.rand(n, 2, generator=g) * 60 + 20
        boxes = torch.cat([ctr - wh / 2, ctr + wh / 2], 1)
        cat = torch.randint(0, ncls, (n, 1), generator=g).float()
        score = torch.rand(n, 1, generator=g)
        pred = torch.cat([boxes, cat, score], 1)
        out[tag + "_pred"] = _np(pred)
        out[tag + "_cfg"] = np.array([thr, iou, md], dtype=np.float64)
        out[tag + "_out"] = _np(ref.non_max_suppression(pred, thr, iou, md))
    np.savez_compressed(os.path.join(OUT, "frcnn_nms.npz"), **out)


def gold_grads(ns):
    """Gradients recorded from the reference's own autograd graph (what loss.backward() in utils/fit.py:57-63 computes):
    Yolov3Loss w.r.t. the three raw head tensors (duplicate cell, clamped cell, unmatched label included), the four IoU
    losses w.r.t. y_pre / y_true in both box modes, and BiCrossEntropyLoss w.r.t. its logits -> tests/golden/grads_small.npz."""
    out = {}
    labels, heads = small_case(7)
    extra = torch.tensor([[0, 1, 1.0, 0.5, 0.3, 0.4],
                          [1, 2, 0.26, 0.26, 0.3, 0.35],
                          [1, 3, 0.27, 0.27, 0.32, 0.33],
                          [1, 1, 0.265, 0.262, 0.31, 0.36],
                          [2, 0, 0.5, 0.5, 0.001, 0.001]], dtype=torch.float32)
    labels = torch.cat([labels, extra], 0)
    anc = SMALL.anchors_levels()

    class Model:
        anchors_per_level = anc
        backbone_strides_per_level = SMALL.strides

    lossf = ns.Yolov3Loss(Model(), 0.5, 0.05, 1.0, 0.5)
    out["labels"] = _np(labels)
    hs = [h.clone().requires_grad_(True) for h in heads]
    loss = lossf(hs, labels)
    (loss * 1.7).sum().backward()                      # a non-unit upstream gradient
    out["loss"] = _np(loss)
    out["upstream"] = np.float32(1.7)
    for i, h in enumerate(hs):
        out["head%d" % i] = _np(heads[i])
        out["grad%d" % i] = _np(h.grad)
    hs = [h.clone().requires_grad_(True) for h in heads]
    lossf(hs, labels[:0]).sum().backward()
    for i, h in enumerate(hs):
        out["grad_nolabels%d" % i] = _np(h.grad)
    # IoU losses
    g = torch.Generator().manual_seed(21)
    n = 48
    a = rand_boxes_xyxy(n, g, degenerate=False)
    b = rand_boxes_xyxy(n, g, degenerate=False)
    b[:24] = a[:24] + torch.randn(24, 4, generator=g) * 3.0
    t = ns.tools
    aw, bw = t.xyxy2xywh(a), t.xyxy2xywh(b)
    w = torch.rand(n, 1, generator=g)
    out["a"], out["b"], out["a_xywh"], out["b_xywh"], out["w"] = _np(a), _np(b), _np(aw), _np(bw), _np(w)
    for name, cls in [("iou", ns.loss.IOULoss), ("giou", ns.loss.GIOULoss), ("diou", ns.loss.DIOULoss), ("ciou", ns.loss.CIOULoss)]:
        for tag, (x, y, mode, red, ww) in {"xyxy_mean": (a, b, "xyxy", "mean", None), "xywh_sum": (aw, bw, "xywh", "sum", None),
                                           "xyxy_mean_w": (a, b, "xyxy", "mean", w)}.items():
            xx, yy = x.clone().requires_grad_(True), y.clone().requires_grad_(True)
            cls(red)(xx, yy, weights=ww, mode=mode).backward()
            out["g_%s_%s_pre" % (name, tag)] = _np(xx.grad)
            out["g_%s_%s_true" % (name, tag)] = _np(yy.grad)
    xx, yy = aw[:, 2:].clone().requires_grad_(True), bw[:, 2:].clone().requires_grad_(True)
    ns.loss.IOULoss("mean")(xx, yy, mode="wh").backward()
    out["g_iou_wh_mean_pre"], out["g_iou_wh_mean_true"] = _np(xx.grad), _np(yy.grad)
    # BCE
    logits = torch.randn(33, 7, generator=g) * 3
    idx = torch.randint(0, 7, (33,), generator=g)
    out["bce_logits"], out["bce_idx"] = _np(logits), _np(idx)
    for red in ("mean", "sum"):
        x = logits.clone().requires_grad_(True)
        ns.loss.BiCrossEntropyLoss(red)(x, idx).backward()
        out["g_bce_%s" % red] = _np(x.grad)
    x = logits.sigmoid().clone().requires_grad_(True)
    ns.loss.BiCrossEntropyLoss("mean")(x, idx, already_sigmoid=True).backward()
    out["g_bce_sig_mean"] = _np(x.grad)
    one = torch.randn(50, 1, generator=g) * 3
    tgt = torch.rand(50, 1, generator=g)
    out["bce1_logits"], out["bce1_tgt"] = _np(one), _np(tgt)
    x = one.clone().requires_grad_(True)
    ns.loss.BiCrossEntropyLoss("mean")(x, tgt).backward()
    out["g_bce1_mean"] = _np(x.grad)
    np.savez_compressed(os.path.join(OUT, "grads_small.npz"), **out)


def to_nchw(h):
    """[B,A,H,W,K] -> the conv output layout [B,A*K,H,W] the demos' loss consumes."""
    b, a, hh, ww, k = h.shape
    return h.permute(0, 1, 4, 2, 3).reshape(b, a * k, hh, ww).contiguous()


def gold_demo_loss(ns):
    """ComputeLoss of both demos (forward values and autograd gradients w.r.t. the conv outputs) -> tests/golden/demo_loss.npz."""
    import contextlib
    import io
    import types
    out = {}
    labels, heads = small_case(13, batch=3)
    extra = torch.tensor([[1, 2, 0.26, 0.26, 0.3, 0.35],
                          [1, 3, 0.27, 0.27, 0.31, 0.34],      # same cell, same best anchor as the row above on some levels
                          [2, 0, 0.5, 0.5, 0.9, 0.8]], dtype=torch.float32)
    labels = torch.cat([labels, extra], 0)
    nchw = [to_nchw(h) for h in heads]
    anchors = [a.reshape(-1, 2) / s for a, s in zip(SMALL.anchors_levels(), SMALL.strides)]
    model = types.SimpleNamespace(anchors=anchors)
    out["labels"] = _np(labels)
    for i, h in enumerate(nchw):
        out["head%d" % i] = _np(h)
        out["anchors%d" % i] = _np(anchors[i])
    ship = ns.load_demo("yolov3_huaweiShip", "lossv3").ComputeLoss()
    hs = [h.clone().requires_grad_(True) for h in nchw]
    lb, lc, lo = ship(hs, labels, model)
    out["ship_box"], out["ship_cls"], out["ship_conf"] = _np(lb), _np(lc), _np(lo)
    w = torch.tensor([0.05, 0.5, 1.0])
    out["ship_up"] = _np(w)
    (lb * w[0] + lc * w[1] + lo * w[2]).sum().backward()
    for i, h in enumerate(hs):
        out["ship_grad%d" % i] = _np(h.grad)
    u = ns.load_demo("yolov3_u", "lossv3").ComputeLoss()
    hs = [h.clone().requires_grad_(True) for h in nchw]
    with contextlib.redirect_stdout(io.StringIO()):
        lu = u(hs, labels, model)
    out["u_loss"] = _np(lu)
    (lu * 0.3).sum().backward()
    out["u_up"] = np.float32(0.3)
    for i, h in enumerate(hs):
        out["u_grad%d" % i] = _np(h.grad)
    np.savez_compressed(os.path.join(OUT, "demo_loss.npz"), **out)


def _ref_function(path, name, namespace):
    """Compile ONE function of a reference file (unmodified source text) in ``namespace``: the demo scripts import cv2 /
    albumentations at module level and cannot be imported as a whole here."""
    import ast
    src = open(path).read()
    tree = ast.parse(src)
    node = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == name)
    code = compile(ast.Module(body=[node], type_ignores=[]), path, "exec")
    exec(code, namespace)
    return namespace[name]


def gold_postprocess(ns):
    """demos/<x>/inference.py postProcess (decode + un-letterbox + filter + class-aware NMS), both demos -> tests/golden/postprocess.npz."""
    out = {}
    labels, heads = small_case(17, batch=1)
    nchw = [to_nchw(h) for h in heads]
    anchors = [a.reshape(-1, 2) / s for a, s in zip(SMALL.anchors_levels(), SMALL.strides)]
    args = dict(conf_thres=0.2, iou_thres=0.4, resize_ratio=0.8, padding_left=3, padding_top=5, ori_width=72, ori_height=66)
    for i, h in enumerate(nchw):
        out["head%d" % i] = _np(h)
        out["anchors%d" % i] = _np(anchors[i])
    out["args"] = np.array([args[k] for k in ("conf_thres", "iou_thres", "resize_ratio", "padding_left", "padding_top", "ori_width", "ori_height")],
                           dtype=np.float64)
    for demo in ("yolov3_u", "yolov3_huaweiShip"):
        box = ns.load_demo(demo, "box")
        nms = ns.load_demo(demo, "nms")
        env = {"torch": torch, "grid": box.grid, "xywh2xyxy": box.xywh2xyxy, "non_max_suppression": nms.non_max_suppression}
        fn = _ref_function(os.path.join(ref_shim.REFERENCE_ROOT, "demos", demo, "inference.py"), "postProcess", env)
        scores, cats, boxes = fn([h.clone() for h in nchw], SMALL.strides, anchors, **args)
        out[demo + "_scores"], out[demo + "_cats"], out[demo + "_boxes"] = _np(scores), _np(cats), _np(boxes)
    np.savez_compressed(os.path.join(OUT, "postprocess.npz"), **out)


def anchor_samples(n, seed):
    """Normalised (w, h) samples around 9 COCO-like anchor shapes."""
    g = torch.Generator().manual_seed(seed)
    base = torch.tensor(synth.COCO416.anchors_px, dtype=torch.float32) / 416.0
    pick = torch.randint(0, 9, (n,), generator=g)
    wh = base[pick] * torch.exp(torch.randn(n, 2, generator=g) * 0.25)
    return wh.clamp(0.005, 1.0).numpy().astype(np.float32)


def gold_anchor(ns):
    """KMeans of detection/tools/ANCHOR.py (seeded numpy RNG) -> tests/golden/anchor_kmeans.npz."""
    import importlib
    A = importlib.import_module("fastvision.detection.tools.ANCHOR")
    out = {}
    for tag, (n, k, iters, seed) in {"a": (600, 9, 25, 3), "b": (50, 12, 8, 4)}.items():
        xs = anchor_samples(n, seed)
        out[tag + "_samples"] = xs.copy()
        out[tag + "_cfg"] = np.array([k, iters, seed], dtype=np.int64)
        np.random.seed(seed)
        centers, cats = A.KMeans(xs=xs.copy(), k=k).fit(iters=iters)
        out[tag + "_centers"], out[tag + "_categories"] = np.asarray(centers), np.asarray(cats)
    np.savez_compressed(os.path.join(OUT, "anchor_kmeans.npz"), **out)


def main():
    if "--only-frcnn-nms" in sys.argv:
        torch.set_num_threads(1)
        gold_frcnn_nms(ref_shim.load())
        return
    if "--only-demo-loss" in sys.argv:
        torch.set_num_threads(1)
        gold_demo_loss(ref_shim.load())
        return
    if "--only-postprocess" in sys.argv:
        torch.set_num_threads(1)
        gold_postprocess(ref_shim.load())
        return
    if "--only-anchor" in sys.argv:
        gold_anchor(ref_shim.load())
        return
    if "--only-grads" in sys.argv:
        torch.set_num_threads(1)
        gold_grads(ref_shim.load())
        return
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)
    ns = ref_shim.load()
    gold_boxes_iou(ns)
    gold_decode_nms_loss(ns)
    gold_tv_nms()
    gold_map(ns)
    gold_rpn(ns)
    gold_frcnn_nms(ns)
    gold_grads(ns)
    gold_demo_loss(ns)
    gold_postprocess(ns)
    gold_anchor(ns)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
