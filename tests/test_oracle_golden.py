"""The oracle against the golden vectors recorded from the real reference (CPU, no GPU).

Every comparison is bit-exact unless stated: the oracle restates the reference with the same torch
ops in the same order, so on the same torch build the bits agree; ``CLOSE`` (rtol 1e-6) is used so
the fixtures also hold on a host whose torch-CPU vector path (AVX2 vs AVX512 exp/atan) differs.
"""
import numpy as np
import pytest
import torch

import oracle
from oracle import iou as oi, loss as ol, nms as on, boxes as ob
from conftest import T
from small_cfg import SMALL


def CLOSE(a, b, rtol=1e-6, atol=1e-7):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol)
    assert a.shape == b.shape, (a.shape, b.shape)


def test_box_and_grid(golden_iou):
    g = golden_iou
    CLOSE(ob.xywh2xyxy(T(g["a_xywh"])), g["xywh2xyxy_a"])
    CLOSE(ob.xyxy2xywh(T(g["a"])), g["a_xywh"])
    CLOSE(ob.xyxy2xywhn(T(g["a"]), 80, 120), g["xyxy2xywhn_a"])
    for mode in ("xy", "yx"):
        assert np.array_equal(ob.grid(3, 5, mode, "torch").numpy(), g["grid_torch_" + mode])
        assert np.array_equal(ob.grid(3, 5, mode, "numpy"), g["grid_numpy_" + mode])


@pytest.mark.parametrize("kind,fn", [("iou", oi.cal_iou), ("giou", oi.GIOU), ("diou", oi.DIOU), ("ciou", oi.CIOU)])
def test_iou_elementwise(golden_iou, kind, fn):
    g = golden_iou
    CLOSE(fn(T(g["a"]), T(g["b"]), mode="xyxy"), g["ew_%s_xyxy" % kind])
    CLOSE(fn(T(g["a_xywh"]), T(g["b_xywh"]), mode="xywh"), g["ew_%s_xywh" % kind])


@pytest.mark.parametrize("kind,fn", [("iou", oi.cal_iou_batch), ("giou", oi.GIOU_batch), ("diou", oi.DIOU_batch), ("ciou", oi.CIOU_batch)])
def test_iou_pairwise(golden_iou, kind, fn):
    g = golden_iou
    CLOSE(fn(T(g["a"]), T(g["c"]), mode="xyxy"), g["pw_%s_xyxy" % kind])
    CLOSE(fn(T(g["a_xywh"]), T(g["c_xywh"]), mode="xywh"), g["pw_%s_xywh" % kind])


def test_iou_wh_and_demo_variant(golden_iou):
    g = golden_iou
    CLOSE(oi.cal_iou(T(g["a_xywh"])[:, 2:], T(g["b_xywh"])[:, 2:], mode="wh"), g["ew_iou_wh"])
    CLOSE(oi.cal_iou_batch(T(g["a_xywh"])[:, 2:], T(g["c_xywh"])[:, 2:], mode="wh"), g["pw_iou_wh"])
    CLOSE(oi.DIOU(T(g["a"]), T(g["b"]), variant="demo"), g["demo_ew_diou_xyxy"])
    CLOSE(oi.CIOU(T(g["a_xywh"]), T(g["b_xywh"]), mode="xywh", variant="demo"), g["demo_ew_ciou_xywh"])
    CLOSE(oi.CIOU_batch(T(g["a"]), T(g["c"]), variant="demo"), g["demo_pw_ciou_xyxy"])
    with pytest.raises(Exception, match="mode must be"):
        oi.cal_iou(T(g["a"]), T(g["b"]), mode="nope")


@pytest.mark.parametrize("kind", ["iou", "giou", "diou", "ciou"])
def test_iou_losses(golden_iou, kind):
    g = golden_iou
    a, b, w = T(g["a"]), T(g["b"]), T(g["w"])
    CLOSE(oi.iou_loss(kind, a, b), g["loss_%s_mean" % kind], rtol=1e-5)
    CLOSE(oi.iou_loss(kind, T(g["a_xywh"]), T(g["b_xywh"]), mode="xywh", reduction="sum"), g["loss_%s_sum_xywh" % kind], rtol=1e-5)
    CLOSE(oi.iou_loss(kind, a, b, weights=w), g["loss_%s_mean_w" % kind], rtol=1e-5)


def test_bce(golden_iou):
    g = golden_iou
    lg, idx = T(g["bce_logits"]), T(g["bce_idx"])
    CLOSE(ol.bi_cross_entropy(lg, idx), g["bce_mean"], rtol=1e-5)
    CLOSE(ol.bi_cross_entropy(lg.sigmoid(), idx, already_sigmoid=True), g["bce_mean_sig"], rtol=1e-5)
    CLOSE(ol.bi_cross_entropy(lg, idx, reduction="sum"), g["bce_sum"], rtol=1e-5)
    CLOSE(ol.bi_cross_entropy(T(g["bce1_logits"]), T(g["bce1_tgt"])), g["bce1_mean"], rtol=1e-5)


def _heads(g, prefix="head"):
    return [T(g["%s%d" % (prefix, i)]) for i in range(3)]


def test_decode(golden_yolo):
    g = golden_yolo
    res = oracle.decode.decode(_heads(g), SMALL.anchors_levels(), SMALL.strides)
    CLOSE(res, g["decoded"])


def test_build_target_and_loss(golden_yolo):
    g = golden_yolo
    heads, labels = _heads(g), T(g["labels"])
    locs, cats, xywh, anchs = ol.build_target(heads, labels, SMALL.anchors_levels(), SMALL.strides)
    for l in range(3):
        assert np.array_equal(locs[l][0].numpy(), g["bt_b%d" % l])
        assert np.array_equal(locs[l][1].numpy(), g["bt_gxy%d" % l])
        assert np.array_equal(locs[l][2].numpy(), g["bt_a%d" % l])
        assert np.array_equal(cats[l].numpy(), g["bt_cls%d" % l])
        CLOSE(xywh[l], g["bt_xywh%d" % l])
        CLOSE(anchs[l], g["bt_anc%d" % l])
    loss, parts = ol.yolov3_loss(heads, labels, SMALL.anchors_levels(), SMALL.strides, return_partials=True)
    CLOSE(loss, g["loss"], rtol=1e-5)
    # the all-reduce formula reproduces the scalar from the per-level sums
    b, c = heads[0].size(0), SMALL.num_classes
    tot = 0.0
    for l, (s_cls, s_box, s_conf, m) in enumerate(parts):
        cells = heads[l][..., 4].numel()
        if m:
            tot += 0.5 * s_cls / (m * c) + 0.05 * s_box / m
        tot += 1.0 * s_conf / cells
    np.testing.assert_allclose(tot * b, g["loss"][0], rtol=1e-5)
    CLOSE(ol.yolov3_loss(heads, labels[:0], SMALL.anchors_levels(), SMALL.strides), g["loss_nolabels"], rtol=1e-5)
    CLOSE(ol.yolov3_loss(_heads(g, "s_head"), T(g["s_labels"]), SMALL.anchors_levels(), SMALL.strides), g["s_loss"], rtol=1e-5)


@pytest.mark.parametrize("backend", ["numpy", "torchvision"])
def test_nms_lib_frontend(golden_yolo, backend):
    if backend == "torchvision":
        pytest.importorskip("torchvision")
    g = golden_yolo
    res = T(g["decoded"])
    for i in range(res.size(0)):
        for tag, (ct, it, md) in {"a": (0.25, 0.45, 300), "b": (0.05, 0.3, 20)}.items():
            s, c, b = on.nms_lib(res[i], ct, it, md, backend=backend)
            CLOSE(s, g["nms_%s_s%d" % (tag, i)])
            assert np.array_equal(c.numpy(), g["nms_%s_c%d" % (tag, i)])
            CLOSE(b, g["nms_%s_b%d" % (tag, i)])
    sres = oracle.decode.decode(_heads(g, "s_head"), SMALL.anchors_levels(), SMALL.strides)
    for i in range(2):
        s, c, b = on.nms_lib(sres[i], 0.25, 0.45, 300, backend=backend)
        CLOSE(s, g["s_nms_s%d" % i])
        assert np.array_equal(c.numpy(), g["s_nms_c%d" % i])
        CLOSE(b, g["s_nms_b%d" % i])


def test_nms_empty_returns_cpu_float_triplet():
    s, c, b = on.nms_lib(torch.zeros(5, 9), 0.25, 0.45, 300)
    assert s.shape == (0, 1) and c.shape == (0, 1) and b.shape == (0, 4)
    assert s.dtype == torch.float32 and c.dtype == torch.float32


def test_nms_demo_frontends(golden_yolo):
    g = golden_yolo
    res = T(g["decoded"])
    for i in range(res.size(0)):
        CLOSE(on.nms_demo(T(g["demo_in%d" % i]), 0.1, 0.3, 50), g["demo_nms%d" % i])
    outs = on.nms_demo_batch([res[i] for i in range(res.size(0))], 0.1, 0.3, 50)
    for i, o in enumerate(outs):
        CLOSE(o, g["demo_batch%d" % i])


@pytest.mark.parametrize("case", ["cluster", "ties", "degenerate", "gap", "single", "rpn_like"])
def test_nms_greedy_vs_torchvision_golden(golden_nms, case):
    g = golden_nms
    keep = on.nms_greedy(T(g[case + "_boxes"]), T(g[case + "_scores"]), float(g[case + "_thr"]))
    assert np.array_equal(keep.numpy(), g[case + "_keep"])


def test_nms_greedy_vs_torchvision_live():
    tv = pytest.importorskip("torchvision")
    gen = torch.Generator().manual_seed(123)
    for n in (0, 1, 2, 65, 257, 1000):
        xy = torch.rand(n, 2, generator=gen) * 60
        wh = torch.rand(n, 2, generator=gen) * 30 + 1
        b = torch.cat([xy, xy + wh], 1)
        s = torch.rand(n, generator=gen)
        for thr in (0.3, 0.45, 0.7):
            assert torch.equal(on.nms_greedy(b, s, thr), tv.ops.nms(b, s, thr))


def test_map(golden_map):
    g = golden_map
    est = oracle.map_.MapOracle(g["thr"])
    for i in range(int(g["n_img"])):
        before = len(est.correct_all_images)
        yp, yt = T(g["pred%d" % i]), T(g["true%d" % i])
        est.process_one(yp, yt)
        got = est.correct_all_images[-1] if len(est.correct_all_images) > before else np.zeros((0, 12))
        assert np.array_equal(got, g["correct%d" % i]), i
        rule = oracle.map_.match_rule(yp, yt, g["thr"])
        assert np.array_equal(rule, g["correct%d" % i][:, 2:].astype(bool)), i
    m_iou, m_cls, ids = est.fetch()
    np.testing.assert_allclose(m_iou, g["map_each_iou"], rtol=1e-12)
    np.testing.assert_allclose(m_cls, g["map_each_cls"], rtol=1e-12)
    assert ids == g["cls_ids"].tolist()


def test_rpn(golden_rpn):
    g = golden_rpn
    base = T(g["base_anchors_px"]) / 16
    for tag in ("a", "b"):
        pre, post, thr = g[tag + "_cfg"]
        props = oracle.rpn.filter_proposals(T(g[tag + "_cls"]), T(g[tag + "_reg"]), base, int(pre), int(post), float(thr))
        for i, p in enumerate(props):
            CLOSE(p, g["%s_prop%d" % (tag, i)])


def test_nms_frcnn_flavour_vs_reference_golden():
    """oracle.nms.nms_frcnn == demos/faster_rcnn/utils/nms.py:5-39 (recorded by oracle/make_golden.py gold_frcnn_nms)."""
    from conftest import load_golden, T
    from oracle import nms as on
    g = load_golden("frcnn_nms.npz")
    for tag in "abc":
        thr, iou, md = g[tag + "_cfg"]
        out = on.nms_frcnn(T(g[tag + "_pred"]), float(thr), float(iou), int(md))
        assert np.array_equal(out.numpy(), g[tag + "_out"]), tag


def GCLOSE(a, b):
    """Gradients: rtol 1e-5 with an absolute floor of 1e-6 of the largest entry (sums of signed terms cancel)."""
    CLOSE(a, b, rtol=1e-5, atol=1e-6 * float(np.abs(b).max()))


# ---- gradients: autograd through the oracle vs gradients recorded from the reference's own graph ------------------------
def test_yolov3_loss_grads(golden_grads):
    g = golden_grads
    heads = [T(g["head%d" % i]) for i in range(3)]
    labels = T(g["labels"])
    loss, grads = oracle.grad.yolov3_loss_grad(heads, labels, SMALL.anchors_levels(), SMALL.strides, upstream=float(g["upstream"]))
    CLOSE(loss, g["loss"])
    for i in range(3):
        GCLOSE(grads[i], g["grad%d" % i])
    _, grads0 = oracle.grad.yolov3_loss_grad(heads, labels[:0], SMALL.anchors_levels(), SMALL.strides)
    for i in range(3):
        GCLOSE(grads0[i], g["grad_nolabels%d" % i])


@pytest.mark.parametrize("kind", ["iou", "giou", "diou", "ciou"])
def test_iou_loss_grads(golden_grads, kind):
    g = golden_grads
    cases = {"xyxy_mean": ("a", "b", "xyxy", "mean", None), "xywh_sum": ("a_xywh", "b_xywh", "xywh", "sum", None),
             "xyxy_mean_w": ("a", "b", "xyxy", "mean", T(g["w"]))}
    for tag, (x, y, mode, red, w) in cases.items():
        _, ga, gb = oracle.grad.iou_loss_grad(kind, T(g[x]), T(g[y]), w, mode, red)
        GCLOSE(ga, g["g_%s_%s_pre" % (kind, tag)])
        GCLOSE(gb, g["g_%s_%s_true" % (kind, tag)])
    if kind == "iou":
        _, ga, gb = oracle.grad.iou_loss_grad("iou", T(g["a_xywh"])[:, 2:], T(g["b_xywh"])[:, 2:], None, "wh", "mean")
        GCLOSE(ga, g["g_iou_wh_mean_pre"])
        GCLOSE(gb, g["g_iou_wh_mean_true"])


def test_bce_loss_grads(golden_grads):
    g = golden_grads
    logits, idx = T(g["bce_logits"]), T(g["bce_idx"])
    for red in ("mean", "sum"):
        GCLOSE(oracle.grad.bce_loss_grad(logits, idx, reduction=red)[1], g["g_bce_%s" % red])
    GCLOSE(oracle.grad.bce_loss_grad(logits.sigmoid(), idx, already_sigmoid=True)[1], g["g_bce_sig_mean"])
    GCLOSE(oracle.grad.bce_loss_grad(T(g["bce1_logits"]), T(g["bce1_tgt"]))[1], g["g_bce1_mean"])


# ---- the demos' ComputeLoss (both flavours), values and gradients ---------------------------------------------------------
def test_demo_compute_loss(golden_demo_loss):
    g = golden_demo_loss
    heads = [T(g["head%d" % i]) for i in range(3)]
    anchors = [T(g["anchors%d" % i]) for i in range(3)]
    labels = T(g["labels"])
    hs = [h.clone().requires_grad_(True) for h in heads]
    lb, lc, lo = oracle.demo_loss.compute_loss(hs, labels, anchors, "ship")
    CLOSE(lb, g["ship_box"]); CLOSE(lc, g["ship_cls"]); CLOSE(lo, g["ship_conf"])
    w = g["ship_up"]
    (lb * float(w[0]) + lc * float(w[1]) + lo * float(w[2])).sum().backward()
    for i in range(3):
        GCLOSE(hs[i].grad, g["ship_grad%d" % i])
    hs = [h.clone().requires_grad_(True) for h in heads]
    lu = oracle.demo_loss.compute_loss(hs, labels, anchors, "u")
    CLOSE(lu, g["u_loss"])
    (lu * float(g["u_up"])).sum().backward()
    for i in range(3):
        GCLOSE(hs[i].grad, g["u_grad%d" % i])
    with pytest.raises(IndexError):                     # an image without targets: lossv3.py:107
        oracle.demo_loss.compute_loss(heads, labels[labels[:, 0] != 1], anchors, "ship")


# ---- the demos' postProcess (decode + un-letterbox + min-size filter + class-aware NMS) -----------------------------------
@pytest.mark.parametrize("demo,form", [("yolov3_u", "v5"), ("yolov3_huaweiShip", "v3")])
def test_demo_postprocess(golden_postprocess, demo, form):
    g = golden_postprocess
    heads = [T(g["head%d" % i]) for i in range(3)]
    anchors = [T(g["anchors%d" % i]) for i in range(3)]
    ct, it, rr, pl, pt, ow, oh = g["args"].tolist()
    s, c, b, _ = oracle.postprocess.post_process(heads, SMALL.strides, anchors, ct, it, rr, int(pl), int(pt), int(ow), int(oh), form)
    CLOSE(s, g[demo + "_scores"]); CLOSE(b, g[demo + "_boxes"])
    assert np.array_equal(c.numpy(), g[demo + "_cats"])


# ---- anchor k-means (detection/tools/ANCHOR.py) ----------------------------------------------------------------------------
@pytest.mark.parametrize("tag", ["a", "b"])
def test_anchor_kmeans(golden_anchor, tag):
    g = golden_anchor
    k, iters, seed = (int(v) for v in g[tag + "_cfg"])
    np.random.seed(seed)
    centers, cats = oracle.anchor.KMeans(g[tag + "_samples"].copy(), k).fit(iters)
    assert np.array_equal(np.asarray(centers), g[tag + "_centers"]) and np.array_equal(cats, g[tag + "_categories"])
