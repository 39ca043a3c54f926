"""GPU parity tests: CUDA kernels (through the C ABI) vs the golden vectors and the CPU oracle.

Tolerances: bit-exact for indices / classes / keep sets / mAP bits; rtol 1e-5 (atol 1e-6) for fp32
boxes, IoUs and losses -- the tolerance BASELINE.json's north_star states.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import oracle
from oracle import iou as oi, loss as ol, nms as on
from conftest import T
from gpu_util import cuda, close
from small_cfg import SMALL
from fastvision_b200 import synth
from fastvision_b200.detection import tools as ft
from fastvision_b200.detection.models import yolov3_decode, DecodeContext
from fastvision_b200 import loss as fl
from fastvision_b200.metrics import CalculateMAP
from fastvision_b200.pipeline import ValStep, ValPipeline, shard_labels


def _heads(g, prefix="head"):
    return [T(g["%s%d" % (prefix, i)]) for i in range(3)]


# ---------------------------------------------------------------- decode
@pytest.mark.parametrize("precise", [False, True])
def test_decode_golden(golden_yolo, precise):
    heads = [h.cuda() for h in _heads(golden_yolo)]
    res = yolov3_decode(heads, SMALL.anchors_levels(), SMALL.strides, precise=precise)
    close(res, golden_yolo["decoded"])


@pytest.mark.parametrize("cfg,batch", [(synth.COCO416, 8), (synth.COCO416, 3), (synth.SHIP608, 5)])
def test_decode_vs_oracle(cfg, batch):
    g = synth.make_generator(1)
    labels = synth.make_labels(cfg, batch, g)
    heads = synth.make_heads(cfg, batch, labels, g)
    want = oracle.decode.decode(heads, cfg.anchors_levels(), cfg.strides)
    got = yolov3_decode([h.cuda() for h in heads], cfg.anchors_levels(), cfg.strides)
    close(got, want)
    # v5 form: xy = (2*sigmoid - 0.5 + g)*stride cancels near the cell origin, so the absolute error of the
    # sigmoid (<= 2e-7) times 2*stride is the floor there: atol 2e-5 px on top of rtol 1e-5
    got5 = yolov3_decode([h.cuda() for h in heads], cfg.anchors_levels(), cfg.strides, form="v5")
    close(got5, oracle.decode.decode(heads, cfg.anchors_levels(), cfg.strides, form="v5"), atol=2e-5)
    got5p = yolov3_decode([h.cuda() for h in heads], cfg.anchors_levels(), cfg.strides, form="v5", precise=True)
    close(got5p, oracle.decode.decode(heads, cfg.anchors_levels(), cfg.strides, form="v5"), atol=4e-6)


def test_decode_nonsquare_odd_shapes():
    # ragged shapes: H != W, K even, 2 levels, 5 anchors, batch 1 -- exercises tile tails and index maths
    gen = torch.Generator().manual_seed(5)
    heads = [torch.randn(1, 5, 3, 7, 12, generator=gen), torch.randn(1, 5, 9, 4, 12, generator=gen)]
    anchors = [torch.rand(5, 1, 1, 2, generator=gen) * 50 + 1 for _ in range(2)]
    want = oracle.decode.decode(heads, anchors, [16, 8])
    got = yolov3_decode([h.cuda() for h in heads], anchors, [16, 8])
    close(got, want)


@pytest.mark.parametrize("cfg,stress", [(synth.COCO416, False), (synth.COCO416, True), (synth.SHIP608, False), (synth.SHIP608, True), (SMALL, True)],
                         ids=["coco416", "coco416-dense", "ship608", "ship608-dense", "tiny-dense"])
def test_decode_fused_side_outputs(cfg, stress):
    # the records have two code paths (warp-serial for sparse tiles with many classes, lane-per-row for few classes or
    # candidate-dense tiles): the dense variants (objectness ~N(0,1)) and the 10- / 4-class configurations cover the second
    batch = 4 if cfg is not synth.SHIP608 else 2
    g = synth.make_generator(1)
    labels = synth.make_labels(cfg, batch, g)
    heads = [h.cuda() for h in synth.make_heads(cfg, batch, labels, g, stress=stress)]
    ctx = DecodeContext(heads, cfg.anchors_levels(), cfg.strides)
    res = yolov3_decode(heads, cfg.anchors_levels(), cfg.strides, ctx=ctx, conf_thres=0.25, want_bce0=True)
    # bitmap == (conf > thr) on the values the kernel itself stored
    mask = (res[..., 4] > 0.25).cpu().numpy()
    words = ctx.bitmap().cpu().numpy().view(np.uint32)
    bits = ((words[:, :, None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(batch, -1)[:, :ctx.rows].astype(bool)
    assert np.array_equal(bits, mask)
    # candidate records: {row[0:5], max_c(cls*conf), argmax} for every candidate row, computed on the stored values
    rec = ctx.records().cpu()
    res_c = res.cpu()
    for b in range(batch):
        rows = np.nonzero(mask[b])[0]
        cand = res_c[b, rows]
        prod = cand[:, 5:] * cand[:, 4:5]
        best, arg = prod.max(1)
        assert torch.equal(rec[b, rows, :5], cand[:, :5])
        assert torch.equal(rec[b, rows, 5], best)
        assert np.array_equal(rec[b, rows, 6].numpy().view(np.int32), arg.numpy().astype(np.int32))
    # zero-target objectness BCE: sum of the partials == oracle sum over all cells
    want = 0.0
    for h in heads:
        p = h[..., 4].cpu().sigmoid().reshape(-1, 1)
        want += float(ol.bce_terms(p, torch.zeros_like(p)).double().sum())
    np.testing.assert_allclose(float(ctx.bce0().sum()), want, rtol=1e-6)


# ---------------------------------------------------------------- IoU family / losses
@pytest.mark.parametrize("kind,fn", [("iou", ft.cal_iou), ("giou", ft.GIOU), ("diou", ft.DIOU), ("ciou", ft.CIOU)])
def test_iou_elementwise_golden(golden_iou, kind, fn):
    g = golden_iou
    close(fn(cuda(g["a"]), cuda(g["b"]), mode="xyxy"), g["ew_%s_xyxy" % kind])
    close(fn(cuda(g["a_xywh"]), cuda(g["b_xywh"]), mode="xywh"), g["ew_%s_xywh" % kind])


@pytest.mark.parametrize("kind,fn", [("iou", ft.cal_iou_batch), ("giou", ft.GIOU_batch), ("diou", ft.DIOU_batch), ("ciou", ft.CIOU_batch)])
def test_iou_pairwise_golden(golden_iou, kind, fn):
    g = golden_iou
    close(fn(cuda(g["a"]), cuda(g["c"]), mode="xyxy"), g["pw_%s_xyxy" % kind])
    close(fn(cuda(g["a_xywh"]), cuda(g["c_xywh"]), mode="xywh"), g["pw_%s_xywh" % kind])


def test_iou_wh_demo_boxes_and_errors(golden_iou):
    g = golden_iou
    close(ft.cal_iou(cuda(g["a_xywh"])[:, 2:], cuda(g["b_xywh"])[:, 2:], mode="wh"), g["ew_iou_wh"])
    close(ft.cal_iou_batch(cuda(g["a_xywh"])[:, 2:], cuda(g["c_xywh"])[:, 2:], mode="wh"), g["pw_iou_wh"])
    close(ft.DIOU(cuda(g["a"]), cuda(g["b"]), variant="demo"), g["demo_ew_diou_xyxy"])
    close(ft.CIOU(cuda(g["a_xywh"]), cuda(g["b_xywh"]), mode="xywh", variant="demo"), g["demo_ew_ciou_xywh"])
    close(ft.CIOU_batch(cuda(g["a"]), cuda(g["c"]), variant="demo"), g["demo_pw_ciou_xyxy"])
    close(ft.xywh2xyxy(cuda(g["a_xywh"])), g["xywh2xyxy_a"])
    close(ft.xyxy2xywh(cuda(g["a"])), g["a_xywh"])
    close(ft.xyxy2xywhn(cuda(g["a"]), 80, 120), g["xyxy2xywhn_a"])
    with pytest.raises(Exception, match="mode must be"):
        ft.cal_iou(cuda(g["a"]), cuda(g["b"]), mode="nope")
    with pytest.raises(RuntimeError, match="CUDA only"):
        ft.cal_iou(T(g["a"]), T(g["b"]))
    assert ft.cal_iou_batch(cuda(g["a"])[:0], cuda(g["c"])).shape == (0, g["c"].shape[0])


def test_iou_pairwise_large_vs_oracle():
    gen = torch.Generator().manual_seed(2)
    xy = torch.rand(700, 2, generator=gen) * 300
    a = torch.cat([xy, xy + torch.rand(700, 2, generator=gen) * 80 + 1], 1)
    xy = torch.rand(333, 2, generator=gen) * 300
    b = torch.cat([xy, xy + torch.rand(333, 2, generator=gen) * 80 + 1], 1)
    for fn_g, fn_o in [(ft.cal_iou_batch, oi.cal_iou_batch), (ft.GIOU_batch, oi.GIOU_batch),
                       (ft.DIOU_batch, oi.DIOU_batch), (ft.CIOU_batch, oi.CIOU_batch)]:
        close(fn_g(a.cuda(), b.cuda()), fn_o(a, b))


@pytest.mark.parametrize("kind,cls", [("iou", fl.IOULoss), ("giou", fl.GIOULoss), ("diou", fl.DIOULoss), ("ciou", fl.CIOULoss)])
def test_iou_losses_golden(golden_iou, kind, cls):
    g = golden_iou
    a, b, w = cuda(g["a"]), cuda(g["b"]), cuda(g["w"])
    close(cls("mean")(a, b), g["loss_%s_mean" % kind])
    close(cls("sum")(cuda(g["a_xywh"]), cuda(g["b_xywh"]), mode="xywh"), g["loss_%s_sum_xywh" % kind])
    close(cls("mean")(a, b, weights=w), g["loss_%s_mean_w" % kind])


def test_bce_golden(golden_iou):
    g = golden_iou
    lg, idx = cuda(g["bce_logits"]), cuda(g["bce_idx"])
    close(fl.BiCrossEntropyLoss("mean")(lg, idx), g["bce_mean"])
    close(fl.BiCrossEntropyLoss("mean")(lg.sigmoid(), idx, already_sigmoid=True), g["bce_mean_sig"])
    close(fl.BiCrossEntropyLoss("sum")(lg, idx), g["bce_sum"])
    close(fl.BiCrossEntropyLoss("mean")(cuda(g["bce1_logits"]), cuda(g["bce1_tgt"])), g["bce1_mean"])


# ---------------------------------------------------------------- NMS
def test_nms_lib_golden(golden_yolo):
    g = golden_yolo
    res = cuda(g["decoded"])
    for i in range(res.size(0)):
        for tag, (ct, it, md) in {"a": (0.25, 0.45, 300), "b": (0.05, 0.3, 20)}.items():
            s, c, b = ft.non_max_suppression(res[i], ct, it, md)
            assert np.array_equal(s.cpu().numpy(), g["nms_%s_s%d" % (tag, i)])
            assert np.array_equal(c.cpu().numpy(), g["nms_%s_c%d" % (tag, i)])
            assert np.array_equal(b.cpu().numpy(), g["nms_%s_b%d" % (tag, i)])
            assert c.dtype == torch.int64


def test_nms_empty_and_flattened_batch(golden_yolo):
    s, c, b = ft.non_max_suppression(torch.zeros(50, 9, device="cuda"), 0.25, 0.45, 300)
    assert s.shape == (0, 1) and c.shape == (0, 1) and b.shape == (0, 4) and not s.is_cuda and s.dtype == torch.float32
    res = cuda(golden_yolo["decoded"])                      # [B,N,K] is flattened across images (NMS.py:7-8)
    s, c, b = ft.non_max_suppression(res, 0.05, 0.3, 50)
    ws, wc, wb = on.nms_lib(res.cpu(), 0.05, 0.3, 50)
    assert np.array_equal(s.cpu().numpy(), ws.numpy()) and np.array_equal(c.cpu().numpy(), wc.numpy())
    assert np.array_equal(b.cpu().numpy(), wb.numpy())


def test_nms_demo_flavours_golden(golden_yolo):
    from fastvision_b200.detection.tools.nms import non_max_suppression_demo, non_max_suppression_batch
    g = golden_yolo
    res = cuda(g["decoded"])
    for i in range(res.size(0)):
        got = non_max_suppression_demo(cuda(g["demo_in%d" % i]), 0.1, 0.3, 50)
        assert np.array_equal(got.cpu().numpy(), g["demo_nms%d" % i])
    outs = non_max_suppression_batch([res[i] for i in range(res.size(0))], 0.1, 0.3, 50)
    for i, o in enumerate(outs):
        assert np.array_equal(o.numpy(), g["demo_batch%d" % i])


@pytest.mark.parametrize("case", ["cluster", "ties", "degenerate", "gap", "single", "rpn_like"])
def test_nms_segmented_vs_torchvision_golden(golden_nms, case):
    g = golden_nms
    keep = ft.nms(cuda(g[case + "_boxes"]), cuda(g[case + "_scores"]), float(g[case + "_thr"]))
    assert np.array_equal(keep.cpu().numpy(), g[case + "_keep"])


def test_nms_segmented_batched_and_large():
    gen = torch.Generator().manual_seed(77)
    sizes = [0, 1, 33, 700, 3000, 64]                       # 3000 > shared-memory capacity -> workspace path
    boxes, scores = [], []
    for n in sizes:
        xy = torch.rand(n, 2, generator=gen) * 200
        boxes.append(torch.cat([xy, xy + torch.rand(n, 2, generator=gen) * 40 + 1], 1))
        scores.append(torch.rand(n, generator=gen))
    off = torch.tensor(np.concatenate([[0], np.cumsum(sizes)]), dtype=torch.int32)
    keep, cnt = ft.nms(torch.cat(boxes).cuda(), torch.cat(scores).cuda(), 0.5, seg_offsets=off.cuda(), max_keep=3000)
    for i, n in enumerate(sizes):
        want = on.nms_greedy(boxes[i], scores[i], 0.5).numpy()
        got = keep[i, :int(cnt[i])].cpu().numpy()
        assert np.array_equal(got, want), (i, n)


def test_nms_single_segment_keeps_more_than_shared_memory_holds():
    """``nms(boxes, scores, thr)`` is the torchvision call: it returns EVERY survivor.  12 000 sparse boxes keep ~10 000 -- far
    more than the kept list's shared-memory capacity -- and 20 000 clustered ones exercise the long sort + long keep together."""
    gen = torch.Generator().manual_seed(77)
    for n, spread, thr in ((12000, 4000.0, 0.5), (20000, 600.0, 0.6)):
        xy = torch.rand(n, 2, generator=gen) * spread
        boxes = torch.cat([xy, xy + torch.rand(n, 2, generator=gen) * 40 + 2], 1)
        scores = torch.rand(n, generator=gen)
        want = on.nms_greedy(boxes, scores, thr)
        got = ft.nms(boxes.cuda(), scores.cuda(), thr).cpu()
        assert got.numel() == want.numel() and want.numel() > 4096, (n, got.numel(), want.numel())
        assert torch.equal(got, want)


def test_nms_segmented_sort_boundaries_and_score_ties():
    """Every size class of the in-kernel sort (bitonic network: 64..1024 keys at 2 per thread, 2048 at 4 per thread; radix
    beyond) with scores drawn from a handful of values, so the order is decided by the tie rule (lower index first)."""
    gen = torch.Generator().manual_seed(91)
    sizes = [1, 2, 63, 64, 65, 127, 128, 129, 511, 512, 513, 1023, 1024, 1025, 1400, 2047, 2048, 2049]
    boxes, scores = [], []
    for n in sizes:
        xy = torch.rand(n, 2, generator=gen) * 300
        boxes.append(torch.cat([xy, xy + torch.rand(n, 2, generator=gen) * 30 + 1], 1))
        scores.append(torch.randint(0, 7, (n,), generator=gen).float() / 8 + 0.125)
    off = torch.tensor(np.concatenate([[0], np.cumsum(sizes)]), dtype=torch.int32)
    keep, cnt = ft.nms(torch.cat(boxes).cuda(), torch.cat(scores).cuda(), 0.5, seg_offsets=off.cuda(), max_keep=2049)
    for i, n in enumerate(sizes):
        want = on.nms_greedy(boxes[i], scores[i], 0.5).numpy()
        got = keep[i, :int(cnt[i])].cpu().numpy()
        assert np.array_equal(got, want), (i, n)


def test_nms_batched_config1_exact_on_same_decoded_tensor():
    cfg, batch = synth.COCO416, 8
    g = synth.make_generator(1)
    labels = synth.make_labels(cfg, batch, g)
    heads = [h.cuda() for h in synth.make_heads(cfg, batch, labels, g)]
    ctx = DecodeContext(heads, cfg.anchors_levels(), cfg.strides)
    res = yolov3_decode(heads, cfg.anchors_levels(), cfg.strides, ctx=ctx, conf_thres=0.25)
    boxes, scores, cls, cnt, rows = ft.non_max_suppression_batched(res, 0.25, 0.45, 300, cand_bitmap=ctx.bitmap(), cand_records=ctx.records(), want_rows=True)
    assert int(ctx.bitmap().abs().sum()) == 0                # consumed words are cleared for the next step
    res_cpu = res.cpu()
    for i in range(batch):
        ws, wc, wb = on.nms_lib(res_cpu[i], 0.25, 0.45, 300)
        k = int(cnt[i])
        assert k == ws.size(0)
        assert np.array_equal(scores[i, :k].cpu().numpy(), ws.view(-1).numpy())
        assert np.array_equal(cls[i, :k].cpu().numpy(), wc.view(-1).numpy())
        assert np.array_equal(boxes[i, :k].cpu().numpy(), wb.numpy())
    # stand-alone path (kernel scans the objectness channel itself) gives the same
    b2, s2, c2, n2 = ft.non_max_suppression_batched(res, 0.25, 0.45, 300)
    assert torch.equal(n2, cnt)
    for i in range(batch):
        k = int(cnt[i])
        assert torch.equal(s2[i, :k], scores[i, :k]) and torch.equal(b2[i, :k], boxes[i, :k]) and torch.equal(c2[i, :k], cls[i, :k])


def test_nms_stress_many_candidates_workspace_path():
    cfg, batch = synth.COCO416, 2
    g = synth.make_generator(1, rank=3)
    labels = synth.make_labels(cfg, batch, g)
    heads = synth.make_heads(cfg, batch, labels, g, stress=True)   # ~4600 candidates / image > 2048 smem slots
    res = yolov3_decode([h.cuda() for h in heads], cfg.anchors_levels(), cfg.strides)
    boxes, scores, cls, cnt = ft.non_max_suppression_batched(res, 0.25, 0.45, 300)
    res_cpu = res.cpu()
    for i in range(batch):
        assert int((res_cpu[i, :, 4] > 0.25).sum()) > 2048
        ws, wc, wb = on.nms_lib(res_cpu[i], 0.25, 0.45, 300)
        k = int(cnt[i])
        assert k == ws.size(0)
        assert np.array_equal(scores[i, :k].cpu().numpy(), ws.view(-1).numpy())
        assert np.array_equal(boxes[i, :k].cpu().numpy(), wb.numpy())


# ---------------------------------------------------------------- loss
class _Model:
    def __init__(self, cfg):
        self.anchors_per_level = cfg.anchors_levels()
        self.backbone_strides_per_level = cfg.strides


def test_loss_golden(golden_yolo):
    g = golden_yolo
    lossf = fl.Yolov3Loss(_Model(SMALL), 0.5, 0.05, 1.0, 0.5)
    heads, labels = [h.cuda() for h in _heads(g)], cuda(g["labels"])
    out = lossf(heads, labels)
    assert out.shape == (1,)
    close(out, g["loss"])
    close(lossf(heads, labels[:0]), g["loss_nolabels"])
    close(lossf([h.cuda() for h in _heads(g, "s_head")], cuda(g["s_labels"])), g["s_loss"])
    # shuffled label order: grouped-by-image fast path off, same "last (t,a) wins" semantics as the oracle
    perm = torch.randperm(labels.size(0), generator=torch.Generator().manual_seed(0))
    want = ol.yolov3_loss(_heads(g), T(g["labels"])[perm], SMALL.anchors_levels(), SMALL.strides)
    close(lossf(heads, labels[perm.cuda()]), want)


def test_build_target_golden(golden_yolo):
    g = golden_yolo
    lossf = fl.Yolov3Loss(_Model(SMALL), 0.5, 0.05, 1.0, 0.5)
    locs, cats, xywh, anchs = lossf.build_target([h.cuda() for h in _heads(g)], cuda(g["labels"]))
    for l in range(3):
        assert np.array_equal(locs[l][0].cpu().numpy(), g["bt_b%d" % l])
        assert np.array_equal(locs[l][1].cpu().numpy(), g["bt_gxy%d" % l])
        assert np.array_equal(locs[l][2].cpu().numpy(), g["bt_a%d" % l])
        assert np.array_equal(cats[l].cpu().numpy(), g["bt_cls%d" % l])
        close(xywh[l], g["bt_xywh%d" % l])
        close(anchs[l], g["bt_anc%d" % l])


@pytest.mark.parametrize("cfg,batch", [(synth.COCO416, 8), (synth.SHIP608, 4)])
def test_loss_vs_oracle_fused_and_standalone(cfg, batch):
    g = synth.make_generator(1)
    labels = synth.make_labels(cfg, batch, g)
    heads = synth.make_heads(cfg, batch, labels, g)
    want, parts = ol.yolov3_loss(heads, labels, cfg.anchors_levels(), cfg.strides, return_partials=True)
    lossf = fl.Yolov3Loss(_Model(cfg), 0.5, 0.05, 1.0, 0.5)
    dh, dl = [h.cuda() for h in heads], labels.cuda()
    close(lossf(dh, dl), want)                                   # stand-alone (streams channel 4 itself)
    close(lossf.partials, np.asarray(parts), rtol=1e-5)
    ctx = DecodeContext(dh, cfg.anchors_levels(), cfg.strides)
    yolov3_decode(dh, cfg.anchors_levels(), cfg.strides, ctx=ctx, want_bce0=True)
    close(lossf(dh, dl, conf_bce0=ctx.bce0(), ctx=ctx), want)    # fused with decode
    # sharding: partials of two half batches add up to the whole; combine reproduces the scalar
    half = batch // 2
    p = torch.zeros(3, 4, dtype=torch.float64, device="cuda")
    for lo, hi in ((0, half), (half, batch)):
        lossf([h[lo:hi].contiguous() for h in dh], shard_labels(dl, lo, hi))
        p += lossf.partials
    close(p, np.asarray(parts), rtol=1e-6)
    full_ctx = DecodeContext(dh, cfg.anchors_levels(), cfg.strides)
    close(lossf.combine(p, batch, ctx=full_ctx), want)


# ---------------------------------------------------------------- mAP
def test_map_golden(golden_map):
    g = golden_map
    est = CalculateMAP(g["thr"])
    for i in range(int(g["n_img"])):
        before = len(est.correct_all_images)
        est.process_one(cuda(g["pred%d" % i]), cuda(g["true%d" % i]))
        got = est.correct_all_images[-1] if len(est.correct_all_images) > before else np.zeros((0, 12))
        assert np.array_equal(got, g["correct%d" % i]), i
    m_iou, m_cls, ids = est.fetch()
    np.testing.assert_allclose(m_iou, g["map_each_iou"], rtol=1e-12)
    np.testing.assert_allclose(m_cls, g["map_each_cls"], rtol=1e-12)
    assert ids == g["cls_ids"].tolist()
    # one batched launch over all images gives the same bits
    dets = [T(g["pred%d" % i]) for i in range(int(g["n_img"]))]
    gts = [T(g["true%d" % i]) for i in range(int(g["n_img"]))]
    doff = torch.tensor(np.concatenate([[0], np.cumsum([d.size(0) for d in dets])]), dtype=torch.int32)
    goff = torch.tensor(np.concatenate([[0], np.cumsum([t.size(0) for t in gts])]), dtype=torch.int32)
    est2 = CalculateMAP(g["thr"])
    correct = est2.match(torch.cat(dets).cuda(), doff.cuda(), torch.cat(gts).cuda(), goff.cuda()).cpu().numpy()
    want = np.concatenate([g["correct%d" % i][:, 2:] for i in range(int(g["n_img"]))]).astype(np.uint8)
    assert np.array_equal(correct, want)


# ---------------------------------------------------------------- pipeline (config 1, B=8) + properties at B=256
def test_val_step_config1_matches_oracle_and_graph_replay():
    cfg, batch = synth.COCO416, 8
    g = synth.make_generator(1)
    labels = synth.make_labels(cfg, batch, g)
    heads = synth.make_heads(cfg, batch, labels, g)
    step = ValStep(cfg.anchors_levels(), cfg.strides)
    dh, dl = [h.cuda() for h in heads], labels.cuda()
    out = step(dh, dl)
    torch.cuda.synchronize()
    want_res = oracle.decode.decode(heads, cfg.anchors_levels(), cfg.strides)
    close(out["results"], want_res)
    close(out["loss"], ol.yolov3_loss(heads, labels, cfg.anchors_levels(), cfg.strides))
    res_cpu = out["results"].cpu()
    est_g, est_o = CalculateMAP(np.linspace(0.5, 0.95, 10)), oracle.map_.MapOracle(np.linspace(0.5, 0.95, 10))
    dets = step.detections()
    for i in range(batch):
        ws, wc, wb = on.nms_lib(res_cpu[i], 0.25, 0.45, 300)
        assert np.array_equal(dets[i][:, 1].cpu().numpy(), ws.view(-1).numpy())
        assert np.array_equal(dets[i][:, 2:].cpu().numpy(), wb.numpy())
        tgt = synth.labels_to_pixel_targets(labels, i, cfg.img, cfg.img)
        est_o.process_one(torch.cat([wc.float(), ws, wb], 1), tgt)
        est_g.process_one(dets[i], tgt.cuda())
    mo, mg = est_o.fetch(), est_g.fetch()
    np.testing.assert_allclose(mg[0], mo[0], rtol=1e-12)
    assert mg[2] == mo[2]
    snap = {k: v.clone() for k, v in out.items()}
    replay = step.capture(dh, dl)
    for v in out.values():
        v.zero_()
    replay()
    torch.cuda.synchronize()
    for k in ("results", "loss", "cnt"):
        assert torch.equal(out[k], snap[k]), k
    for i in range(batch):
        kk = int(snap["cnt"][i])
        assert torch.equal(out["boxes"][i, :kk], snap["boxes"][i, :kk])


@pytest.mark.parametrize("cfg,batch", [(synth.COCO416, 8), (SMALL, 3), (synth.SHIP608, 2)], ids=["coco416", "tiny", "ship608"])
def test_two_part_loss_equals_one_call(cfg, batch):
    """fvb_yolov3_loss_match_f32 (beside the decode) + fvb_yolov3_loss_finish_f32 (after it) == fvb_yolov3_loss_f32: same scalar,
    fp64 partials equal up to the association of the sums, with labels and without."""
    g = synth.make_generator(1)
    labels = synth.make_labels(cfg, batch, g)
    dh = [h.cuda() for h in synth.make_heads(cfg, batch, labels, g)]
    step = ValStep(cfg.anchors_levels(), cfg.strides)
    for dl in (labels.cuda(), labels[:1].cuda(), labels[:0].cuda()):
        o = step(dh, dl)                                             # two-part form (ValStep._run)
        torch.cuda.synchronize()
        l2, p2 = o["loss"].clone(), o["partials"].clone()
        step._decode(dh)                                             # one-call form on the same decode side outputs
        step._nms()
        l1 = step.loss_fn(dh, dl, conf_bce0=step.ctx.bce0(), ctx=step.ctx)
        torch.cuda.synchronize()
        close(l2, l1, rtol=2e-7, atol=0)
        np.testing.assert_allclose(p2.cpu().numpy(), step.loss_fn.partials.cpu().numpy(), rtol=1e-13, atol=0)


def test_full_size_properties_b256():
    """BASELINE config 2 (B=256): size-independent properties instead of a full CPU oracle run."""
    cfg, batch = synth.COCO416, 256
    g = synth.make_generator(2)
    labels = synth.make_labels(cfg, batch, g)
    heads = synth.make_heads(cfg, batch, labels, g)
    step = ValStep(cfg.anchors_levels(), cfg.strides)
    dh, dl = [h.cuda() for h in heads], labels.cuda()
    out = {k: v.clone() for k, v in step(dh, dl).items()}
    # (1) batch independence: images 0..7 decoded / suppressed alone give the same bits as inside the batch
    sub = ValStep(cfg.anchors_levels(), cfg.strides)
    o8 = sub([h[:8].contiguous() for h in dh], shard_labels(dl, 0, 8))
    assert torch.equal(o8["results"], out["results"][:8])
    assert torch.equal(o8["cnt"], out["cnt"][:8])
    for i in range(8):
        k = int(o8["cnt"][i])
        assert torch.equal(o8["boxes"][i, :k], out["boxes"][i, :k])
    # (2) oracle spot check on 4 images spread over the batch
    res_cpu = out["results"][[0, 85, 170, 255]].cpu()
    for j, i in enumerate([0, 85, 170, 255]):
        ws, wc, wb = on.nms_lib(res_cpu[j], 0.25, 0.45, 300)
        k = int(out["cnt"][i])
        assert k == ws.size(0)
        assert np.array_equal(out["boxes"][i, :k].cpu().numpy(), wb.numpy())
    # (3) NMS invariants for every image: scores sorted, no kept pair above the threshold
    cnt = out["cnt"].cpu()
    for i in range(0, batch, 17):
        k = int(cnt[i])
        s = out["scores"][i, :k]
        assert bool((s[:-1] >= s[1:]).all())
        iou = ft.cal_iou_batch(out["boxes"][i, :k].contiguous(), out["boxes"][i, :k].contiguous())
        iou.fill_diagonal_(0)
        assert float(iou.max()) <= 0.45 + 1e-6
    # (4) loss linearity: per-level partial sums of 4 shards add up to the full-batch partials
    p = torch.zeros(3, 4, dtype=torch.float64, device="cuda")
    lossf = fl.Yolov3Loss(_Model(cfg), 0.5, 0.05, 1.0, 0.5)
    for lo in range(0, batch, 64):
        lossf([h[lo:lo + 64].contiguous() for h in dh], shard_labels(dl, lo, lo + 64))
        p += lossf.partials
    close(p, out["partials"], rtol=1e-9, atol=0)
    close(lossf.combine(p, batch, ctx=step.ctx), out["loss"])


def test_pipeline_overlapped_batches_match_serial_steps():
    """ValPipeline (tail of batch i overlaps the decode of batch i+1, two buffer sets) == one ValStep per batch."""
    cfg, batch = synth.COCO416, 6
    batches = []
    for r in range(5):
        g = synth.make_generator(1, rank=10 + r)
        labels = synth.make_labels(cfg, batch, g)
        batches.append(([h.cuda() for h in synth.make_heads(cfg, batch, labels, g)], labels.cuda()))
    ref = ValStep(cfg.anchors_levels(), cfg.strides)
    want = []
    for dh, dl in batches:
        o = ref(dh, dl)
        torch.cuda.synchronize()
        want.append({k: v.clone() for k, v in o.items()})
    pipe = ValPipeline(cfg.anchors_levels(), cfg.strides)
    got = []
    for i, (dh, dl) in enumerate(batches):
        o = pipe.submit(dh, dl)
        if i >= 1:                                   # read batch i-1 (still valid: its slot is reused at i+1)
            pipe.wait((i - 1) % 2)
            got.append({k: v.clone() for k, v in pipe.steps[(i - 1) % 2].out.items()})
    pipe.flush()
    got.append({k: v.clone() for k, v in o.items()})
    torch.cuda.synchronize()
    for w, g_ in zip(want, got):
        assert torch.equal(w["results"], g_["results"]) and torch.equal(w["cnt"], g_["cnt"])
        close(g_["loss"], w["loss"], rtol=1e-7, atol=0)
        for i in range(batch):
            k = int(w["cnt"][i])
            assert torch.equal(w["boxes"][i, :k], g_["boxes"][i, :k]) and torch.equal(w["cls"][i, :k], g_["cls"][i, :k])


def test_nms_frcnn_flavour_golden():
    """demos/faster_rcnn/utils/nms.py:5-39 (class-tagged rows [x1,y1,x2,y2,cat,score])."""
    from conftest import load_golden
    g = load_golden("frcnn_nms.npz")
    for tag in "abc":
        thr, iou, md = g[tag + "_cfg"]
        out = ft.non_max_suppression_frcnn(cuda(g[tag + "_pred"]), float(thr), float(iou), int(md))
        assert tuple(out.shape) == g[tag + "_out"].shape, tag
        assert np.array_equal(out.cpu().numpy(), g[tag + "_out"]), tag


@pytest.mark.parametrize("cfg,batch", [(synth.COCO416, 5), (synth.SHIP608, 3)])
def test_decode_nchw_heads(cfg, batch):
    """SURVEY 8 a1': the conv outputs [B,A*K,H,W] decoded directly == decoding their permuted copies (bitwise), and
    == the demos' own formulation (feature-unit anchors x stride; (a,y,x) and (y,x,a) row orders) to rtol 1e-5."""
    g = synth.make_generator(3, rank=7)
    labels = synth.make_labels(cfg, batch, g)
    heads = synth.make_heads(cfg, batch, labels, g)                               # [B,A,H,W,K]
    nchw = [h.permute(0, 1, 4, 2, 3).reshape(h.size(0), -1, h.size(2), h.size(3)).contiguous() for h in heads]
    anc, st = cfg.anchors_levels(), cfg.strides
    want = yolov3_decode([h.cuda() for h in heads], anc, st)
    ctx = DecodeContext([h.cuda() for h in nchw], anc, st, layout="nchw")
    got = yolov3_decode([h.cuda() for h in nchw], anc, st, layout="nchw", ctx=ctx, conf_thres=0.25, want_bce0=True)
    assert torch.equal(got, want)
    ctx0 = DecodeContext([h.cuda() for h in heads], anc, st)
    yolov3_decode([h.cuda() for h in heads], anc, st, ctx=ctx0, conf_thres=0.25, want_bce0=True)
    assert torch.equal(ctx.bitmap(), ctx0.bitmap()) and torch.equal(ctx.bce0(), ctx0.bce0())
    anc_feat = [a.view(-1, 2) / s for a, s in zip(anc, st)]
    close(got, oracle.decode.decode_demo_nchw(nchw, anc_feat, st, order="ayx"))
    got_yxa = yolov3_decode([h.cuda() for h in nchw], anc, st, layout="nchw", row_order="yxa")
    close(got_yxa, oracle.decode.decode_demo_nchw(nchw, anc_feat, st, order="yxa"))



def test_full_size_properties_config3_ship608():
    """BASELINE configs[2] at its 8-GPU per-rank share (YOLOv3-608, 10 classes, 128 images, K = 15 channels -- the narrow-row
    decode path, ~1400 NMS candidates per image): batch independence, oracle spot checks of decode / NMS / loss, NMS invariants,
    and the sharded loss (4 ranks' partials) against the full batch."""
    cfg, batch = synth.SHIP608, 128
    g = synth.make_generator(3)
    labels = synth.make_labels(cfg, batch, g)
    heads = synth.make_heads(cfg, batch, labels, g)
    step = ValStep(cfg.anchors_levels(), cfg.strides)
    dh, dl = [h.cuda() for h in heads], labels.cuda()
    out = {k: v.clone() for k, v in step(dh, dl).items()}
    sub = ValStep(cfg.anchors_levels(), cfg.strides)
    o4 = sub([h[60:64].contiguous() for h in dh], shard_labels(dl, 60, 64))
    assert torch.equal(o4["results"], out["results"][60:64]) and torch.equal(o4["cnt"], out["cnt"][60:64])
    for i in range(4):
        k = int(o4["cnt"][i])
        assert torch.equal(o4["boxes"][i, :k], out["boxes"][60 + i, :k])
    pick = [0, 63, 127]
    want_res = oracle.decode.decode([h[pick] for h in heads], cfg.anchors_levels(), cfg.strides)
    close(out["results"][pick], want_res)
    res_cpu = out["results"][pick].cpu()
    for j, i in enumerate(pick):
        ws, wc, wb = on.nms_lib(res_cpu[j], 0.25, 0.45, 300)
        k = int(out["cnt"][i])
        assert k == ws.size(0)
        assert np.array_equal(out["boxes"][i, :k].cpu().numpy(), wb.numpy())
        assert np.array_equal(out["cls"][i, :k].cpu().numpy(), wc.view(-1).numpy())
    for i in range(0, batch, 13):
        k = int(out["cnt"][i])
        s = out["scores"][i, :k]
        assert bool((s[:-1] >= s[1:]).all())
        iou = ft.cal_iou_batch(out["boxes"][i, :k].contiguous(), out["boxes"][i, :k].contiguous())
        iou.fill_diagonal_(0)
        assert float(iou.max()) <= 0.45 + 1e-6
    sl = slice(0, 16)
    sub_labels = shard_labels(labels, 0, 16)
    want_loss = ol.yolov3_loss([h[sl] for h in heads], sub_labels, cfg.anchors_levels(), cfg.strides)
    lossf = fl.Yolov3Loss(_Model(cfg), 0.5, 0.05, 1.0, 0.5)
    close(lossf([h[sl].contiguous() for h in dh], sub_labels.cuda()), want_loss)
    p = torch.zeros(3, 4, dtype=torch.float64, device="cuda")
    for lo in range(0, batch, 32):
        lossf([h[lo:lo + 32].contiguous() for h in dh], shard_labels(dl, lo, lo + 32))
        p += lossf.partials
    close(p, out["partials"], rtol=1e-9, atol=0)
    close(lossf.combine(p, batch, ctx=step.ctx), out["loss"])


def test_full_size_properties_config3_b1024():
    """BASELINE configs[2] at its FULL global batch (YOLOv3-608, 10 classes, B = 1024: 1.4 GB of heads, 23.3 M rows): what a single
    GPU computes for the whole batch must be, image for image, what the 128-image shards of the 8-GPU run compute -- detections
    bit-equal, loss partials additive over the shards -- plus oracle spot checks (decode rows, keep lists) on images of the first,
    a middle and the last shard, and the NMS invariants on a stride of images."""
    cfg, batch, shard = synth.SHIP608, 1024, 128
    g = synth.make_generator(3)
    labels = synth.make_labels(cfg, batch, g)
    heads = synth.make_heads(cfg, batch, labels, g)
    dh, dl = [h.cuda() for h in heads], labels.cuda()
    full = ValStep(cfg.anchors_levels(), cfg.strides)
    out = full(dh, dl)
    torch.cuda.synchronize()
    assert int(out["cnt"].min()) >= 0
    part = ValStep(cfg.anchors_levels(), cfg.strides)
    psum = torch.zeros(3, 4, dtype=torch.float64, device="cuda")
    for lo in range(0, batch, shard):
        o = part([h[lo:lo + shard].contiguous() for h in dh], shard_labels(dl, lo, lo + shard))
        psum += o["partials"]
        assert torch.equal(o["cnt"], out["cnt"][lo:lo + shard]), lo
        valid = torch.arange(o["boxes"].size(1), device="cuda")[None, :] < o["cnt"][:, None]
        for key in ("boxes", "scores", "cls"):
            assert torch.equal(o[key][valid], out[key][lo:lo + shard][valid]), (lo, key)
        if lo in (0, 512):
            assert torch.equal(o["results"], out["results"][lo:lo + shard])
    close(psum, out["partials"], rtol=1e-9, atol=0)                       # loss/yolov3_loss.py:52,58,64: sums are additive
    close(full.loss_fn.combine(psum, batch, ctx=full.ctx), out["loss"])
    pick = [5, 517, 1023]
    want_res = oracle.decode.decode([h[pick] for h in heads], cfg.anchors_levels(), cfg.strides)
    close(out["results"][pick], want_res)
    res_cpu = out["results"][pick].cpu()
    for j, i in enumerate(pick):
        ws, wc, wb = on.nms_lib(res_cpu[j], 0.25, 0.45, 300)
        k = int(out["cnt"][i])
        assert k == ws.size(0)
        assert np.array_equal(out["boxes"][i, :k].cpu().numpy(), wb.numpy())
        assert np.array_equal(out["cls"][i, :k].cpu().numpy(), wc.view(-1).numpy())
    for i in range(0, batch, 97):
        k = int(out["cnt"][i])
        s = out["scores"][i, :k]
        assert bool((s[:-1] >= s[1:]).all())
        iou = ft.cal_iou_batch(out["boxes"][i, :k].contiguous(), out["boxes"][i, :k].contiguous())
        iou.fill_diagonal_(0)
        assert float(iou.max()) <= 0.45 + 1e-6
