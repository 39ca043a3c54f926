"""GPU parity of the backward kernels: gradients through the C ABI vs gradients recorded from the
reference's own autograd graph (tests/golden/grads_small.npz) and vs autograd through the CPU oracle on seeded inputs.

Tolerance: rtol 1e-5 (the north-star fp32 tolerance) with an absolute floor of 2e-6 of the largest gradient entry --
a gradient entry is a sum of signed terms, so the floor is relative to the tensor, not to the entry.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import oracle
from conftest import T
from gpu_util import cuda, close
from small_cfg import SMALL
from fastvision_b200 import synth
from fastvision_b200 import loss as fl
from fastvision_b200.pipeline import shard_labels


def gclose(got, want, floor=2e-6):
    want = want.detach().cpu().numpy() if isinstance(want, torch.Tensor) else np.asarray(want)
    close(got, want, rtol=1e-5, atol=floor * float(np.abs(want).max()))


class _Model:
    def __init__(self, cfg):
        self.anchors_per_level = cfg.anchors_levels()
        self.backbone_strides_per_level = cfg.strides


def test_yolov3_loss_backward_golden(golden_grads):
    """loss.backward() on the drop-in module == the reference's autograd gradients (duplicate cell, clamped cell included)."""
    g = golden_grads
    lossf = fl.Yolov3Loss(_Model(SMALL), 0.5, 0.05, 1.0, 0.5)
    heads = [cuda(g["head%d" % i]).requires_grad_(True) for i in range(3)]
    labels = cuda(g["labels"])
    loss = lossf(heads, labels)
    close(loss, g["loss"])
    (loss * float(g["upstream"])).sum().backward()
    for i in range(3):
        gclose(heads[i].grad, g["grad%d" % i])
    # no labels: only the objectness term
    heads = [cuda(g["head%d" % i]).requires_grad_(True) for i in range(3)]
    lossf(heads, labels[:0]).sum().backward()
    for i in range(3):
        gclose(heads[i].grad, g["grad_nolabels%d" % i])
    # shuffled labels (grouped-by-image fast path off): same gradient semantics as the oracle's (t,a) order
    perm = torch.randperm(labels.size(0), generator=torch.Generator().manual_seed(0))
    _, want = oracle.grad.yolov3_loss_grad([T(g["head%d" % i]) for i in range(3)], T(g["labels"])[perm],
                                           SMALL.anchors_levels(), SMALL.strides)
    heads = [cuda(g["head%d" % i]).requires_grad_(True) for i in range(3)]
    lossf(heads, labels[perm.cuda()]).sum().backward()
    for i in range(3):
        gclose(heads[i].grad, want[i])


@pytest.mark.parametrize("cfg,batch", [(synth.COCO416, 4), (synth.SHIP608, 3)])
def test_yolov3_loss_backward_vs_oracle(cfg, batch):
    g = synth.make_generator(1)
    labels = synth.make_labels(cfg, batch, g)
    # force duplicate cells: repeat some labels with a small jitter so several matches share a cell/anchor
    dup = labels[::3].clone()
    dup[:, 2:4] += 1e-4
    dup[:, 1] = (dup[:, 1] + 1) % cfg.num_classes
    labels = torch.cat([labels, dup], 0)
    labels = labels[torch.argsort(labels[:, 0], stable=True)]
    heads = synth.make_heads(cfg, batch, labels, g)
    _, want = oracle.grad.yolov3_loss_grad(heads, labels, cfg.anchors_levels(), cfg.strides, upstream=0.5)
    lossf = fl.Yolov3Loss(_Model(cfg), 0.5, 0.05, 1.0, 0.5)
    dh = [h.cuda().requires_grad_(True) for h in heads]
    (lossf(dh, labels.cuda()) * 0.5).sum().backward()
    for i in range(3):
        gclose(dh[i].grad, want[i])
    # run-to-run reproducibility (single writer per row, no atomics)
    first = [h.grad.clone() for h in dh]
    for h in dh:
        h.grad = None
    (lossf(dh, labels.cuda()) * 0.5).sum().backward()
    for i in range(3):
        assert torch.equal(first[i], dh[i].grad)
    # data-parallel form: two half batches with the global partials / batch reproduce the full-batch gradient
    half = batch // 2
    dl = labels.cuda()
    with torch.no_grad():
        lossf([h.detach() for h in dh], dl)
        parts = lossf.partials.clone()
    for lo, hi in ((0, half), (half, batch)):
        shard = [h.detach()[lo:hi].contiguous() for h in dh]
        grads = lossf.backward_heads(shard, shard_labels(dl, lo, hi), torch.full((1,), 0.5, device="cuda"), parts, batch)
        for i in range(3):
            gclose(grads[i], want[i][lo:hi])


def test_yolov3_loss_backward_unaligned_views():
    """Gradient buffers that are not 16-byte aligned take the scalar store path."""
    cfg, batch = SMALL, 2
    g = synth.make_generator(3)
    labels = synth.make_labels(cfg, batch, g)
    heads = synth.make_heads(cfg, batch, labels, g)
    _, want = oracle.grad.yolov3_loss_grad(heads, labels, cfg.anchors_levels(), cfg.strides)
    lossf = fl.Yolov3Loss(_Model(cfg), 0.5, 0.05, 1.0, 0.5)
    dh = [h.cuda() for h in heads]
    with torch.no_grad():
        lossf(dh, labels.cuda())
    bufs = [torch.zeros(h.numel() + 1, device="cuda") for h in dh]
    grads = [b[1:].view_as(h) for b, h in zip(bufs, dh)]
    lossf.backward_heads(dh, labels.cuda(), None, lossf.partials, batch, grads=grads)
    for i in range(3):
        gclose(grads[i], want[i])


@pytest.mark.parametrize("kind,cls", [("iou", fl.IOULoss), ("giou", fl.GIOULoss), ("diou", fl.DIOULoss), ("ciou", fl.CIOULoss)])
def test_iou_loss_backward_golden(golden_grads, kind, cls):
    g = golden_grads
    cases = {"xyxy_mean": ("a", "b", "xyxy", "mean", None), "xywh_sum": ("a_xywh", "b_xywh", "xywh", "sum", None),
             "xyxy_mean_w": ("a", "b", "xyxy", "mean", cuda(g["w"]))}
    for tag, (x, y, mode, red, w) in cases.items():
        xx, yy = cuda(g[x]).requires_grad_(True), cuda(g[y]).requires_grad_(True)
        cls(red)(xx, yy, weights=w, mode=mode).backward()
        gclose(xx.grad, g["g_%s_%s_pre" % (kind, tag)])
        gclose(yy.grad, g["g_%s_%s_true" % (kind, tag)])
    if kind == "iou":
        xx, yy = cuda(g["a_xywh"][:, 2:]).requires_grad_(True), cuda(g["b_xywh"][:, 2:]).requires_grad_(True)
        fl.IOULoss("mean")(xx, yy, mode="wh").backward()
        gclose(xx.grad, g["g_iou_wh_mean_pre"])
        gclose(yy.grad, g["g_iou_wh_mean_true"])
    # only y_pre needs a gradient: the y_true output is skipped
    xx = cuda(g["a"]).requires_grad_(True)
    cls("mean")(xx, cuda(g["b"])).backward()
    gclose(xx.grad, g["g_%s_xyxy_mean_pre" % kind])


def test_bce_loss_backward_golden(golden_grads):
    g = golden_grads
    idx = cuda(g["bce_idx"])
    for red in ("mean", "sum"):
        x = cuda(g["bce_logits"]).requires_grad_(True)
        fl.BiCrossEntropyLoss(red)(x, idx).backward()
        gclose(x.grad, g["g_bce_%s" % red])
    x = cuda(g["bce_logits"]).sigmoid().detach().requires_grad_(True)
    fl.BiCrossEntropyLoss("mean")(x, idx, already_sigmoid=True).backward()
    gclose(x.grad, g["g_bce_sig_mean"], floor=1e-5)
    x = cuda(g["bce1_logits"]).requires_grad_(True)
    fl.BiCrossEntropyLoss("mean")(x, cuda(g["bce1_tgt"])).backward()
    gclose(x.grad, g["g_bce1_mean"])


def test_backward_full_size_properties_b256():
    """B=256 YOLOv3-416: the gradient is zero outside channel 4 + matched rows, channel 4 matches the closed form on a
    sample, and a sampled image agrees with the oracle run on that image alone with the global normalisers."""
    cfg, batch = synth.COCO416, 256
    g = synth.make_generator(2)
    labels = synth.make_labels(cfg, batch, g)
    heads = synth.make_heads(cfg, batch, labels, g)
    lossf = fl.Yolov3Loss(_Model(cfg), 0.5, 0.05, 1.0, 0.5)
    dh = [h.cuda().requires_grad_(True) for h in heads]
    lossf(dh, labels.cuda()).sum().backward()
    for l, h in enumerate(dh):
        gr = h.grad
        assert torch.isfinite(gr).all()
        nz_rows = (gr[..., :4].abs().sum(-1) + gr[..., 5:].abs().sum(-1)) > 0
        m = int(lossf.partials[l, 3].item())
        assert 0 < int(nz_rows.sum()) <= m          # only matched rows (duplicates share a row) carry box / class gradient
        # dense objectness gradient, target 0, on the unmatched cells: coef * p(1-p)/(1-p+1e-8)
        p = torch.sigmoid(h.detach()[..., 4].double())
        coef = 1.0 / (h.size(1) * h.size(2) * h.size(3))
        want = coef * p * (1 - p) / (1 - p + 1e-8)
        sel = ~nz_rows
        torch.testing.assert_close(gr[..., 4][sel].double(), want[sel], rtol=2e-5, atol=1e-12)


def test_backward_wide_rows_fallback_path():
    """K = 5 + 140 > 128 channels: the matched-row kernels leave the register path and read-modify-write the row."""
    cfg = synth.YoloConfig("wide", 64, 140, [[40, 30], [50, 60], [30, 50], [20, 24], [16, 10], [12, 22], [4, 6], [8, 5], [7, 9]],
                           labels_per_img=4.0, max_labels=9)
    batch = 2
    g = synth.make_generator(5)
    labels = synth.make_labels(cfg, batch, g)
    dup = labels[:2].clone()
    dup[:, 1] = (dup[:, 1] + 3) % cfg.num_classes
    labels = torch.cat([labels, dup], 0)
    labels = labels[torch.argsort(labels[:, 0], stable=True)]
    heads = synth.make_heads(cfg, batch, labels, g)
    _, want = oracle.grad.yolov3_loss_grad(heads, labels, cfg.anchors_levels(), cfg.strides)
    lossf = fl.Yolov3Loss(_Model(cfg), 0.5, 0.05, 1.0, 0.5)
    dh = [h.cuda().requires_grad_(True) for h in heads]
    lossf(dh, labels.cuda()).sum().backward()
    for i in range(3):
        gclose(dh[i].grad, want[i])
    # the demos' loss on the same wide head
    import types
    from fastvision_b200.loss import ComputeLoss
    nchw = [h.permute(0, 1, 4, 2, 3).reshape(h.size(0), -1, h.size(2), h.size(3)).contiguous() for h in heads]
    anchors = [a.reshape(-1, 2) / s for a, s in zip(cfg.anchors_levels(), cfg.strides)]
    hs = [h.clone().requires_grad_(True) for h in nchw]
    lb, lc, lo = oracle.demo_loss.compute_loss(hs, labels, anchors, "ship")
    (lb + lc + lo).sum().backward()
    dn = [h.cuda().requires_grad_(True) for h in nchw]
    gb, gc, go = ComputeLoss()(dn, labels.cuda(), types.SimpleNamespace(anchors=anchors))
    close(gb, lb); close(gc, lc); close(go, lo)
    (gb + gc + go).sum().backward()
    for i in range(3):
        gclose(dn[i].grad, hs[i].grad)
