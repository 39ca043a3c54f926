"""GPU parity of the demos' training loss ComputeLoss, forward and backward, through the C ABI:
golden vectors recorded from the reference's demos (tests/golden/demo_loss.npz) and the CPU oracle on seeded inputs.
Values: rtol 1e-5; gradients: rtol 1e-5 with a floor of 2e-6 of the largest entry."""
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import oracle
from conftest import T
from gpu_util import cuda, close
from fastvision_b200 import synth
from fastvision_b200.loss import ComputeLoss, ComputeLossU
from fastvision_b200.pipeline import shard_labels


def gclose(got, want, floor=2e-6):
    want = want.detach().cpu().numpy() if isinstance(want, torch.Tensor) else np.asarray(want)
    close(got, want, rtol=1e-5, atol=floor * float(np.abs(want).max()))


def to_nchw(h):
    b, a, hh, ww, k = h.shape
    return h.permute(0, 1, 4, 2, 3).reshape(b, a * k, hh, ww).contiguous()


def _model(anchors):
    return types.SimpleNamespace(anchors=[a.cuda() for a in anchors])


def test_demo_loss_golden(golden_demo_loss):
    g = golden_demo_loss
    anchors = [T(g["anchors%d" % i]) for i in range(3)]
    labels = cuda(g["labels"])
    heads = [cuda(g["head%d" % i]).requires_grad_(True) for i in range(3)]
    lb, lc, lo = ComputeLoss()(heads, labels, _model(anchors))
    assert lb.shape == lc.shape == lo.shape == (1,)
    close(lb, g["ship_box"]); close(lc, g["ship_cls"]); close(lo, g["ship_conf"])
    w = g["ship_up"]
    (lb * float(w[0]) + lc * float(w[1]) + lo * float(w[2])).sum().backward()
    for i in range(3):
        gclose(heads[i].grad, g["ship_grad%d" % i])
    heads = [cuda(g["head%d" % i]).requires_grad_(True) for i in range(3)]
    wrapped = types.SimpleNamespace(module=_model(anchors))         # DataParallel-style wrapper (get_model, lossv3.py:13-17)
    lu = ComputeLossU()(heads, labels, wrapped)
    assert lu.shape == (1,)
    close(lu, g["u_loss"])
    (lu * float(g["u_up"])).sum().backward()
    for i in range(3):
        gclose(heads[i].grad, g["u_grad%d" % i])
    # an image without targets raises like the reference (strict), or is treated as ignore-free
    keep = labels[labels[:, 0] != 1]
    with pytest.raises(IndexError):
        ComputeLoss()([h.detach() for h in heads], keep, _model(anchors))
    out = ComputeLoss(strict=False)([h.detach() for h in heads], keep, _model(anchors))
    assert all(torch.isfinite(o).all() for o in out)


@pytest.mark.parametrize("flavour", ["ship", "u"])
@pytest.mark.parametrize("cfg,batch", [(synth.SHIP608, 3), (synth.COCO416, 2)])
def test_demo_loss_vs_oracle(cfg, batch, flavour):
    g = synth.make_generator(4)
    labels = synth.make_labels(cfg, batch, g)
    dup = labels[::2].clone()
    dup[:, 2:4] += 1e-4                                            # duplicates: same cell and (mostly) the same best anchor
    dup[:, 1] = (dup[:, 1] + 1) % cfg.num_classes
    labels = torch.cat([labels, dup], 0)
    heads = [to_nchw(h) for h in synth.make_heads(cfg, batch, labels, g)]
    anchors = [a.reshape(-1, 2) / s for a, s in zip(cfg.anchors_levels(), cfg.strides)]
    hs = [h.clone().requires_grad_(True) for h in heads]
    want, parts = oracle.demo_loss.compute_loss(hs, labels, anchors, flavour, partials=True)
    up = [0.05, 0.5, 1.0] if flavour == "ship" else [0.7]
    tot = sum(w * o for w, o in zip(up, want)) if flavour == "ship" else want * up[0]
    tot.sum().backward()
    lossf = (ComputeLoss if flavour == "ship" else ComputeLossU)()
    dh = [h.cuda().requires_grad_(True) for h in heads]
    got = lossf(dh, labels.cuda(), _model(anchors))
    if flavour == "ship":
        for a, b in zip(got, want):
            close(a, b)
        sum(w * o for w, o in zip(up, got)).sum().backward()
    else:
        close(got, want)
        (got * up[0]).sum().backward()
    p = np.asarray(parts)                                             # oracle: [S_a, S_b, S_cls, S_conf, n_valid]
    close(lossf.partials[:, :4], p[:, :4], rtol=1e-5)
    assert np.array_equal(lossf.partials[:, 4].cpu().numpy(), p[:, 4])   # the ignore / positive mask agrees exactly
    for i in range(3):
        gclose(dh[i].grad, hs[i].grad)
    # reproducible, and data-parallel: shard partials add up, combine + backward with the global partials match
    first = [h.grad.clone() for h in dh]
    for h in dh:
        h.grad = None
    got2 = lossf(dh, labels.cuda(), _model(anchors))
    (sum(w * o for w, o in zip(up, got2)) if flavour == "ship" else got2 * up[0]).sum().backward()
    for i in range(3):
        assert torch.equal(first[i], dh[i].grad)
    if batch >= 2:
        full_parts, full_ctx = lossf.partials.clone(), lossf._ctx
        acc = torch.zeros_like(full_parts)
        dl = labels.cuda()
        shards = []
        for lo, hi in ((0, 1), (1, batch)):
            sh = [h.detach()[lo:hi].contiguous() for h in dh]
            sl = shard_labels(dl, lo, hi)
            ctx = lossf._context(sh, anchors)
            _, parts_s, mask_s = lossf._run(sh, sl, ctx, want_mask=True)
            acc += parts_s
            shards.append((lo, hi, sh, sl, mask_s, ctx))
        close(acc, full_parts, rtol=1e-9)
        comb = lossf.combine(acc, ctx=full_ctx)
        for a, b in zip(comb if flavour == "ship" else [comb], want if flavour == "ship" else [want]):
            close(a, b)
        gout = torch.tensor(up, device="cuda")
        for lo, hi, sh, sl, mask_s, ctx in shards:
            grads = lossf.backward_heads(sh, sl, gout, acc, mask_s, ctx=ctx)
            for i in range(3):
                gclose(grads[i], hs[i].grad[lo:hi])


def test_demo_loss_unaligned_grad_buffers(golden_demo_loss):
    g = golden_demo_loss
    anchors = [T(g["anchors%d" % i]) for i in range(3)]
    labels = cuda(g["labels"])
    heads = [cuda(g["head%d" % i]) for i in range(3)]
    lossf = ComputeLoss()
    ctx = lossf._context(heads, anchors)
    _, parts, mask = lossf._run(heads, labels, ctx, want_mask=True)
    bufs = [torch.zeros(h.numel() + 1, device="cuda") for h in heads]
    grads = [b[1:].view_as(h) for b, h in zip(bufs, heads)]
    lossf.backward_heads(heads, labels, cuda(g["ship_up"]), parts, mask, ctx=ctx, grads=grads)
    for i in range(3):
        gclose(grads[i], g["ship_grad%d" % i])


def test_demo_loss_edge_cases(golden_demo_loss):
    g = golden_demo_loss
    anchors = [T(g["anchors%d" % i]) for i in range(3)]
    heads = [cuda(g["head%d" % i]) for i in range(3)]
    # no targets at all: the reference cannot run (IndexError); non-strict gives mean-of-empty NaNs for the target terms and the
    # all-negative objectness mean
    none = torch.zeros(0, 6, device="cuda")
    with pytest.raises(IndexError):
        ComputeLoss()(heads, none, _model(anchors))
    lb, lc, lo = ComputeLoss(strict=False)(heads, none, _model(anchors))
    assert torch.isnan(lb).all() and torch.isnan(lc).all()
    want = sum(torch.nn.functional.binary_cross_entropy_with_logits(
        h.view(h.size(0), 3, -1, h.size(2), h.size(3))[:, :, 4].double(), torch.zeros(h.size(0), 3, h.size(2), h.size(3), dtype=torch.float64, device="cuda"))
        for h in heads)
    close(lo, want.float().view(1))
    # batch of one image, one target
    one = [h[:1].contiguous() for h in heads]
    lab = torch.tensor([[0, 1, 0.5, 0.5, 0.4, 0.3]], device="cuda")
    hs = [h.cpu().clone().requires_grad_(True) for h in one]
    wb, wc, wo = oracle.demo_loss.compute_loss(hs, lab.cpu(), anchors, "ship")
    (wb + wc + wo).sum().backward()
    dh = [h.clone().requires_grad_(True) for h in one]
    gb, gc, go = ComputeLoss()(dh, lab, _model(anchors))
    close(gb, wb); close(gc, wc); close(go, wo)
    (gb + gc + go).sum().backward()
    for a, b in zip(dh, hs):
        gclose(a.grad, b.grad)
    # CPU tensors are rejected (no fallback)
    with pytest.raises(RuntimeError, match="CUDA only"):
        ComputeLoss()([h.cpu() for h in heads], lab.cpu(), _model(anchors))


def test_demo_loss_many_targets_per_image():
    """More targets in one image than one shared-memory round holds (64): the mask state is carried across rounds."""
    cfg = synth.SHIP608
    g = synth.make_generator(9)
    n = 150
    lab = torch.cat([torch.zeros(n, 1), torch.randint(0, cfg.num_classes, (n, 1), generator=g).float(),
                     torch.rand(n, 2, generator=g) * 0.9 + 0.05, torch.rand(n, 2, generator=g) * 0.25 + 0.02], 1)
    lab = torch.cat([lab, torch.tensor([[1, 2, 0.5, 0.5, 0.3, 0.3]])], 0)
    heads = [to_nchw(h) for h in synth.make_heads(cfg, 2, lab, g)]
    anchors = [a.reshape(-1, 2) / s for a, s in zip(cfg.anchors_levels(), cfg.strides)]
    hs = [h.clone().requires_grad_(True) for h in heads]
    want, parts = oracle.demo_loss.compute_loss(hs, lab, anchors, "ship", partials=True)
    sum(want).sum().backward()
    dh = [h.cuda().requires_grad_(True) for h in heads]
    lossf = ComputeLoss()
    got = lossf(dh, lab.cuda(), _model(anchors))
    for a, b in zip(got, want):
        close(a, b)
    assert np.array_equal(lossf.partials[:, 4].cpu().numpy(), np.asarray(parts)[:, 4])
    sum(got).sum().backward()
    for i in range(3):
        gclose(dh[i].grad, hs[i].grad)
    # without autograd (no mask requested) the same values come out of the workspace-mask path
    with torch.no_grad():
        got2 = ComputeLoss()([h.detach() for h in dh], lab.cuda(), _model(anchors))
    for a, b in zip(got2, got):
        assert torch.equal(a, b.detach())
