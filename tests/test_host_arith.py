"""CPU check of the kernels' hand-derived gradient arithmetic (no GPU): tests/host_check.cu compiles the SAME
FVB_HD source the CUDA kernels use (csrc/iou_grad.cuh, csrc/loss_common.cuh) as host code; the results are compared with
torch autograd through the oracle.  The oracle is only the checker here."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest
import torch

from oracle import iou as oi
from oracle import loss as ol

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_check.cu")
SO = os.path.join(HERE, "_host_check.so")
CSRC = os.path.join(os.path.dirname(HERE), "fastvision_b200", "csrc")


@pytest.fixture(scope="module")
def hc():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    deps = [SRC] + [os.path.join(CSRC, f) for f in ("common.cuh", "iou_grad.cuh", "loss_common.cuh")]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.run([nvcc, "-O1", "-std=c++17", "-Xcompiler", "-fPIC,-ffp-contract=off", "-shared", SRC, "-o", SO], check=True)
    return C.CDLL(SO)


def fp(x):
    return x.ctypes.data_as(C.c_void_p)


def boxes(n, seed, overlap=True):
    g = torch.Generator().manual_seed(seed)
    xy = torch.rand(n, 2, generator=g) * 50
    wh = torch.rand(n, 2, generator=g) * 30 + 1
    a = torch.cat([xy, xy + wh], 1)
    b = a + torch.randn(n, 4, generator=g) * 4
    b[:, 2:] = torch.maximum(b[:, 2:], b[:, :2] + 0.5)
    if not overlap:
        b[: n // 4] += 200.0                     # disjoint pairs: clamp(0) branch
    return a, b


@pytest.mark.parametrize("kind", ["iou", "giou", "diou", "ciou"])
@pytest.mark.parametrize("mode", ["xyxy", "xywh"])
@pytest.mark.parametrize("variant", ["lib", "demo"])
def test_iou_family_grad_matches_autograd(hc, kind, mode, variant):
    if variant == "demo" and kind in ("iou", "giou"):
        pytest.skip("the demo variant only differs for DIoU/CIoU")
    n = 257
    a, b = boxes(n, 5, overlap=False)
    if mode == "xywh":
        from oracle.boxes import xyxy2xywh
        a, b = xyxy2xywh(a), xyxy2xywh(b)
    g = torch.randn(n, generator=torch.Generator().manual_seed(1))
    ta, tb = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
    fn = {"iou": oi.cal_iou, "giou": oi.GIOU, "diou": oi.DIOU, "ciou": oi.CIOU}[kind]
    kw = {"variant": variant} if kind in ("diou", "ciou") else {}
    val = fn(ta, tb, mode=mode, **kw).reshape(-1)
    (val * g).sum().backward()
    an, bn, gn = a.numpy().copy(), b.numpy().copy(), g.numpy().copy()
    value = np.empty(n, np.float32)
    ga, gb = np.empty((n, 4), np.float32), np.empty((n, 4), np.float32)
    hc.hc_iou_family(fp(an), fp(bn), n, {"xyxy": 0, "xywh": 1}[mode], {"iou": 0, "giou": 1, "diou": 2, "ciou": 3}[kind],
                     {"lib": 0, "demo": 1}[variant], C.c_float(1e-7), fp(gn), fp(value), fp(ga), fp(gb))
    assert not np.isnan(value).any(), "backward's forward value differs from iou_family<false>"
    np.testing.assert_allclose(value, val.detach().numpy(), rtol=1e-5, atol=1e-6)
    scale = max(float(ta.grad.abs().max()), float(tb.grad.abs().max()))
    np.testing.assert_allclose(ga, ta.grad.numpy(), rtol=1e-4, atol=2e-6 * scale)
    np.testing.assert_allclose(gb, tb.grad.numpy(), rtol=1e-4, atol=2e-6 * scale)


def test_tie_and_touching_conventions(hc):
    """torch.minimum/maximum split ties 1/2:1/2, clamp(0) passes the gradient at exactly 0."""
    a = torch.tensor([[0., 0., 10., 10.], [0., 0., 10., 10.], [0., 0., 10., 10.]])
    b = torch.tensor([[0., 0., 10., 10.], [10., 0., 20., 10.], [2., 0., 10., 8.]])   # identical, touching, shared edges
    for kind, fn in [("iou", oi.cal_iou), ("giou", oi.GIOU), ("diou", oi.DIOU), ("ciou", oi.CIOU)]:
        ta, tb = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
        fn(ta, tb, mode="xyxy").sum().backward()
        n = a.size(0)
        value = np.empty(n, np.float32)
        ga, gb = np.empty((n, 4), np.float32), np.empty((n, 4), np.float32)
        ones = np.ones(n, np.float32)
        hc.hc_iou_family(fp(a.numpy().copy()), fp(b.numpy().copy()), n, 0, {"iou": 0, "giou": 1, "diou": 2, "ciou": 3}[kind], 0,
                         C.c_float(1e-7), fp(ones), fp(value), fp(ga), fp(gb))
        np.testing.assert_allclose(ga, ta.grad.numpy(), rtol=1e-4, atol=1e-7)
        np.testing.assert_allclose(gb, tb.grad.numpy(), rtol=1e-4, atol=1e-7)


def test_match_row_grad_matches_autograd(hc):
    """Box logits of a matched row: w_box * d(1-CIoU) + g_tgt * d IoU, through sigmoid / exp*anchor / xywh->xyxy."""
    n = 200
    g = torch.Generator().manual_seed(3)
    r = torch.randn(n, 4, generator=g)
    tgt = torch.cat([torch.rand(n, 2, generator=g), torch.rand(n, 2, generator=g) * 4 + 0.3], 1)
    anc = torch.rand(n, 2, generator=g) * 4 + 0.3
    g_tgt = torch.randn(n, generator=g) * 0.01
    w_box = 0.37
    rr = r.clone().requires_grad_(True)
    pred = torch.cat([rr[:, :2].sigmoid(), torch.exp(rr[:, 2:4]) * anc], 1)
    ciou = oi.CIOU(pred, tgt, mode="xywh")
    iou = oi.cal_iou(pred, tgt, mode="xywh")
    (w_box * (1 - ciou).sum() + (iou.view(-1) * g_tgt).sum()).backward()
    grad, iou_o = np.empty((n, 4), np.float32), np.empty(n, np.float32)
    hc.hc_match_row_grad(fp(r.numpy().copy()), fp(tgt.numpy().copy()), fp(anc.numpy().copy()), n, C.c_float(w_box),
                         fp(g_tgt.numpy().copy()), fp(grad), fp(iou_o))
    np.testing.assert_allclose(iou_o, iou.detach().view(-1).numpy(), rtol=1e-5, atol=1e-6)
    scale = float(rr.grad.abs().max())
    np.testing.assert_allclose(grad, rr.grad.numpy(), rtol=1e-4, atol=2e-6 * scale)


def test_bce_derivatives(hc):
    n = 300
    g = torch.Generator().manual_seed(4)
    p = torch.rand(n, generator=g).clamp(1e-4, 1 - 1e-4)
    t = torch.rand(n, generator=g)
    t[:50] = 0.0
    t[50:100] = 1.0
    pp, tt = p.clone().requires_grad_(True), t.clone().requires_grad_(True)
    val = ol.bce_terms(pp, tt)
    val.sum().backward()
    value, dp, dt = (np.empty(n, np.float32) for _ in range(3))
    hc.hc_bce(fp(p.numpy().copy()), fp(t.numpy().copy()), n, fp(value), fp(dp), fp(dt))
    np.testing.assert_allclose(value, val.detach().numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(dp, pp.grad.numpy(), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(dt, tt.grad.numpy(), rtol=1e-5, atol=1e-6)
