"""End-to-end parity of the caller glue (fastvision_b200.utils.Fit, the drop-in for utils/fit.py): a tiny YOLO-shaped model
trained for a few SGD steps through the CUDA loss forward + backward ends with the same losses and parameters as the same model
trained on the CPU through the oracle loss + torch autograd; validation (fused decode/NMS/loss step, batched matcher, device AP
integration) reproduces the oracle pipeline run on the model's own head outputs."""
import copy

import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

import oracle
from small_cfg import SMALL
from fastvision_b200 import synth
from fastvision_b200.loss import Yolov3Loss
from fastvision_b200.utils import Fit


class TinyYolo(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        self.anchors_per_level = cfg.anchors_levels()
        self.backbone_strides_per_level = cfg.strides
        self.heads = nn.ModuleList([nn.Conv2d(3, cfg.anchors_per_level * cfg.k, 1) for _ in cfg.strides])

    def forward(self, images, val=False):
        out = []
        for s, conv in zip(self.cfg.strides, self.heads):
            y = conv(F.avg_pool2d(images, int(s)))
            b, _, h, w = y.shape
            out.append(y.view(b, self.cfg.anchors_per_level, self.cfg.k, h, w).permute(0, 1, 3, 4, 2).contiguous())
        return out


def make_loader(cfg, batches, batch, seed):
    g = torch.Generator().manual_seed(seed)
    data = []
    for _ in range(batches):
        labels = synth.make_labels(cfg, batch, g)
        data.append((torch.rand(batch, 3, cfg.img, cfg.img, generator=g) * 4 - 2, labels))
    return data


def test_fit_train_matches_cpu_reference_training():
    cfg = SMALL
    torch.manual_seed(0)
    model = TinyYolo(cfg)
    ref = copy.deepcopy(model)
    loader = make_loader(cfg, 3, 4, 5)
    model = model.cuda()
    opt = torch.optim.SGD(model.parameters(), lr=0.05, momentum=0.9)
    fit = Fit(model, torch.device("cuda"), opt, torch.optim.lr_scheduler.StepLR(opt, 1), Yolov3Loss(model, 0.5, 0.05, 1.0, 0.5),
              end_epoch=2, train_loader=loader, verbose=False)
    fit.run_epoches()
    # the same training on the CPU: oracle loss + autograd
    ropt = torch.optim.SGD(ref.parameters(), lr=0.05, momentum=0.9)
    rsched = torch.optim.lr_scheduler.StepLR(ropt, 1)
    want = []
    for epoch in range(2):
        for images, labels in loader:
            ropt.zero_grad()
            loss = oracle.loss.yolov3_loss(ref(images), labels, cfg.anchors_levels(), cfg.strides)
            loss.sum().backward()
            ropt.step()
            want.append(float(loss.detach()))
        rsched.step()
    got = [h[2] for h in fit.history]
    np.testing.assert_allclose(got, want, rtol=2e-5)
    for p, q in zip(model.parameters(), ref.parameters()):
        np.testing.assert_allclose(p.detach().cpu().numpy(), q.detach().numpy(), rtol=1e-4, atol=1e-6)


def test_fit_val_matches_oracle_pipeline():
    cfg = SMALL
    torch.manual_seed(1)
    model = TinyYolo(cfg).cuda()
    with torch.no_grad():
        for conv in model.heads:                      # make some cells confident so that NMS and the matcher have work
            conv.bias.view(cfg.anchors_per_level, cfg.k)[:, 4] += 1.0
    loader = make_loader(cfg, 2, 5, 9)
    opt = torch.optim.SGD(model.parameters(), lr=0.0)
    fit = Fit(model, torch.device("cuda"), opt, torch.optim.lr_scheduler.StepLR(opt, 1), Yolov3Loss(model, 0.5, 0.05, 1.0, 0.5),
              end_epoch=1, train_loader=loader, val_loader=loader, verbose=False)
    loss, m_iou, m_cls, ids = fit._val()
    thr = np.linspace(0.5, 0.95, 10)
    est = oracle.map_.MapOracle(thr)
    last = None
    model.eval()
    with torch.no_grad():
        for images, labels in loader:
            heads = [h.cpu() for h in model(images.cuda())]
            last = oracle.loss.yolov3_loss(heads, labels, cfg.anchors_levels(), cfg.strides)
            res = oracle.decode.decode(heads, cfg.anchors_levels(), cfg.strides)
            for i in range(images.size(0)):                                                   # utils/fit.py:93-101
                s, c, b = oracle.nms.nms_lib(res[i], 0.25, 0.45, 300)
                pred = torch.cat([c.float(), s, b], 1) if s.numel() else torch.zeros(0, 6)
                est.process_one(pred, synth.labels_to_pixel_targets(labels, i, cfg.img, cfg.img))
    w_iou, w_cls, w_ids = est.fetch()
    np.testing.assert_allclose(loss, float(last), rtol=1e-5)
    assert ids == w_ids
    np.testing.assert_allclose(m_iou, w_iou, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(m_cls, w_cls, rtol=1e-9, atol=1e-12)
