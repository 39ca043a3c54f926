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


class _Backbone(nn.Module):
    """Stand-in with the constructor / accessor contract fastvision's darknet53 offers to Yolov3 (classfication/models/darknet53.py)."""

    def __init__(self, in_channels=3, including_top=False):
        super().__init__()
        self.convs = nn.ModuleList([nn.Conv2d(in_channels, 8, 1) for _ in SMALL.strides])

    def backbone_strides_per_level(self):
        return list(SMALL.strides)

    def backbone_channels_per_level(self):
        return [8 for _ in SMALL.strides]

    def forward(self, x):
        return [conv(F.avg_pool2d(x, int(s))) for s, conv in zip(SMALL.strides, self.convs)]


class _Neck(nn.Module):
    def __init__(self, feature_channels):
        super().__init__()

    def forward(self, feats):
        return feats


class _Head(nn.Module):
    """[B, A, H, W, 5 + C] per level, like detection/head/yolov3head.py:52-67."""

    def __init__(self, feature_channels, num_levels, num_anchors_per_level, num_classes):
        super().__init__()
        self.a, self.k = num_anchors_per_level, 5 + num_classes
        self.convs = nn.ModuleList([nn.Conv2d(c, a * self.k, 1) for c, a in zip(feature_channels, num_anchors_per_level)])

    def forward(self, feats):
        out = []
        for f, conv, a in zip(feats, self.convs, self.a):
            y = conv(f)
            b, _, h, w = y.shape
            out.append(y.view(b, a, self.k, h, w).permute(0, 1, 3, 4, 2).contiguous())
        return out


def test_fit_val_with_the_package_yolov3_in_eval_mode():
    """Fit._val with fastvision_b200.detection.models.Yolov3: after model.eval() its forward returns the TUPLE (head_out, results)
    like the reference's (detection/models/yolov3.py:33-54); _val must unpack it, must not decode twice, and must reproduce the
    oracle pipeline on the heads -- including with labels that do NOT arrive grouped by image."""
    from fastvision_b200.detection.models import Yolov3
    cfg = SMALL
    torch.manual_seed(3)
    anchors = torch.tensor(cfg.anchors_px, dtype=torch.float32)
    model = Yolov3(_Backbone, _Neck, _Head, anchors, [cfg.anchors_per_level] * cfg.levels, num_classes=cfg.num_classes).cuda()
    with torch.no_grad():
        for conv in model.head.convs:
            conv.bias.view(cfg.anchors_per_level, cfg.k)[:, 4] += 1.5
    loader = make_loader(cfg, 2, 5, 21)
    images, labels = loader[1]
    loader[1] = (images, labels[torch.randperm(labels.size(0), generator=torch.Generator().manual_seed(1))])   # ungrouped labels
    opt = torch.optim.SGD(model.parameters(), lr=0.0)
    fit = Fit(model, torch.device("cuda"), opt, torch.optim.lr_scheduler.StepLR(opt, 1), Yolov3Loss(model, 0.5, 0.05, 1.0, 0.5),
              end_epoch=1, train_loader=loader, val_loader=loader, verbose=False)
    calls = []
    orig = model.decode
    model.decode = lambda h: (calls.append(1), orig(h))[1]
    loss, m_iou, m_cls, ids = fit._val()
    assert not calls and model.decode_in_forward is True        # no redundant decode, flag restored
    model.eval()
    out = model(loader[0][0].cuda())
    assert isinstance(out, tuple) and out[1] is not None and out[1].size(1) == cfg.cells      # the reference's eval contract
    thr = np.linspace(0.5, 0.95, 10)
    est = oracle.map_.MapOracle(thr)
    last = None
    with torch.no_grad():
        for images, labels in loader:
            heads = [h.cpu() for h in model(images.cuda())[0]]
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


def test_map_matcher_many_targets_and_padded_batches():
    """More targets per image than the matcher stages in shared memory (1024) take the global-memory path; process_padded
    (capacity-shaped evidence, counts read once at fetch) accumulates the same rows as process_batch on exact tensors."""
    from fastvision_b200.metrics import CalculateMAP
    g = torch.Generator().manual_seed(5)
    thr = np.linspace(0.5, 0.95, 10)
    n_t, n_d = 1500, 260
    tb = torch.rand(n_t, 2, generator=g) * 380
    gts = torch.cat([torch.randint(0, 3, (n_t, 1), generator=g).float(), tb, tb + torch.rand(n_t, 2, generator=g) * 30 + 4], 1)
    src = torch.randint(0, n_t, (n_d,), generator=g)
    dets = torch.cat([gts[src, 0:1], torch.rand(n_d, 1, generator=g), gts[src, 1:] + torch.randn(n_d, 4, generator=g) * 1.5], 1)
    dets[5] = dets[4]
    eo, eg = oracle.map_.MapOracle(thr), CalculateMAP(thr)
    eo.process_one(dets, gts)
    eg.process_one(dets.cuda(), gts.cuda())
    assert np.array_equal(eo.correct_all_images[-1], eg.correct_all_images[-1])
    # padded path == exact path on a ValStep output
    cfg, batch = SMALL, 6
    gg = synth.make_generator(1, rank=5)
    labels = synth.make_labels(cfg, batch, gg)
    heads = [h.cuda() for h in synth.make_heads(cfg, batch, labels, gg)]
    from fastvision_b200.pipeline import ValStep
    step = ValStep(cfg.anchors_levels(), cfg.strides)
    o = step(heads, labels.cuda())
    a, b = CalculateMAP(thr), CalculateMAP(thr)
    a.process_padded(o["boxes"], o["scores"], o["cls"], o["cnt"], labels.cuda(), cfg.img, cfg.img)
    for i, d in enumerate(step.detections()):
        b.process_one(d, synth.labels_to_pixel_targets(labels, i, cfg.img, cfg.img).cuda())
    ra, rb = a.fetch(), b.fetch()
    assert ra[2] == rb[2] and np.array_equal(ra[0], rb[0]) and np.array_equal(ra[1], rb[1])
    assert np.array_equal(np.concatenate(a.correct_all_images), np.concatenate(b.correct_all_images))


def _evidence_reference(boxes, scores, cls, cnt, labels, w, h):
    """utils/fit.py:96-99 per image, concatenated: detections [cls, conf, xyxy] and targets [cls, xyxy px] grouped by image."""
    b = boxes.size(0)
    dets, doff, gts, goff = [], [0], [], [0]
    for i in range(b):
        k = max(int(cnt[i]), 0)
        dets.append(torch.cat([cls[i, :k].float()[:, None], scores[i, :k, None], boxes[i, :k]], 1))
        doff.append(doff[-1] + k)
        t = labels[labels[:, 0] == i][:, 1:]
        half = t[:, 3:5] / 2
        xyxy = torch.cat([t[:, 1:3] - half, t[:, 1:3] + half], 1) * torch.tensor([w, h, w, h], dtype=t.dtype)
        gts.append(torch.cat([t[:, 0:1], xyxy], 1))
        goff.append(goff[-1] + t.size(0))
    return torch.cat(dets), doff, torch.cat(gts), goff


@pytest.mark.parametrize("seed", range(6))
def test_val_evidence_kernel_random_cases(seed):
    """fvb_val_evidence_f32 against the per-image Python of utils/fit.py:94-99: empty images, images without labels, labels in
    arbitrary order, labels of images outside the batch (dropped), a failed image (cnt = -1), no labels at all."""
    from fastvision_b200 import _lib
    g = torch.Generator().manual_seed(300 + seed)
    b, md = int(torch.randint(1, 40, (1,), generator=g)), int(torch.randint(1, 50, (1,), generator=g))
    t = 0 if seed == 5 else int(torch.randint(0, 200, (1,), generator=g))
    boxes = torch.rand(b, md, 4, generator=g) * 400
    scores = torch.rand(b, md, generator=g)
    cls = torch.randint(0, 80, (b, md), generator=g)
    cnt = torch.randint(0, md + 1, (b,), generator=g).int()
    cnt[0] = -1 if seed % 2 else 0
    img = torch.randint(0, b + (2 if seed % 3 == 0 else 0), (t,), generator=g).float()     # sometimes ids beyond the batch
    if seed % 2 == 0:
        img = torch.sort(img)[0]                                                              # collate order (fast path)
    labels = torch.cat([img[:, None], torch.randint(0, 80, (t, 1), generator=g).float(), torch.rand(t, 4, generator=g)], 1)
    w, h = 416.0, 320.0
    want_d, want_doff, want_g, want_goff = _evidence_reference(boxes, scores, cls, cnt, labels, w, h)
    db, ds, dc, dn, dl = boxes.cuda(), scores.cuda(), cls.cuda(), cnt.cuda(), labels.cuda()
    dets = torch.full((b * md, 6), -7.0, device="cuda")
    gts = torch.full((max(t, 1), 5), -7.0, device="cuda")
    doff = torch.empty(b + 1, dtype=torch.int32, device="cuda")
    goff = torch.empty(b + 1, dtype=torch.int32, device="cuda")
    _lib.check(_lib.load().fvb_val_evidence_f32(_lib.dptr(db), _lib.dptr(ds), _lib.dptr(dc), _lib.dptr(dn), b, md, _lib.dptr(dl), t,
                                                w, h, _lib.dptr(dets), _lib.dptr(doff), _lib.dptr(gts), _lib.dptr(goff),
                                                _lib.stream()), "val_evidence")
    torch.cuda.synchronize()
    assert doff.cpu().tolist() == want_doff and goff.cpu().tolist() == want_goff
    assert torch.equal(dets[:want_doff[-1]].cpu(), want_d)
    np.testing.assert_allclose(gts[:want_goff[-1]].cpu().numpy(), want_g.numpy(), rtol=1e-6, atol=1e-4)
