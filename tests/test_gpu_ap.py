"""GPU parity of the device AP integration (fvb_map_ap_f64) against the oracle's numpy restatement of
CalculateMAP.fetch (metrics/map.py:85-141).  float64 results: rtol 1e-12; class lists and positive counts: exact."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import oracle
from fastvision_b200.metrics import CalculateMAP

THR = np.linspace(0.5, 0.95, 10)


def synth_eval(images, classes, max_det, seed, tie_conf=False):
    """Per image: a few targets and detections (jittered copies of targets + false positives), classes as floats."""
    g = torch.Generator().manual_seed(seed)
    dets, gts, doff, goff = [], [], [0], [0]
    for _ in range(images):
        nt = int(torch.randint(0, 9, (1,), generator=g))
        xy = torch.rand(nt, 2, generator=g) * 300
        wh = torch.rand(nt, 2, generator=g) * 80 + 10
        tb = torch.cat([xy, xy + wh], 1)
        tc = torch.randint(0, classes, (nt, 1), generator=g).float()
        gts.append(torch.cat([tc, tb], 1))
        rows = []
        for j in range(nt):
            for _ in range(int(torch.randint(0, 4, (1,), generator=g))):
                jit = torch.randn(4, generator=g) * float(torch.rand(1, generator=g)) * 8
                c = tc[j] if float(torch.rand(1, generator=g)) < 0.85 else torch.randint(0, classes + 2, (1,), generator=g).float()
                rows.append(torch.cat([c, torch.rand(1, generator=g), tb[j] + jit]))
        for _ in range(int(torch.randint(0, max_det, (1,), generator=g))):
            p = torch.rand(2, generator=g) * 300
            rows.append(torch.cat([torch.randint(0, classes + 2, (1,), generator=g).float(), torch.rand(1, generator=g), p, p + 30]))
        d = torch.stack(rows) if rows else torch.zeros(0, 6)
        if tie_conf and d.size(0):
            d[:, 1] = torch.round(d[:, 1] * 8) / 8          # heavy confidence ties
        dets.append(d)
        doff.append(doff[-1] + d.size(0))
        goff.append(goff[-1] + nt)
    return dets, gts, torch.tensor(doff, dtype=torch.int32), torch.tensor(goff, dtype=torch.int32)


def run_both(dets, gts, doff, goff, per_image=False):
    est = CalculateMAP(THR)
    if per_image:
        for d, t in zip(dets, gts):
            est.process_one(d.cuda(), t.cuda())
    else:
        est.process_batch(torch.cat(dets).cuda(), doff.cuda(), torch.cat(gts).cuda(), goff.cuda())
    ora = oracle.map_.MapOracle(THR)
    # the oracle integrates the SAME correct bits (the matcher has its own parity test): isolate the AP integration
    ora.correct_all_images = est.correct_all_images
    ora.seen_all_targets_cls = est.seen_all_targets_cls
    return est, ora


@pytest.mark.parametrize("images,classes,max_det,tie", [(40, 5, 6, False), (300, 20, 40, False), (120, 3, 30, True), (60, 300, 10, False)])
def test_ap_fetch_vs_oracle(images, classes, max_det, tie):
    dets, gts, doff, goff = synth_eval(images, classes, max_det, seed=images + classes, tie_conf=tie)
    est, ora = run_both(dets, gts, doff, goff)
    if tie:
        # the reference's argsort(-conf) leaves ties unspecified; the device order is "input order", i.e. a stable sort
        orig = np.argsort

        def stable(a, *args, **kw):
            kw["kind"] = "stable"
            return orig(a, *args, **kw)
        np.argsort = stable
        try:
            want = ora.fetch()
        finally:
            np.argsort = orig
    else:
        want = ora.fetch()
    got = est.fetch()
    np.testing.assert_allclose(got[0], want[0], rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(got[1], want[1], rtol=1e-12, atol=1e-15)
    assert got[2] == want[2]
    ap, pos = est.ap_table()
    seen = np.concatenate(est.seen_all_targets_cls)
    for c in got[2]:
        assert int(pos[c]) == int((seen == c).sum())


def test_ap_edge_cases():
    # a class with targets but no detections scores 0.5 at every threshold (interp of [0,1] -> [1,0])
    est = CalculateMAP(THR)
    est.process_one(torch.tensor([[1.0, 0.9, 0, 0, 10, 10]]).cuda(), torch.tensor([[1.0, 0, 0, 10, 10], [2.0, 50, 50, 60, 60]]).cuda())
    m_iou, m_cls, ids = est.fetch()
    assert ids == [1, 2]
    np.testing.assert_allclose(m_cls, [0.995, 0.5], rtol=1e-12)   # a perfect class integrates to 0.995: np.interp returns the trailing sentinel 0 at recall 1.0
    # per-image calls and one batched call accumulate the same evidence
    dets, gts, doff, goff = synth_eval(25, 4, 8, seed=5)
    a, _ = run_both(dets, gts, doff, goff, per_image=True)
    b, _ = run_both(dets, gts, doff, goff)
    fa, fb = a.fetch(), b.fetch()
    assert np.array_equal(fa[0], fb[0]) and np.array_equal(fa[1], fb[1]) and fa[2] == fb[2]
    # state()/load_state() round trip (what data-parallel ranks all-gather)
    rows, tg = a.state()
    c = CalculateMAP(THR)
    c.load_state(rows, tg)
    fc = c.fetch()
    assert np.array_equal(fa[0], fc[0]) and fa[2] == fc[2]
    with pytest.raises(ValueError):
        CalculateMAP(THR).fetch()


def test_ap_config5_scale():
    """BASELINE configs[4] scale: 5000 images, <= 300 detections each (~0.75 M rows), 80 classes."""
    g = torch.Generator().manual_seed(7)
    n, classes = 750_000, 80
    dets = torch.zeros(n, 6)
    dets[:, 0] = torch.randint(0, classes, (n,), generator=g).float()
    dets[:, 1] = torch.rand(n, generator=g)
    correct = (torch.rand(n, 10, generator=g) < torch.linspace(0.6, 0.1, 10)).to(torch.uint8)
    correct = torch.cummin(correct, dim=1)[0]            # a match at a higher IoU threshold implies the lower ones
    targets = torch.randint(0, classes, (37_000,), generator=g).float()
    est = CalculateMAP(THR)
    est._dets, est._correct, est._targets = [dets.cuda()], [correct.cuda()], [targets.cuda()]
    got = est.fetch()
    ora = oracle.map_.MapOracle(THR)
    block = np.zeros((n, 12))
    block[:, 0], block[:, 1], block[:, 2:] = dets[:, 1].numpy(), dets[:, 0].numpy(), correct.numpy()
    ora.correct_all_images, ora.seen_all_targets_cls = [block], [targets.numpy()]
    want = ora.fetch()
    np.testing.assert_allclose(got[0], want[0], rtol=1e-12)
    np.testing.assert_allclose(got[1], want[1], rtol=1e-12)
    assert got[2] == want[2]
