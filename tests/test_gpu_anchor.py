"""GPU parity of the anchor k-means drop-in (fvb_kmeans_step_f32) against golden vectors recorded from the reference's
KMeans (detection/tools/ANCHOR.py:11-46) and the numpy oracle.  Assignments: exact; centres: rtol 1e-5 (fp64 sums on the
device vs numpy's float32 pairwise mean)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import oracle
from fastvision_b200.detection.tools import KMeans, AnchorGenerator


@pytest.mark.parametrize("tag", ["a", "b"])
def test_kmeans_golden(golden_anchor, tag):
    g = golden_anchor
    k, iters, seed = (int(v) for v in g[tag + "_cfg"])
    xs = g[tag + "_samples"].copy()
    np.random.seed(seed)
    centers, cats = KMeans(xs=xs, k=k).fit(iters=iters)
    # the drop-in shuffles the caller's array in place exactly like the reference
    ref = g[tag + "_samples"].copy()
    np.random.seed(seed)
    np.random.shuffle(ref)
    assert np.array_equal(xs, ref)
    assert centers.shape == g[tag + "_centers"].shape and centers.dtype == np.float32
    np.testing.assert_allclose(centers, g[tag + "_centers"], rtol=1e-5, atol=1e-7)
    assert np.array_equal(cats, g[tag + "_categories"])


def test_kmeans_vs_oracle_large_and_generator(tmp_path):
    gen = torch.Generator().manual_seed(8)
    base = torch.rand(9, 2, generator=gen) * 0.5 + 0.02
    wh = (base[torch.randint(0, 9, (20000,), generator=gen)] * torch.exp(torch.randn(20000, 2, generator=gen) * 0.2)).clamp(0.003, 1.0)
    xs = wh.numpy().astype(np.float32)
    np.random.seed(1)
    want_c, want_cat = oracle.anchor.KMeans(xs.copy(), 9).fit(15)
    np.random.seed(1)
    got_c, got_cat = KMeans(xs.copy(), 9).fit(15)
    np.testing.assert_allclose(got_c, want_c, rtol=2e-5, atol=1e-7)
    assert (got_cat != want_cat).mean() < 1e-3          # a sample sitting on a cluster boundary may flip with the mean's last bit

    class Loader(list):
        pass
    loader = Loader([(torch.zeros(2, 3, 416, 416), torch.cat([torch.zeros(500, 4), wh[i * 500:(i + 1) * 500]], 1)) for i in range(4)])
    np.random.seed(2)
    anchors = AnchorGenerator([loader], k=9, iters=10, plot=False, cache=str(tmp_path)).get_anchors()
    assert anchors.shape == (9, 2) and np.all(np.diff(anchors[:, 0] * anchors[:, 1]) <= 0)   # sorted by area, scaled to pixels
    cached = AnchorGenerator([loader], k=9, iters=10, plot=False, cache=str(tmp_path), use_cache=True).get_anchors()
    np.testing.assert_allclose(cached, anchors)
