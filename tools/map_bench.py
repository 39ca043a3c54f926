"""Time the batched mAP matcher (fvb_map_match_f32) and the device AP integration on synthetic evidence."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastvision_b200.metrics import CalculateMAP  # noqa: E402

images, per, ngt = 250, 300, 7
g = torch.Generator().manual_seed(1)
xy = torch.rand(images * ngt, 2, generator=g) * 300
wh = torch.rand(images * ngt, 2, generator=g) * 80 + 10
gts = torch.cat([torch.randint(0, 80, (images * ngt, 1), generator=g).float(), xy, xy + wh], 1).cuda()
goff = (torch.arange(images + 1) * ngt).int().cuda()
src = torch.randint(0, ngt, (images, per), generator=g) + (torch.arange(images) * ngt)[:, None]
base = gts.cpu()[src.view(-1)]
dets = torch.cat([base[:, :1], torch.rand(images * per, 1, generator=g), base[:, 1:] + torch.randn(images * per, 4, generator=g) * 6], 1).cuda()
doff = (torch.arange(images + 1) * per).int().cuda()
est = CalculateMAP(np.linspace(0.5, 0.95, 10))
for _ in range(3):
    est.match(dets, doff, gts, goff)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20):
    c = est.match(dets, doff, gts, goff)
b.record()
torch.cuda.synchronize()
print("match ms per %d images x %d dets: %.4f  (correct rate %.3f)" % (images, per, a.elapsed_time(b) / 20, c.float().mean().item()))
for _ in range(20):
    est.process_batch(dets, doff, gts, goff)
torch.cuda.synchronize()
t0 = time.perf_counter()
r = est.fetch()
print("fetch (1.5M rows) ms: %.3f  mAP50 %.4f" % ((time.perf_counter() - t0) * 1e3, r[0][0]))
t0 = time.perf_counter()
r = est.fetch()
print("fetch again ms: %.3f" % ((time.perf_counter() - t0) * 1e3))
