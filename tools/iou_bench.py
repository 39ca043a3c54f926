"""Pairwise / element-wise IoU-family kernel throughput (K2): the output write is the algorithmic traffic of the pairwise form."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastvision_b200.detection import tools as ft  # noqa: E402
from fastvision_b200 import loss as fl  # noqa: E402

g = torch.Generator().manual_seed(0)


def boxes(n):
    xy = torch.rand(n, 2, generator=g) * 500
    wh = torch.rand(n, 2, generator=g) * 100 + 1
    return torch.cat([xy, xy + wh], 1).cuda()


def timeit(fn, n=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


out = {}
a, b = boxes(8192), boxes(8192)
for name, fn in [("iou", ft.cal_iou_batch), ("giou", ft.GIOU_batch), ("diou", ft.DIOU_batch), ("ciou", ft.CIOU_batch)]:
    ms = timeit(lambda: fn(a, b))
    out["pairwise_%s_8192x8192" % name] = {"ms": ms, "Gpairs_per_s": 8192 * 8192 / ms / 1e6, "write_GBps": 8192 * 8192 * 4 / ms / 1e6}
big_a, big_b = boxes(4_000_000), boxes(4_000_000)
for name, fn in [("iou", ft.cal_iou), ("ciou", ft.CIOU)]:
    ms = timeit(lambda: fn(big_a, big_b))
    out["elementwise_%s_4M" % name] = {"ms": ms, "GBps": 4e6 * 36 / ms / 1e6}
ms = timeit(lambda: fl.CIOULoss("mean")(big_a, big_b))
out["ciou_loss_mean_4M"] = {"ms": ms, "GBps": 4e6 * 32 / ms / 1e6}
x = big_a.clone().requires_grad_(True)


def fb():
    x.grad = None
    fl.CIOULoss("mean")(x, big_b).backward()


out["ciou_loss_fwd_bwd_4M"] = {"ms": timeit(fb)}
print(json.dumps(out, indent=1))
