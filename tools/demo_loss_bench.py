"""Time the demos' ComputeLoss forward + backward on the huaweiShip shape (BASELINE configs[2]: YOLOv3-608, 10 classes).
Algorithmic bytes: forward = 5 of K head planes read once (20 B per predicted box) + 1 B mask; backward = the whole
gradient written once (K*4 B per box) + the objectness plane and mask read (5 B per box)."""
import argparse
import json
import os
import sys
import types

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastvision_b200 import synth  # noqa: E402
from fastvision_b200.loss import ComputeLoss, ComputeLossU  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="yolov3-608-ship")
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--steps", type=int, default=30)
ap.add_argument("--flavour", default="ship")
args = ap.parse_args()
cfg = synth.CONFIGS[args.config]
g = synth.make_generator(3)
labels = synth.make_labels(cfg, args.batch, g)
heads5 = synth.make_heads(cfg, args.batch, labels, g)
heads = [h.permute(0, 1, 4, 2, 3).reshape(h.size(0), -1, h.size(2), h.size(3)).contiguous().cuda() for h in heads5]
del heads5
dl = labels.cuda()
anchors = [a.reshape(-1, 2) / s for a, s in zip(cfg.anchors_levels(), cfg.strides)]
lossf = (ComputeLoss if args.flavour == "ship" else ComputeLossU)(strict=False)
ctx = lossf._context(heads, anchors)
_, parts, mask = lossf._run(heads, dl, ctx, want_mask=True)
grads = [torch.empty_like(h) for h in heads]
gout = torch.ones(3, device="cuda")


def timeit(fn):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.steps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / args.steps


t_f = timeit(lambda: lossf._run(heads, dl, ctx, want_mask=True))
t_b = timeit(lambda: lossf.backward_heads(heads, dl, gout, parts, mask, ctx=ctx, grads=grads))
boxes = sum(h.numel() for h in heads) // ctx.geom.channels
k = ctx.geom.channels
fwd_bytes, bwd_bytes = boxes * 21, boxes * (k * 4 + 5)
peak = 6552.6
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
print(json.dumps({"config": cfg.name, "flavour": args.flavour, "batch": args.batch, "labels": int(labels.size(0)), "boxes": boxes,
                  "forward_ms": t_f, "forward_GBps": fwd_bytes / t_f / 1e6, "forward_frac": fwd_bytes / t_f / 1e6 / peak,
                  "backward_ms": t_b, "backward_GBps": bwd_bytes / t_b / 1e6, "backward_frac": bwd_bytes / t_b / 1e6 / peak,
                  "images_per_s_fwd_bwd": args.batch / ((t_f + t_b) * 1e-3)}))
