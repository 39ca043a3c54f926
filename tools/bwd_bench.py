"""Time Yolov3Loss forward + backward (B=256 YOLOv3-416) -- the training-side call of the path (utils/fit.py:57-63).
Algorithmic bytes of the backward: one write of the head gradients (N*K*4 per image) + one read of the objectness channel."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastvision_b200 import synth  # noqa: E402
from fastvision_b200 import loss as fl  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="yolov3-416-coco")
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--steps", type=int, default=50)
args = ap.parse_args()
cfg = synth.CONFIGS[args.config]
g = synth.make_generator(2)
labels = synth.make_labels(cfg, args.batch, g)
heads = [h.cuda() for h in synth.make_heads(cfg, args.batch, labels, g)]
dl = labels.cuda()


class M:
    anchors_per_level = cfg.anchors_levels()
    backbone_strides_per_level = cfg.strides


lossf = fl.Yolov3Loss(M(), 0.5, 0.05, 1.0, 0.5)
from fastvision_b200 import _lib  # noqa: E402
ctx = lossf._context(heads)
saved = torch.empty(int(_lib.load().fvb_yolov3_saved_conf_floats(ctx.geom)), device="cuda")
lossf._forward_impl(heads, dl, ctx, None, None, None, saved_conf=saved)
parts = lossf.partials
grads = [torch.empty_like(h) for h in heads]
one = torch.ones(1, device="cuda")


def fwd():
    lossf._forward_impl(heads, dl, ctx, None, None, None, saved_conf=saved)


def bwd():
    lossf.backward_heads(heads, dl, one, parts, args.batch, grads=grads, saved_conf=saved)


def bwd_strided():
    lossf.backward_heads(heads, dl, one, parts, args.batch, grads=grads)


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


def fill():
    for gr in grads:
        gr.zero_()


t_f, t_b, t_z, t_bs = timeit(fwd), timeit(bwd), timeit(fill), timeit(bwd_strided)
n_floats = sum(h.numel() for h in heads)
rows = n_floats // heads[0].size(-1)
alg = n_floats * 4 + rows * 4   # gradient write + compact objectness read
peak = 6552.6
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
print(json.dumps({"config": cfg.name, "batch": args.batch, "loss_forward_ms": t_f, "loss_backward_ms": t_b,
                  "backward_algorithmic_bytes": alg, "backward_GBps": alg / (t_b * 1e-3) / 1e9,
                  "backward_frac_of_hbm_peak": alg / (t_b * 1e-3) / 1e9 / peak,
                  "loss_backward_strided_reread_ms": t_bs, "torch_zero_fill_same_buffers_ms": t_z, "zero_fill_GBps": n_floats * 4 / (t_z * 1e-3) / 1e9}))
