"""Launch the decode kernel a few times (plain, then with the fused side outputs) -- the command profiled with ncu."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastvision_b200 import synth  # noqa: E402
from fastvision_b200.detection.models import yolov3_decode, DecodeContext  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="yolov3-416-coco")
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--reps", type=int, default=2)
args = ap.parse_args()
cfg = synth.CONFIGS[args.config]
g = synth.make_generator(2)
labels = synth.make_labels(cfg, args.batch, g)
heads = [h.cuda() for h in synth.make_heads(cfg, args.batch, labels, g)]
anc, st = cfg.anchors_levels(), cfg.strides
ctx = DecodeContext(heads, anc, st)
res = torch.empty(args.batch, ctx.rows, ctx.k, device="cuda")
for _ in range(args.reps):
    yolov3_decode(heads, anc, st, ctx=ctx, out=res)
for _ in range(args.reps):
    ctx.bitmap().zero_()
    yolov3_decode(heads, anc, st, ctx=ctx, out=res, conf_thres=0.25, want_bce0=True)
torch.cuda.synchronize()
print("ok", float(res[0, 0, 4]))
