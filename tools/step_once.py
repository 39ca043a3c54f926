"""Run the fused val step a few times (eager, no graph) -- the command profiled with ncu."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastvision_b200 import synth  # noqa: E402
from fastvision_b200.pipeline import ValStep  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="yolov3-416-coco")
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--steps", type=int, default=4)
args = ap.parse_args()
cfg = synth.CONFIGS[args.config]
g = synth.make_generator(2)
labels = synth.make_labels(cfg, args.batch, g)
heads = synth.make_heads(cfg, args.batch, labels, g)
dh, dl = [h.cuda() for h in heads], labels.cuda()
step = ValStep(cfg.anchors_levels(), cfg.strides)
for _ in range(args.steps):
    out = step(dh, dl)
torch.cuda.synchronize()
print("loss", out["loss"].item(), "kept", out["cnt"].float().mean().item())
