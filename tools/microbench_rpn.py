"""RPN proposal filter timing (BASELINE config 4: B=64, 50x50x9 anchors) on one GPU, CUDA events."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastvision_b200 import synth  # noqa: E402
from fastvision_b200.detection import tools as ft  # noqa: E402
from microbench import timeit  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
args = ap.parse_args()
gen = synth.make_generator(4)
cls, reg = synth.make_rpn_inputs(args.batch, 50, 50, 9, gen)
dc, dr = cls.cuda(), reg.cuda()
base = ft.get_base_anchor([128, 256, 512], [1, 0.5, 2]) / 16
out = {}
for pre, post in [(12000, 2000), (6000, 300), (2000, 2000)]:
    med, mn = timeit(lambda: ft.filter_proposals_batched(dc, dr, base, pre, post, 0.7), iters=10, warm=2)
    _, cnt = ft.filter_proposals_batched(dc, dr, base, pre, post, 0.7)
    out["rpn_pre%d_post%d" % (pre, post)] = {"ms_median": med, "ms_min": mn, "images_per_s": args.batch / med * 1e3,
                                             "kept_mean": cnt.float().mean().item()}
    print(pre, post, out["rpn_pre%d_post%d" % (pre, post)], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/microbench_rpn_b%d.json" % args.batch, "w") as f:
    json.dump(out, f, indent=1)
