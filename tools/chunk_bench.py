"""ValStep vs ChunkedValStep (graph replay and eager) at the headline shape."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastvision_b200 import synth  # noqa: E402
from fastvision_b200.pipeline import ValStep, ChunkedValStep  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="yolov3-416-coco")
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--steps", type=int, default=100)
args = ap.parse_args()
cfg = synth.CONFIGS[args.config]
g = synth.make_generator(2)
labels = synth.make_labels(cfg, args.batch, g)
dh = [h.cuda() for h in synth.make_heads(cfg, args.batch, labels, g)]
dl = labels.cuda()


def timeit(fn):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.steps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / args.steps


plain = ValStep(cfg.anchors_levels(), cfg.strides)
plain(dh, dl)
print("plain eager  %.4f ms" % timeit(lambda: plain(dh, dl)))
print("plain graph  %.4f ms" % timeit(plain.capture(dh, dl)))
for chunks in (2, 3, 4):
    st = ChunkedValStep(cfg.anchors_levels(), cfg.strides, chunks=chunks)
    st(dh, dl)
    print("chunks=%d eager %.4f ms" % (chunks, timeit(lambda: st(dh, dl))))
    print("chunks=%d graph %.4f ms" % (chunks, timeit(st.capture(dh, dl))))
