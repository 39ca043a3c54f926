"""Timeline of the pipelined step (decode / NMS / loss start+stop per batch) from CUDA events on each stream."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastvision_b200 import synth  # noqa: E402
from fastvision_b200.pipeline import ValPipeline  # noqa: E402

cfg = synth.COCO416
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
g = synth.make_generator(2)
labels = synth.make_labels(cfg, B, g)
heads = synth.make_heads(cfg, B, labels, g)
dh, dl = [h.cuda() for h in heads], labels.cuda()
pipe = ValPipeline(cfg.anchors_levels(), cfg.strides)
for _ in range(6):
    pipe.submit(dh, dl)
pipe.flush()
torch.cuda.synchronize()
E = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
base = E()
base.record()
recs = []
for i in range(8):
    d = (E(), E())
    done = E()
    pipe.submit(dh, dl, decode_events=d)
    pipe.wait()
    done.record()
    recs.append((d, done))
pipe.flush()
torch.cuda.synchronize()
for i, (d, done) in enumerate(recs):
    f = lambda e: base.elapsed_time(e) * 1e3  # noqa: E731
    print("batch %d  decode %7.1f -> %7.1f (%5.1f us)   batch complete at %7.1f (tail %5.1f us after its decode)" % (
        i, f(d[0]), f(d[1]), f(d[1]) - f(d[0]), f(done), f(done) - f(d[1])))
