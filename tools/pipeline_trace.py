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
    tr = {"nms": (E(), E()), "loss": (E(), E())}
    pipe.submit(dh, dl, decode_events=d, trace=tr)
    recs.append((d, tr))
pipe.flush()
torch.cuda.synchronize()
for i, (d, tr) in enumerate(recs):
    f = lambda e: base.elapsed_time(e) * 1e3  # noqa: E731
    print("batch %d  decode %7.1f -> %7.1f (%5.1f us)   nms %7.1f -> %7.1f (%5.1f)   loss %7.1f -> %7.1f (%5.1f)" % (
        i, f(d[0]), f(d[1]), f(d[1]) - f(d[0]), f(tr["nms"][0]), f(tr["nms"][1]), f(tr["nms"][1]) - f(tr["nms"][0]),
        f(tr["loss"][0]), f(tr["loss"][1]), f(tr["loss"][1]) - f(tr["loss"][0])))
