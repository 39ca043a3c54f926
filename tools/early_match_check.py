"""Two-part loss (match beside the decode + finish after it) vs the one-call form on the same batch, and the graph-replayed step time."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastvision_b200 import synth  # noqa: E402
from fastvision_b200.pipeline import ValStep  # noqa: E402

cfg, B = synth.COCO416, int(sys.argv[1]) if len(sys.argv) > 1 else 256
g = synth.make_generator(2)
labels = synth.make_labels(cfg, B, g)
dh = [h.cuda() for h in synth.make_heads(cfg, B, labels, g)]
dl = labels.cuda()
step = ValStep(cfg.anchors_levels(), cfg.strides)
o = step(dh, dl)                                   # two-part form
torch.cuda.synchronize()
l2, p2, c2 = float(o["loss"]), o["partials"].clone(), o["cnt"].clone()
step._decode(dh)
step._tail(dh, dl, reduce_inside=False, early_match=False)   # one-call form
torch.cuda.synchronize()
l1, p1 = float(o["loss"]), o["partials"].clone()
print("loss two-part %.9g one-call %.9g rel diff %.3g; partials max rel diff %.3g; detections equal %s" % (
    l2, l1, abs(l2 - l1) / abs(l1), float(((p2 - p1).abs() / p1.abs().clamp_min(1e-300)).max()), bool(torch.equal(c2, o["cnt"]))))
o0 = step(dh, dl[:0])                              # no labels: only the objectness term
torch.cuda.synchronize()
print("no-label loss %.9g" % float(o0["loss"]))
for early in (True, False):
    gph = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.graph(gph, stream=s):
        for _ in range(20):
            step._decode(dh)
            step._tail(dh, dl, early_match=early)
    torch.cuda.current_stream().wait_stream(s)
    gph.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        gph.replay()
    b.record()
    torch.cuda.synchronize()
    print("early_match=%s: %.4f ms/step" % (early, a.elapsed_time(b) / 100))
