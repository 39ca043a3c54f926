"""Cost of the decode kernel's fused side outputs (candidate bitmap + records, objectness-BCE partials), per configuration."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastvision_b200 import synth  # noqa: E402
from fastvision_b200.detection.models import yolov3_decode, DecodeContext  # noqa: E402
from microbench import timeit  # noqa: E402

cfg = synth.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "yolov3-608-ship"]
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
g = synth.make_generator(2)
labels = synth.make_labels(cfg, batch, g)
heads = [h.cuda() for h in synth.make_heads(cfg, batch, labels, g)]
anc, st = cfg.anchors_levels(), cfg.strides
ctx = DecodeContext(heads, anc, st)
res = torch.empty(batch, ctx.rows, ctx.k, device="cuda")
nbytes = 2 * batch * ctx.rows * ctx.k * 4
for name, kw in [("plain", {}), ("bce0", {"want_bce0": True}), ("bitmap+records", {"conf_thres": 0.25}),
                 ("both", {"conf_thres": 0.25, "want_bce0": True}), ("bitmap+records thr 0.999", {"conf_thres": 0.999})]:
    med, mn = timeit(lambda: yolov3_decode(heads, anc, st, ctx=ctx, out=res, **kw), iters=30, warm=5)
    if "conf_thres" in kw:
        ctx.bitmap().zero_()
    print("%-26s %.4f ms (min %.4f)  %.0f GB/s" % (name, med, mn, nbytes / med / 1e6), flush=True)
