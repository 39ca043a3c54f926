"""Sweep the decode kernel's tuning knobs (env: FVB_DECODE_WARPS / _STAGES / _TILE_ROWS) on one GPU."""
import itertools
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastvision_b200 import synth, _lib  # noqa: E402
from fastvision_b200.detection.models import yolov3_decode, DecodeContext  # noqa: E402
from microbench import timeit  # noqa: E402

cfg = synth.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "yolov3-416-coco"]
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 256
g = synth.make_generator(2)
labels = synth.make_labels(cfg, batch, g)
heads = [h.cuda() for h in synth.make_heads(cfg, batch, labels, g)]
anc, st = cfg.anchors_levels(), cfg.strides
rows = None
for warps, gb in [(16, 8), (16, 1), (16, 4), (16, 16), (20, 8), (16, 8), (12, 8), (16, 2)]:
    stages, tr = 1, gb
    os.environ["FVB_DECODE_BATCH"] = str(gb)
    os.environ["FVB_DECODE_WARPS"] = str(warps)
    _lib.load().fvb_debug_reload_knobs()          # the library caches the knobs after its first launch
    ctx = DecodeContext(heads, anc, st)
    res = torch.empty(batch, ctx.rows, ctx.k, device="cuda")
    nbytes = 2 * batch * ctx.rows * ctx.k * 4
    try:
        plain, _ = timeit(lambda: yolov3_decode(heads, anc, st, ctx=ctx, out=res), iters=40, warm=5)
        fused, _ = timeit(lambda: yolov3_decode(heads, anc, st, ctx=ctx, out=res, conf_thres=0.25, want_bce0=True), iters=40, warm=5)
        ctx.bitmap().zero_()
    except Exception as e:  # noqa: BLE001
        print(stages, warps, tr, "failed:", str(e)[:100], flush=True)
        continue
    print("stages %d warps %2d tile_rows %3d  plain %.4f ms %.0f GB/s   fused %.4f ms %.0f GB/s" %
          (stages, warps, tr, plain, nbytes / plain / 1e6, fused, nbytes / fused / 1e6), flush=True)
