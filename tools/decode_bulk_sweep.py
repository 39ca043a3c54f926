"""Decode kernel alone and the whole step, graph-replayed, on one GPU -- for every FVB_DECODE_BULK mode when run on branch
exp/decode-shifted-bulk (0: cp.async + 128-bit stores, 1: cp.async.bulk fetch, 2: cp.async.bulk copy-out, 3: both); on main the
knob does not exist and every mode times the same (cp.async) kernel.  Results of the experiment: profiles/r2_decode_data_path.md."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastvision_b200 import synth, _lib  # noqa: E402
from fastvision_b200.detection.models import yolov3_decode, DecodeContext  # noqa: E402
from fastvision_b200.pipeline import ValStep  # noqa: E402

cfg = synth.CONFIGS[os.environ.get("CFG", "yolov3-416-coco")]
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
g = synth.make_generator(2)
labels = synth.make_labels(cfg, B, g)
dh = [h.cuda() for h in synth.make_heads(cfg, B, labels, g)]
dl = labels.cuda()
anc, st = cfg.anchors_levels(), cfg.strides


def graph_time(fn, steps=50):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    gph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gph):
        for _ in range(steps):
            fn()
    gph.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    gph.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


for mode in os.environ.get("MODES", "0,1,2,3,0").split(","):
    os.environ["FVB_DECODE_BULK"] = mode
    _lib.load().fvb_debug_reload_knobs()
    ctx = DecodeContext(dh, anc, st)
    res = torch.empty(B, ctx.rows, ctx.k, device="cuda")
    nbytes = 2 * B * ctx.rows * ctx.k * 4
    plain = graph_time(lambda: yolov3_decode(dh, anc, st, ctx=ctx, out=res))
    fused = graph_time(lambda: (yolov3_decode(dh, anc, st, ctx=ctx, out=res, conf_thres=0.25, want_bce0=True), ctx.bitmap().zero_()))
    step = ValStep(anc, st)
    step(dh, dl)

    def full():
        step._head(dh, dl)
        step._decode(dh)
        step._tail(dh, dl)
    print("FVB_DECODE_BULK=%s  plain decode %.4f ms (%.0f GB/s)   fused decode (+bitmap memset) %.4f ms   step %.4f ms" %
          (mode, plain, nbytes / plain / 1e6, fused, graph_time(full)), flush=True)
