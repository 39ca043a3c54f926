import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from fastvision_b200 import synth
from fastvision_b200.detection.models import yolov3_decode, DecodeContext
from fastvision_b200.pipeline import ValStep
cfg = synth.SHIP608
B = 128
g = synth.make_generator(3)
labels = synth.make_labels(cfg, B, g)
dh = [h.cuda() for h in synth.make_heads(cfg, B, labels, g)]
dl = labels.cuda()
anc, st = cfg.anchors_levels(), cfg.strides
def graph_time(fn, steps=50):
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3): fn()
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    gph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gph):
        for _ in range(steps): fn()
    gph.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); gph.replay(); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / steps
ctx = DecodeContext(dh, anc, st)
res = torch.empty(B, ctx.rows, ctx.k, device="cuda")
plain = graph_time(lambda: yolov3_decode(dh, anc, st, ctx=ctx, out=res))
step = ValStep(anc, st); step(dh, dl)
def full():
    step._head(dh, dl); step._decode(dh); step._tail(dh, dl)
print("batch knob %s warps %s: B=128 plain decode %.4f ms (%.0f GB/s)  step %.4f ms" % (os.environ.get("FVB_DECODE_BATCH"), os.environ.get("FVB_DECODE_WARPS"), plain, 2*B*ctx.rows*ctx.k*4/plain/1e6, graph_time(full)), flush=True)
