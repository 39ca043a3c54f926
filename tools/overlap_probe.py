"""Is the NMS kernel really running under the decode kernel?  Times the graph-replayed step with and without the
programmatic-dependent NMS, the decode alone with and without progress publishing, and prints when each image's NMS started
and ended relative to the step (globaltimer stamps, debug hook fvb_debug_set_nms_trace)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastvision_b200 import synth, _lib  # noqa: E402
from fastvision_b200.detection.models import yolov3_decode  # noqa: E402
from fastvision_b200.pipeline import ValStep  # noqa: E402

cfg = synth.CONFIGS[os.environ.get("CFG", "yolov3-416-coco")]
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
STEPS = 50
g = synth.make_generator(2)
labels = synth.make_labels(cfg, B, g)
dh = [h.cuda() for h in synth.make_heads(cfg, B, labels, g)]
dl = labels.cuda()
lib = _lib.load()


def graph_time(fn, steps=STEPS):
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


for overlap in (False, True):
    step = ValStep(cfg.anchors_levels(), cfg.strides, overlap_nms=overlap)
    step(dh, dl)
    torch.cuda.synchronize()

    def full():
        step._head(dh, dl)
        step._decode(dh)
        step._tail(dh, dl)

    def dec_only():
        step._decode(dh)
        if overlap:
            step.ctx.tile_sync().zero_()

    def dec_nms():
        step._decode(dh)
        step._nms()

    print("overlap=%s  step %.4f ms   decode+nms %.4f ms   decode only%s %.4f ms" % (
        overlap, graph_time(full), graph_time(dec_nms), " (+memset of the counters)" if overlap else "", graph_time(dec_only)), flush=True)
    # eager launches
    for _ in range(5):
        full()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(STEPS):
        full()
    b.record()
    torch.cuda.synchronize()
    print("             eager step %.4f ms" % (a.elapsed_time(b) / STEPS), flush=True)
    # when did every image's NMS run?  (one traced step, captured so that launch gaps do not blur it)
    trace = torch.zeros(9 * B + 2, dtype=torch.int64, device="cuda")
    lib.fvb_debug_set_nms_trace(_lib.dptr(trace))
    gph = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        full()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    with torch.cuda.graph(gph):
        full()
    lib.fvb_debug_set_nms_trace(_lib.dptr(None))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    gph.replay()
    torch.cuda.synchronize()
    trace[9 * B] = 2 ** 62
    trace[9 * B + 1] = 0
    torch.cuda.synchronize()
    e0.record()
    gph.replay()
    e1.record()
    torch.cuda.synchronize()
    full_t = trace.cpu()
    t = full_t[:8 * B].view(B, 8).double()
    entry = full_t[8 * B:9 * B].double()
    t0 = float(full_t[9 * B])
    print("             decode: first CTA start 0.0, last warp end %.1f us; NMS CTA entry: first %.1f, image 147 %.1f, image 148 %.1f, last %.1f us" % (
        (float(full_t[9 * B + 1]) - t0) / 1e3, (entry.min().item() - t0) / 1e3, (entry[min(147, B - 1)].item() - t0) / 1e3,
        (entry[min(148, B - 1)].item() - t0) / 1e3, (entry.max().item() - t0) / 1e3))
    st_, en_ = (t[:, 0] - t0) / 1e3, (t[:, 5] - t0) / 1e3
    print("             traced step %.1f us; NMS start offsets (every 16th image): %s" % (e0.elapsed_time(e1) * 1e3, [round(x, 1) for x in st_[::16].tolist()]))
    print("             NMS duration per image: mean %.1f max %.1f us; last NMS end %.1f us after the decode start" % (
        (en_ - st_).mean().item(), (en_ - st_).max().item(), en_.max().item()), flush=True)
