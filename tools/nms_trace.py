"""Per-phase clock64 breakdown of yolo_nms_kernel (debug hook fvb_debug_set_nms_trace)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastvision_b200 import synth, _lib  # noqa: E402
from fastvision_b200.detection.models import yolov3_decode, DecodeContext  # noqa: E402
from fastvision_b200.detection.tools import non_max_suppression_batched  # noqa: E402

cfg = synth.CONFIGS[os.environ.get("CFG", "yolov3-416-coco")]
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
g = synth.make_generator(2)
labels = synth.make_labels(cfg, B, g)
heads = [h.cuda() for h in synth.make_heads(cfg, B, labels, g)]
anc, st = cfg.anchors_levels(), cfg.strides
ctx = DecodeContext(heads, anc, st)
res = yolov3_decode(heads, anc, st, ctx=ctx, conf_thres=0.25)
lib = C.CDLL(_lib.LIB_PATH)
trace_all = torch.zeros(9 * B + 2, dtype=torch.int64, device="cuda")
trace = trace_all[:8 * B].view(B, 8)
for it in range(3):
    non_max_suppression_batched(res, 0.25, 0.45, 300, cand_bitmap=ctx.bitmap(), cand_records=ctx.records(), clear_bitmap=False)
lib.fvb_debug_set_nms_trace(C.c_void_p(trace_all.data_ptr()))
non_max_suppression_batched(res, 0.25, 0.45, 300, cand_bitmap=ctx.bitmap(), cand_records=ctx.records(), clear_bitmap=False)
torch.cuda.synchronize()
lib.fvb_debug_set_nms_trace(C.c_void_p(0))
t = trace.cpu().double()
d = (t[:, 1:6] - t[:, 0:5]) / 1e3  # globaltimer ns -> us
names = ["phase0 bitmap->rows", "phase1 records", "sort", "greedy", "outputs"]
for i, nm in enumerate(names):
    print("%-22s mean %6.1f us  max %6.1f" % (nm, d[:, i].mean().item(), d[:, i].max().item()))
print("total mean %.1f max %.1f; n mean %.0f kept mean %.0f" % ((t[:, 5] - t[:, 0]).mean().item() / 1e3, (t[:, 5] - t[:, 0]).max().item() / 1e3, t[:, 6].mean().item(), t[:, 7].mean().item()))
t0 = t[:, 0].min()
st_, en_ = (t[:, 0] - t0) / 1e3, (t[:, 5] - t0) / 1e3
print("CTA start offsets: min %.1f median %.1f max %.1f us; end: median %.1f max %.1f us" % (st_.min().item(), st_.median().item(), st_.max().item(), en_.median().item(), en_.max().item()))
print("starts sorted (every 16th):", [round(x, 1) for x in st_.sort()[0][::16].tolist()])
