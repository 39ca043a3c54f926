"""Race hunt for the decode -> NMS hand-shake (tile_sync counters, programmatic dependent launch): the overlapped step is
run thousands of times on the same inputs -- eagerly and as a CUDA-graph replay, with the outputs scrubbed in between -- and every
run must reproduce the plain (stream-ordered) step bit for bit.  A consumer that read candidate bits or records before they were
visible would show up as a mismatch in some iteration."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastvision_b200 import synth  # noqa: E402
from fastvision_b200.pipeline import ValStep  # noqa: E402

cfg = synth.CONFIGS[os.environ.get("CFG", "yolov3-416-coco")]
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ITERS = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
g = synth.make_generator(2, 11)
labels = synth.make_labels(cfg, B, g)
dh = [h.cuda() for h in synth.make_heads(cfg, B, labels, g)]
dl = labels.cuda()
plain = ValStep(cfg.anchors_levels(), cfg.strides, overlap_nms=False)
ref = {k: v.clone() for k, v in plain(dh, dl).items()}
torch.cuda.synchronize()
valid = torch.arange(ref["boxes"].size(1), device="cuda")[None, :] < ref["cnt"][:, None]
fast = ValStep(cfg.anchors_levels(), cfg.strides, overlap_nms=True)
fast(dh, dl)
assert fast.overlap_nms
replay = fast.capture(dh, dl)
bad = 0
for mode in ("eager", "graph"):
    for it in range(ITERS):
        o = fast.out
        if it % 7 == 0:                      # scrub: stale outputs must not mask a miss
            o["cnt"].fill_(-5)
            o["boxes"].zero_()
            o["scores"].zero_()
        if mode == "eager":
            fast(dh, dl)
        else:
            replay()
        ok = torch.equal(o["cnt"], ref["cnt"]) and torch.equal(o["boxes"][valid], ref["boxes"][valid]) and \
            torch.equal(o["scores"][valid], ref["scores"][valid]) and torch.equal(o["cls"][valid], ref["cls"][valid]) and \
            torch.equal(o["loss"], ref["loss"])
        if not ok:
            bad += 1
            print("MISMATCH", mode, it, int((o["cnt"] != ref["cnt"]).sum()), flush=True)
            if bad > 5:
                raise SystemExit(1)
    torch.cuda.synchronize()
    print("%s: %d iterations at B=%d, mismatches so far %d" % (mode, ITERS, B, bad), flush=True)
assert int(fast.ctx.tile_sync().abs().sum()) == 0 and int(fast.ctx.bitmap().abs().sum()) == 0
print("stress ok" if bad == 0 else "stress FAILED")
sys.exit(0 if bad == 0 else 1)
