import os, sys, torch
sys.path.insert(0, "/root/repo")
from fastvision_b200 import synth
from fastvision_b200.detection.models import yolov3_decode, DecodeContext
from fastvision_b200.pipeline import ValStep
sys.path.insert(0, "/root/repo/tools")
from microbench import timeit
for name, B in (("yolov3-608-ship", 1024), ("yolov3-416-coco", 256)):
    cfg = synth.CONFIGS[name]
    g = synth.make_generator(3)
    labels = synth.make_labels(cfg, B, g)
    dh = [h.cuda() for h in synth.make_heads(cfg, B, labels, g)]
    dl = labels.cuda()
    anc, st = cfg.anchors_levels(), cfg.strides
    ctx = DecodeContext(dh, anc, st)
    res = torch.empty(B, ctx.rows, ctx.k, device="cuda")
    plain = timeit(lambda: yolov3_decode(dh, anc, st, ctx=ctx, out=res), iters=30)[0]
    fused = timeit(lambda: (ctx.bitmap().zero_(), yolov3_decode(dh, anc, st, ctx=ctx, out=res, conf_thres=0.25, want_bce0=True)), iters=30)[0]
    step = ValStep(anc, st)
    step(dh, dl)
    rep = step.capture(dh, dl)
    full = timeit(rep, iters=30)[0]
    step2 = ValStep(anc, st, overlap_nms=False)
    step2(dh, dl)
    rep2 = step2.capture(dh, dl)
    full2 = timeit(rep2, iters=30)[0]
    print(name, B, "plain %.4f fused(+memset) %.4f step %.4f (overlap=%s) step-no-overlap %.4f" % (plain, fused, full, step.overlap_nms, full2), flush=True)
    del dh, res, step, step2, rep, rep2, ctx
    torch.cuda.empty_cache()
