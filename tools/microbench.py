"""Per-kernel timings of the hot path on one GPU (CUDA events on the launching stream)."""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastvision_b200 import synth, _lib  # noqa: E402
from fastvision_b200.detection.models import yolov3_decode, DecodeContext  # noqa: E402
from fastvision_b200.detection.tools import non_max_suppression_batched  # noqa: E402
from fastvision_b200.loss import Yolov3Loss  # noqa: E402
from fastvision_b200.pipeline import ValStep, ValPipeline  # noqa: E402


def timeit(fn, iters=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="yolov3-416-coco")
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--stress", action="store_true")
    args = ap.parse_args()
    cfg = synth.CONFIGS[args.config]
    g = synth.make_generator(2)
    t0 = time.time()
    labels = synth.make_labels(cfg, args.batch, g)
    heads = synth.make_heads(cfg, args.batch, labels, g, stress=args.stress)
    print("generated in %.1fs; labels %d" % (time.time() - t0, labels.size(0)), flush=True)
    dh, dl = [h.cuda() for h in heads], labels.cuda()
    anc, st = cfg.anchors_levels(), cfg.strides
    ctx = DecodeContext(dh, anc, st)
    res = torch.empty(args.batch, ctx.rows, ctx.k, device="cuda")
    bytes_alg = 2 * args.batch * ctx.rows * ctx.k * 4
    out = {}

    def rep(name, fn, nbytes=None):
        med, mn = timeit(fn)
        out[name] = {"ms_median": med, "ms_min": mn}
        if nbytes:
            out[name]["GBps"] = nbytes / med / 1e6
        print(name, out[name], flush=True)

    rep("decode_plain", lambda: yolov3_decode(dh, anc, st, ctx=ctx, out=res), bytes_alg)
    rep("decode_precise", lambda: yolov3_decode(dh, anc, st, ctx=ctx, out=res, precise=True), bytes_alg)
    rep("decode_fused", lambda: (ctx.bitmap().zero_(), yolov3_decode(dh, anc, st, ctx=ctx, out=res, conf_thres=0.25, want_bce0=True)), bytes_alg)
    rep("torch_copy_same_bytes", lambda: res.copy_(res.view(-1).roll(0).view_as(res)) if False else res.view(-1)[: res.numel() // 2].copy_(res.view(-1)[res.numel() // 2:]), bytes_alg // 2)
    ctx.bitmap().zero_()
    yolov3_decode(dh, anc, st, ctx=ctx, out=res, conf_thres=0.25, want_bce0=True)
    cand = (res[..., 4] > 0.25).sum(1)
    print("candidates/img mean %.1f max %d" % (cand.float().mean().item(), cand.max().item()))
    rep("nms_bitmap", lambda: non_max_suppression_batched(res, 0.25, 0.45, 300, cand_bitmap=ctx.bitmap(), cand_records=ctx.records(), clear_bitmap=False))
    rep("nms_standalone", lambda: non_max_suppression_batched(res, 0.25, 0.45, 300))

    class M:
        anchors_per_level = anc
        backbone_strides_per_level = st
    lossf = Yolov3Loss(M(), 0.5, 0.05, 1.0, 0.5)
    rep("loss_fused", lambda: lossf(dh, dl, conf_bce0=ctx.bce0(), ctx=ctx))
    rep("loss_standalone", lambda: lossf(dh, dl, ctx=ctx))
    step = ValStep(anc, st)
    rep("step_eager", lambda: step(dh, dl), bytes_alg)
    replay = step.capture(dh, dl)
    rep("step_graph", replay, bytes_alg)
    print("kept/img mean %.1f" % step.out["cnt"].float().mean().item(), "loss", step.out["loss"].item())
    pipe = ValPipeline(anc, st)
    for _ in range(5):
        pipe.submit(dh, dl)
    pipe.flush()
    torch.cuda.synchronize()
    for k in (50, 200):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(k):
            pipe.submit(dh, dl)
        t_cpu = time.perf_counter() - t0
        pipe.flush()
        e1.record()
        e1.synchronize()
        ms = e0.elapsed_time(e1) / k
        out["step_pipelined_%d" % k] = {"ms_per_step": ms, "GBps": bytes_alg / ms / 1e6, "cpu_submit_ms": t_cpu / k * 1e3}
        print("step_pipelined", k, out["step_pipelined_%d" % k], flush=True)
    o = pipe.steps[(pipe.count - 1) % 2].out
    print("pipelined kept/img mean %.1f" % o["cnt"].float().mean().item(), "loss", o["loss"].item())
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/microbench_%s_b%d%s.json" % (args.config, args.batch, "_stress" if args.stress else ""), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
