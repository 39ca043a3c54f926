"""A/B of the per-rank step at a data-parallel shard of 128 YOLOv3-608 / 10-class images (the 8-GPU share of configs[2]) under
torchrun: plain NMS launch vs the chained 1024-thread NMS (FVB_NMS_WIDE_CTA), both with the peer-memory reduce in the graph."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastvision_b200 import synth  # noqa: E402
from fastvision_b200.pipeline import ValStep  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
cfg, per = synth.SHIP608, 128
g = synth.make_generator(3)
labels = synth.make_labels(cfg, per, g)
dh = [h.to(dev) for h in synth.make_heads(cfg, per, labels, g)]
dl = labels.to(dev)


def run(overlap):
    step = ValStep(cfg.anchors_levels(), cfg.strides, batch_global=per * world, overlap_nms=overlap)
    for _ in range(3):
        step(dh, dl)
    torch.cuda.synchronize()
    gph = torch.cuda.CUDAGraph()
    cs = torch.cuda.Stream(device=dev)
    cs.wait_stream(torch.cuda.current_stream())
    with torch.cuda.graph(gph, stream=cs):
        for _ in range(50):
            step._head(dh, dl)
            step._decode(dh)
            step._tail(dh, dl, reduce_inside=True)
    torch.cuda.current_stream().wait_stream(cs)
    gph.replay()
    torch.cuda.synchronize()
    out = []
    for _ in range(3):
        dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        gph.replay()
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / 50], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out.append(float(t))
    return out, step.overlap_nms


for ov in (False, None, False, None):
    ms, used = run(ov)
    if rank == 0:
        print("overlap_nms=%s (in effect: %s): %s ms/step" % (ov, used, ["%.4f" % x for x in ms]), flush=True)
dist.barrier()
dist.destroy_process_group()
