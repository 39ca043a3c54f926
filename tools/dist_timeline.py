"""Device timeline of one data-parallel step per rank (torchrun, eager launches, CUDA events on each stream): when the decode
ends, when the NMS branch ends, and when the loss kernels and the peer reduce of the loss branch end.  Written to find out why
the loss branch (about 45 us of kernels when profiled alone) is the 68 us tail of the step on more than one GPU (DESIGN.md
section 7, "known limits").  Runs with any world size (1 GPU: no reduce; B=256: NMS 61.6 us, loss kernels 64.7 us).

usage: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/dist_timeline.py [batch_per_gpu]
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastvision_b200 import synth  # noqa: E402
from fastvision_b200.detection.tools.nms import non_max_suppression_batched  # noqa: E402
from fastvision_b200.pipeline import ValStep  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
cfg = synth.COCO416
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 256
g = synth.make_generator(1, rank)
labels = synth.make_labels(cfg, batch, g)
dh = [h.to(dev) for h in synth.make_heads(cfg, batch, labels, g)]
dl = labels.to(dev)
step = ValStep(cfg.anchors_levels(), cfg.strides, batch_global=batch * world)
for _ in range(5):
    step(dh, dl)
torch.cuda.synchronize()
dist.barrier()

E = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
rows = []
for it in range(6):
    ev = {k: E() for k in ("t0", "decoded", "nms_end", "loss_begin", "loss_end", "reduce_end")}
    main, side = torch.cuda.current_stream(), step._side
    ctx, o = step.ctx, step.out
    dist.barrier()
    torch.cuda.synchronize()
    ev["t0"].record(main)
    with torch.cuda.stream(side):
        side.wait_event(ev["t0"])
        step.loss_fn.match(dh, dl, ctx)              # beside the decode
    step._decode(dh)
    ev["decoded"].record(main)
    with torch.cuda.stream(side):
        side.wait_event(ev["decoded"])
        ev["loss_begin"].record(side)
        step.loss_fn.finish(dl.size(0), ctx, ctx.bce0(), out=o["loss"], partials=o["partials"])
        ev["loss_end"].record(side)
        step._reduce()
        ev["reduce_end"].record(side)
    step._nms()
    ev["nms_end"].record(main)
    main.wait_event(ev["reduce_end"])
    torch.cuda.synchronize()
    f = lambda k: ev["t0"].elapsed_time(ev[k]) * 1e3  # noqa: E731
    rows.append([f(k) for k in ("decoded", "nms_end", "loss_begin", "loss_end", "reduce_end")])
mine = torch.tensor(rows[1:], dtype=torch.float64, device=dev).median(0).values
allr = [torch.zeros_like(mine) for _ in range(world)]
dist.all_gather(allr, mine)
if rank == 0:
    print("us after the step's start (median of 5 eager steps; eager launches add host gaps a graph replay does not have)")
    print("rank  decode_end  nms_end  loss_begin  loss_kernels_end  reduce_end   nms  loss_kernels  reduce")
    for r, v in enumerate(allr):
        d, n, lb, le, re_ = [float(x) for x in v]
        print("%4d  %10.1f  %7.1f  %10.1f  %16.1f  %10.1f  %5.1f  %12.1f  %6.1f" % (r, d, n, lb, le, re_, n - d, le - lb, re_ - le))
dist.destroy_process_group()
