"""Stage-by-stage multi-GPU smoke (torchrun): prints progress so a hang can be located from the log."""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastvision_b200 import synth  # noqa: E402
from fastvision_b200.pipeline import ValStep  # noqa: E402
from fastvision_b200.dist import shard_range, shard_labels  # noqa: E402


def log(msg):
    print("[rank %s %.1fs] %s" % (os.environ.get("RANK"), time.time() - T0, msg), flush=True)


T0 = time.time()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
log("init pg")
dist.init_process_group("nccl", device_id=dev)
log("barrier")
dist.barrier()
cfg, batch = synth.COCO416, 16
g = synth.make_generator(1)
labels = synth.make_labels(cfg, batch, g)          # same global batch on every rank
heads = synth.make_heads(cfg, batch, labels, g)
lo, hi = shard_range(batch, rank, world)
dh = [h[lo:hi].contiguous().to(dev) for h in heads]
dl = shard_labels(labels, lo, hi).to(dev)
step = ValStep(cfg.anchors_levels(), cfg.strides, batch_global=batch)
log("eager step")
out = step(dh, dl)
torch.cuda.synchronize()
loss_sharded = float(out["loss"])
log("eager loss %.6f" % loss_sharded)
log("capture")
replay = step.capture(dh, dl)
log("replay")
replay()
torch.cuda.synchronize()
log("graph loss %.6f" % float(out["loss"]))
if rank == 0:
    single = ValStep(cfg.anchors_levels(), cfg.strides)
    single.pg = None
    single._distributed = lambda: False
    o1 = single([h.to(dev) for h in heads], labels.to(dev))
    torch.cuda.synchronize()
    full = float(o1["loss"])
    log("single-GPU full-batch loss %.6f  sharded %.6f  rel diff %.2e" % (full, loss_sharded, abs(full - loss_sharded) / abs(full)))
    assert abs(full - loss_sharded) <= 1e-5 * abs(full)
    k = int(o1["cnt"][0])
    assert torch.equal(o1["boxes"][0, :k], out["boxes"][0, :k])
dist.barrier()
log("done")
dist.destroy_process_group()
