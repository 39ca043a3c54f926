"""BASELINE configs[2]: YOLOv3-608, 10 classes, GLOBAL batch 1024 sharded over the ranks (strong scaling).  Run directly for one
GPU or under torchrun for G GPUs; rank 0 prints one JSON line (images/s = 1024 / max-over-ranks step time)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastvision_b200 import synth  # noqa: E402
from fastvision_b200.pipeline import ValStep  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
GLOBAL, STEPS = 1024, int(os.environ.get("STEPS", 100))
cfg = synth.SHIP608
per = GLOBAL // world
g = synth.make_generator(3, rank)                      # every rank generates its own share (same distribution)
labels = synth.make_labels(cfg, per, g)
dh = [h.to(dev) for h in synth.make_heads(cfg, per, labels, g)]
dl = labels.to(dev)
step = ValStep(cfg.anchors_levels(), cfg.strides, batch_global=GLOBAL)
step(dh, dl)
replay = step.capture(dh, dl)
for _ in range(10):
    replay()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(STEPS):
    replay()
b.record()
torch.cuda.synchronize()
t = torch.tensor([a.elapsed_time(b) / STEPS], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
ms = float(t.item())
if rank == 0:
    alg = 2 * per * step.ctx.rows * step.ctx.k * 4
    print(json.dumps({"config": cfg.name, "global_batch": GLOBAL, "n_gpus": world, "images_per_gpu": per, "ms_per_step": ms,
                      "images_per_s": GLOBAL / ms * 1e3, "algorithmic_GBps_per_gpu": alg / ms / 1e6,
                      "loss": float(step.out["loss"]), "kept_mean": float(step.out["cnt"].float().mean())}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
