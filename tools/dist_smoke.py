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
log("peer reduce active: %s" % (step._peer() is not None))
nccl_step = ValStep(cfg.anchors_levels(), cfg.strides, batch_global=batch)
nccl_step._peer_reducer = None                     # force the NCCL all-reduce + combine path
o2 = nccl_step(dh, dl)
torch.cuda.synchronize()
log("NCCL-path loss %.6f (peer path %.6f)" % (float(o2["loss"]), loss_sharded))
assert abs(float(o2["loss"]) - loss_sharded) <= 1e-6 * abs(loss_sharded)
assert torch.allclose(o2["partials"], out["partials"], rtol=1e-12)
if step._peer() is not None:
    assert int(step._peer().status[0]) == 0
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
# ---- training side: sharded backward with all-reduced partials == the full-batch gradient of this rank's images --------
import numpy as np  # noqa: E402
from fastvision_b200 import loss as fl  # noqa: E402
from fastvision_b200.dist import allreduce_partials, gather_map_state  # noqa: E402
from fastvision_b200.metrics import CalculateMAP  # noqa: E402


class _M:
    anchors_per_level = cfg.anchors_levels()
    backbone_strides_per_level = cfg.strides


lossf = fl.Yolov3Loss(_M(), 0.5, 0.05, 1.0, 0.5)
with torch.no_grad():
    lossf(dh, dl)
parts = allreduce_partials(lossf.partials.clone())
grads = lossf.backward_heads(dh, dl, None, parts, batch)
log("sharded backward done")
if True:
    full_heads = [h.to(dev).requires_grad_(True) for h in heads]
    fl.Yolov3Loss(_M(), 0.5, 0.05, 1.0, 0.5)(full_heads, labels.to(dev)).sum().backward()
    for gsh, fh in zip(grads, full_heads):
        want = fh.grad[lo:hi]
        err = float((gsh - want).abs().max() / want.abs().max())
        assert err < 1e-5, err
    log("sharded gradients match the full-batch gradients (max rel err < 1e-5)")

# ---- evaluation side: per-rank matcher, all-gather of the evidence, device AP integration on every rank -------------------
thr = np.linspace(0.5, 0.95, 10)
est = CalculateMAP(thr)
dets = step.detections()
for i, d in enumerate(dets):
    est.process_one(d, synth.labels_to_pixel_targets(labels, lo + i, cfg.img, cfg.img).to(dev))
rows, tcls = est.state()
all_rows, all_cls = gather_map_state(rows, tcls)
merged = CalculateMAP(thr)
merged.load_state(all_rows, all_cls)
m_iou, m_cls, ids = merged.fetch()
log("gathered mAP50 %.4f mAP50:95 %.4f over %d classes" % (m_iou[0], m_iou.mean(), len(ids)))
if rank == 0:
    ref = CalculateMAP(thr)
    for i, d in enumerate(single.detections()):
        ref.process_one(d, synth.labels_to_pixel_targets(labels, i, cfg.img, cfg.img).to(dev))
    r_iou, r_cls, r_ids = ref.fetch()
    assert r_ids == ids and np.allclose(r_iou, m_iou, rtol=1e-12) and np.allclose(r_cls, m_cls, rtol=1e-12)
    log("gathered mAP == single-GPU mAP")
dist.barrier()
log("done")
dist.destroy_process_group()
