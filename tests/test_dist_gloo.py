"""World-size-2 gloo tests (CPU) of the data-parallel host logic: image sharding, the all-reduce of the per-level
loss partials + global-batch combine, and the gather of mAP evidence.  The per-shard partials come from the oracle
here (no GPU in this suite); on the GPU box the same plumbing carries the CUDA kernel's partials (bench.py --gpus N)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from small_cfg import SMALL
from fastvision_b200 import synth
from fastvision_b200 import dist as fd


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    import oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg, batch = SMALL, 5                                     # odd batch: uneven shards
    g = torch.Generator().manual_seed(99)
    labels = synth.make_labels(cfg, batch, g)
    heads = synth.make_heads(cfg, batch, labels, g)
    lo, hi = fd.shard_range(batch, rank, world)
    my_heads = [h[lo:hi].contiguous() for h in heads]
    my_labels = fd.shard_labels(labels, lo, hi)
    _, parts = oracle.loss.yolov3_loss(my_heads, my_labels, cfg.anchors_levels(), cfg.strides, return_partials=True)
    p = torch.tensor(parts, dtype=torch.float64)
    fd.allreduce_partials(p)
    cells = [cfg.anchors_per_level * f * f for f in cfg.feat]
    loss = fd.combine_partials_host(p, cells, cfg.num_classes, batch)
    # mAP evidence: ragged rows per rank
    rows = torch.full((rank + 2, 12), float(rank), dtype=torch.float64)
    cls = torch.arange(rank + 1, dtype=torch.float32)
    all_rows, all_cls = fd.gather_map_state(rows, cls)
    if rank == 0:
        want = float(oracle.loss.yolov3_loss(heads, labels, cfg.anchors_levels(), cfg.strides)[0])
        q.put((loss, want, tuple(all_rows.shape), all_rows[:, 0].tolist(), all_cls.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_loss_and_map_gather_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    loss, want, shape, col0, cls = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    np.testing.assert_allclose(loss, want, rtol=1e-5)        # mean of shard means would NOT pass this
    assert shape == (5, 12) and col0 == [0.0, 0.0, 1.0, 1.0, 1.0]
    assert cls == [0.0, 0.0, 1.0]


def test_shard_range_covers_batch():
    for batch, world in [(256, 8), (5, 2), (7, 4), (3, 8)]:
        spans = [fd.shard_range(batch, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == batch
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))


def test_shard_labels_rebases():
    labels = torch.tensor([[0, 1, .5, .5, .1, .1], [2, 3, .5, .5, .1, .1], [3, 0, .2, .2, .1, .1]])
    out = fd.shard_labels(labels, 2, 4)
    assert out[:, 0].tolist() == [0.0, 1.0] and out[:, 1].tolist() == [3.0, 0.0]


def _agree_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import warnings

    class _Fake:                                   # stands in for a PeerReducer whose construction worked on rank 0 only
        def __init__(self, device, group=None):
            if dist.get_rank() != 0:
                raise RuntimeError("symmetric memory unavailable on this rank")

    fd.PeerReducer = _Fake
    fd._peer_reducers.clear()
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        red = fd.peer_reducer(torch.device("cpu"), None)
    q.put((rank, red is None, len(w)))
    dist.barrier()
    dist.destroy_process_group()


def test_peer_vs_nccl_choice_is_agreed_group_wide():
    """If the peer-memory reducer cannot be built on ONE rank, EVERY rank must fall back to the NCCL all-reduce (a rank alone in the
    peer kernel would wait for the others forever): the choice is the MIN of the per-rank success flags."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_agree_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [g[1] for g in got] == [True, True]                 # rank 0's own success does not count
    assert all(g[2] >= 1 for g in got)                         # and both ranks say why
