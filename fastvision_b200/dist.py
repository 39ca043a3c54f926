"""Data-parallel plumbing for the hot path: one process per GPU, images sharded contiguously.

Decode and NMS are per image and need no collective.  The loss normalisers are GLOBAL-batch
(loss/yolov3_loss.py:52,58,64): ranks all-reduce the L*4 fp64 partial sums {S_cls, S_box, S_conf, M} and
every rank forms the scalar with the global batch size.  mAP needs the ``correct`` rows and the target
classes of all ranks on the rank that calls ``fetch`` (all-gather of padded rows).
Works with any torch.distributed backend (nccl on GPUs; gloo in the CPU tests of the host logic).
"""
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(batch_global: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice [start, stop) of the global batch owned by ``rank`` (remainder spread over low ranks)."""
    base, rem = divmod(batch_global, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_labels(labels: torch.Tensor, start: int, stop: int) -> torch.Tensor:
    """Rows of ``labels`` [T,6] whose image index lies in [start, stop), re-based to local image indices."""
    keep = (labels[:, 0] >= start) & (labels[:, 0] < stop)
    out = labels[keep].clone()
    out[:, 0] -= start
    return out


def allreduce_partials(partials: torch.Tensor, group=None) -> torch.Tensor:
    """Sum the [L,4] fp64 loss partials over ranks, in place (96 bytes; latency-bound)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(partials, op=dist.ReduceOp.SUM, group=group)
    return partials


def combine_partials_host(partials, cells_per_image_level: List[int], num_classes: int, batch_global: int,
                          ratio_box: float = 0.05, ratio_conf: float = 1.0, ratio_cls: float = 0.5) -> float:
    """Host mirror of ``fvb_yolov3_loss_combine_f32`` (the formula of loss/yolov3_loss.py:49-72 on global sums).

    partials[l] = (S_cls, S_box, S_conf, M); cells_per_image_level[l] = A*H_l*W_l.
    """
    total = 0.0
    for l, cells in enumerate(cells_per_image_level):
        s_cls, s_box, s_conf, m = (float(x) for x in partials[l])
        if m > 0:
            total += ratio_cls * s_cls / (m * num_classes) + ratio_box * s_box / m
        total += ratio_conf * s_conf / (float(batch_global) * cells)
    return total * batch_global


def all_gather_rows(rows: torch.Tensor, group=None) -> torch.Tensor:
    """Concatenate a per-rank [n_r, C] tensor over ranks (rank order), n_r may differ: counts, then padded rows."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return rows
    world = dist.get_world_size(group)
    n = torch.tensor([rows.size(0)], dtype=torch.int64, device=rows.device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    cap = max(max(counts), 1)
    padded = torch.zeros((cap,) + tuple(rows.shape[1:]), dtype=rows.dtype, device=rows.device)
    padded[:rows.size(0)] = rows
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    return torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)


def gather_map_state(correct_rows: torch.Tensor, target_classes: torch.Tensor, group=None):
    """Bring every rank's mAP evidence together: ``correct`` rows [sum M, 2+n_thr] and target classes [sum N]."""
    return all_gather_rows(correct_rows, group), all_gather_rows(target_classes.view(-1, 1), group).view(-1)


# ---- all-reduce + combine of the loss partials over NVLink peer memory ------------------------------------------------------
_peer_reducers = {}


class PeerReducer:
    """Peer-mapped buffers for ``fvb_yolov3_loss_peer_combine_f32``: one small symmetric-memory allocation per process group,
    shared by every step object of the process (all ranks must issue the same sequence of reductions on it)."""

    def __init__(self, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        from . import _lib
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        nbytes = int(_lib.load().fvb_peer_buffer_bytes())
        self.buf = symm_mem.empty((nbytes + 7) // 8 + 8, dtype=torch.float64, device=device)
        self.buf.zero_()
        torch.cuda.synchronize(device)
        self.handle = symm_mem.rendezvous(self.buf, self.group)
        self.ptrs = torch.tensor([int(p) for p in self.handle.buffer_ptrs], dtype=torch.int64, device=device)
        self.status = torch.zeros(1, dtype=torch.int32, device=device)
        dist.barrier(self.group)          # every rank's buffer is zeroed and mapped before anybody publishes
        torch.cuda.synchronize(device)

    def reduce_combine(self, loss_fn, geom, partials, batch_global, out_loss):
        """In place: ``partials`` <- sum over ranks (same bits everywhere), ``out_loss`` <- the scalar."""
        from . import _lib
        lib = _lib.load()
        with torch.cuda.device(partials.device):
            _lib.check(lib.fvb_yolov3_loss_peer_combine_f32(geom, int(batch_global), _lib.dptr(partials), _lib.dptr(self.ptrs),
                                                            self.rank, self.world, float(loss_fn.ratio_box),
                                                            float(loss_fn.ratio_conf), float(loss_fn.ratio_cls),
                                                            _lib.dptr(partials), _lib.dptr(out_loss), _lib.dptr(self.status),
                                                            _lib.stream()), "loss_peer_combine")

    def raise_if_failed(self):
        """Host check of the sticky status word of the peer reductions (one 4-byte D2H read: call it where the loss is read
        anyway).  1 = a rank did not arrive within ~10 s, 2 = ranks desynchronised; either way the losses since are NaN."""
        code = int(self.status.item())
        if code:
            raise RuntimeError("fastvision_b200: peer-memory loss reduction failed (status %d: %s) on rank %d" %
                               (code, "a rank did not arrive" if code == 1 else "ranks desynchronised", self.rank))



def peer_reducer(device, group=None) -> Optional["PeerReducer"]:
    """The process-wide PeerReducer of ``group`` (created collectively on first use), or None when symmetric memory is not
    available / FVB_PEER_REDUCE=0 -- callers then fall back to the NCCL all-reduce + combine."""
    import os
    if os.environ.get("FVB_PEER_REDUCE", "1") == "0":
        return None
    key = (id(group), device.index)
    if key not in _peer_reducers:
        red, err = None, None
        try:
            red = PeerReducer(device, group)
        except Exception as exc:  # symmetric memory unsupported on this system: NCCL path
            err = exc
        # the choice must be the SAME on every rank (one rank on NCCL and the others in the peer kernel would wait for each other
        # forever): agree on the minimum of the success flags
        agreed = torch.tensor([1 if red is not None else 0], dtype=torch.int32, device=device)
        dist.all_reduce(agreed, op=dist.ReduceOp.MIN, group=group)
        if int(agreed.item()) == 0:
            if err is not None or red is not None:
                import warnings
                warnings.warn("fastvision_b200: peer-memory reduce unavailable on at least one rank (%s); all ranks use the NCCL "
                              "all-reduce" % (err if err is not None else "another rank failed",))
            red = None
        _peer_reducers[key] = red
    return _peer_reducers[key]
