#!/usr/bin/env python
"""bench.py -- images/s of the fastvision detection hot path (decode + NMS + loss) on N B200s.

Contract (see the task brief / BASELINE.json):
    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path (oracle port)
For N > 1 it is launched under torchrun (one rank per GPU, NCCL); rank 0 prints ONE JSON line.

A "step" is one pass of the hot path over one batch of synthetic head tensors: YOLOv3 decode of the three
raw head levels, confidence filter + NMS for every image, and Yolov3Loss (CIoU box + objectness/class BCE).
Workload at N=1 = BASELINE.json configs[1]: YOLOv3-416, COCO shape (80 classes, 3x3 anchors), batch 256.
Under N ranks every rank owns 256 images (weak scaling); the loss partial sums are all-reduced (96 bytes).
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from fastvision_b200 import synth  # noqa: E402

METRIC = "images/sec YOLOv3-416 decode+NMS+loss"
WORKLOAD = "YOLOv3-416 COCO-shape (80 cls, 3x3 anchors) decode+CIoU loss+NMS, batch 256 per GPU (BASELINE.json configs[1])"
UNIT = "images/s"
CONFIG_ID = 2


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def workload(cfg, batch, rank):
    g = synth.make_generator(CONFIG_ID, rank)
    labels = synth.make_labels(cfg, batch, g)
    heads = synth.make_heads(cfg, batch, labels, g)
    return labels, heads


def measured_peak_hbm():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md, 6.65 TB/s)"


class ClockSampler:
    """Samples SM clock / throttle reasons of one GPU with NVML while the timed region runs."""

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
            nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.005)

    def __enter__(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr is not None:
            self._thr.join(timeout=1.0)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------ CPU reference leg
def cpu_reference_pass(heads, labels, cfg, nms_backend):
    """One pass of the reference's CPU path on a batch: decode -> loss -> per-image NMS loop (utils/fit.py:86-95)."""
    import oracle  # test infrastructure; allowed here as the cpu_baseline / --impl reference leg only
    anc, st = cfg.anchors_levels(), cfg.strides
    res = oracle.decode.decode(heads, anc, st)
    loss = oracle.loss.yolov3_loss(heads, labels, anc, st)
    kept = 0
    for i in range(res.size(0)):
        s, c, b = oracle.nms.nms_lib(res[i], 0.25, 0.45, 300, backend=nms_backend)
        kept += s.size(0)
    return float(loss), kept


def pick_nms_backend():
    try:
        import torchvision  # noqa: F401  the op the reference itself calls (detection/tools/NMS.py:18), CPU build
        return "torchvision"
    except Exception:
        return "numpy"


def time_cpu_reference(cfg, sample_batch, rank, budget_s, min_reps, warm):
    labels, heads = workload(cfg, sample_batch, rank)
    backend = pick_nms_backend()
    for _ in range(warm):
        cpu_reference_pass(heads, labels, cfg, backend)
    times = []
    t_start = time.perf_counter()
    while len(times) < min_reps or (time.perf_counter() - t_start) < budget_s:
        t0 = time.perf_counter()
        cpu_reference_pass(heads, labels, cfg, backend)
        times.append(time.perf_counter() - t0)
        if len(times) >= 200:
            break
    med = statistics.median(times)
    return sample_batch / med, med, len(times), backend


def run_reference(args, cfg):
    rank, world, _ = dist_env()
    if rank != 0:
        return  # the CPU arm runs on rank 0 only
    torch.set_num_threads(os.cpu_count() or 1)
    sample = args.cpu_sample
    labels, heads = workload(cfg, sample, 0)
    backend = pick_nms_backend()
    for _ in range(max(args.warmup, 1) if args.warmup < 3 else 3):
        cpu_reference_pass(heads, labels, cfg, backend)
    steps = max(1, min(args.steps, 40))
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_reference_pass(heads, labels, cfg, backend)
    dt = time.perf_counter() - t0
    value = sample * steps / dt
    sample_desc = ("%d-image slice of the %s batch (same generator and seed), decode -> Yolov3Loss -> per-image "
                   "non_max_suppression loop, oracle port with NMS backend '%s'" % (sample, cfg.name, backend))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_gpu": args.batch, "global_batch": args.batch * world,
                   "rows_per_image": cfg.cells, "channels": cfg.k, "conf_thres": 0.25, "iou_thres": 0.45, "max_det": 300,
                   "loss_ratios": [0.05, 1.0, 0.5],
                   "step_sample_images": sample, "timed_on": "host CPU",
                   "note": "each step is a %d-image sample of the %d-image workload (same generator and seed), images/s-normalised" % (sample, args.batch)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                         "sample": sample_desc, "cpu_model": cpu_model_name(), "stages": cpu_stage_baseline()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def cpu_model_name():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.lower().startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def cpu_stage_baseline(reps=7, warm=2):
    """the reference's CPU path on config 1 (B=8, 416, C=80), per stage, with all host threads and with
    one.  When /root/reference is mounted (build container) the stages call the UNMODIFIED reference through oracle/ref_shim.py
    (kind "reference"); on the GPU box that tree does not exist and the oracle port is timed (kind "port")."""
    import numpy as np
    import oracle
    from oracle import ref_shim
    cfg, batch = synth.COCO416, 8
    g = synth.make_generator(1)
    labels = synth.make_labels(cfg, batch, g)
    heads = synth.make_heads(cfg, batch, labels, g)
    anc, st = cfg.anchors_levels(), cfg.strides
    backend = pick_nms_backend()
    thr = np.linspace(0.5, 0.95, 10)
    kind = "port"
    if ref_shim.available():
        try:
            ref = ref_shim.load()
            kind = "reference"
        except Exception:
            ref = None
    if kind == "reference":
        class _M:
            anchors_per_level, backbone_strides_per_level = anc, st
        loss_fn = ref.Yolov3Loss(_M(), 0.5, 0.05, 1.0, 0.5)
        decode = lambda: ref.decode(heads, anc, st, cfg.num_classes)                                # noqa: E731
        loss = lambda: loss_fn(heads, labels)                                                      # noqa: E731
        nms_one = lambda r: ref.tools.non_max_suppression(r, 0.25, 0.45, 300)                       # noqa: E731
        new_map = lambda: ref.metrics.CalculateMAP(thr)                                            # noqa: E731
    else:
        decode = lambda: oracle.decode.decode(heads, anc, st)                                      # noqa: E731
        loss = lambda: oracle.loss.yolov3_loss(heads, labels, anc, st)                             # noqa: E731
        nms_one = lambda r: oracle.nms.nms_lib(r, 0.25, 0.45, 300, backend=backend)                 # noqa: E731
        new_map = lambda: oracle.map_.MapOracle(thr)                                               # noqa: E731
    res = decode()
    dets = [nms_one(res[i]) for i in range(batch)]
    tgts = [synth.labels_to_pixel_targets(labels, i, cfg.img, cfg.img) for i in range(batch)]

    def nms_loop():
        for i in range(batch):
            nms_one(res[i])

    def map_stage():                      # utils/fit.py:96-103: per-image process_one, then fetch
        est = new_map()
        for (s_, c_, b_), t in zip(dets, tgts):
            if s_.numel():
                est.process_one(torch.cat([c_.float(), s_, b_], 1), t)
        est.fetch()

    stages = {"decode": decode, "loss": loss, "nms_loop": nms_loop, "map": map_stage}
    out = {"config": "BASELINE.json configs[0]: YOLOv3-416 COCO-shape, batch 8", "kind": kind, "cpu_model": cpu_model_name(),
           "nms_backend": "torchvision.ops.nms (CPU)" if (kind == "reference" or backend == "torchvision") else "numpy restatement",
           "reps": "median of %d after %d warm-ups" % (reps, warm), "threads": {}}
    prev = torch.get_num_threads()
    for n in sorted({os.cpu_count() or 1, 1}, reverse=True):
        torch.set_num_threads(n)
        row, total = {}, 0.0
        for name, fn in stages.items():
            for _ in range(warm):
                fn()
            ts = []
            for _ in range(reps):
                t0 = time.perf_counter()
                fn()
                ts.append(time.perf_counter() - t0)
            med = statistics.median(ts)
            total += med
            row[name] = {"ms": med * 1e3, "images_per_s": batch / med}
        row["total"] = {"ms": total * 1e3, "images_per_s": batch / total}
        out["threads"][str(n)] = row
    torch.set_num_threads(prev)
    return out


def bind_to_gpu_numa(index):
    """Pin this process to the CPUs nearest to GPU `index` BEFORE any pinned host buffer is allocated (first-touch places the pages
    on that NUMA node), so that N ranks do not all stream their H2D copies out of node 0.  Returns the affinity size or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(index))
        return len(os.sched_getaffinity(0))
    except Exception:
        return None


def decode_source_sha():
    import hashlib
    h = hashlib.sha256()
    for name in ("decode.cu", "common.cuh"):
        with open(os.path.join(ROOT, "fastvision_b200", "csrc", name), "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def measured_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per decode launch from the tracked ncu capture -- only if that capture was
    taken from THIS decode kernel source (profiles/decode_traffic.json is keyed by the source hash; tools/update_traffic.py)."""
    try:
        with open(os.path.join(ROOT, "profiles", "decode_traffic.json")) as f:
            rec = json.load(f)
        if rec.get("source_sha16") == decode_source_sha():
            return rec.get("dram_bytes_per_launch"), rec.get("capture")
        return None, "stale: profiles/decode_traffic.json was captured from another decode.cu (%s)" % rec.get("source_sha16")
    except Exception:
        return None, "no capture"


def time_step_loop(step, dh, dl, k, dev, barrier, in_graph, distributed):
    """ms per step of k back-to-back steps (device events); one CUDA graph when `in_graph`, eager launches otherwise."""
    for _ in range(3):
        step(dh, dl)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g = None
    if in_graph:
        g = torch.cuda.CUDAGraph()
        cs = torch.cuda.Stream(device=dev)
        cs.wait_stream(torch.cuda.current_stream())
        with torch.cuda.graph(g, stream=cs):
            for _ in range(k):
                step._head(dh, dl)
                step._decode(dh)
                step._tail(dh, dl, reduce_inside=distributed)
        torch.cuda.current_stream().wait_stream(cs)
        g.replay()
        torch.cuda.synchronize()
    barrier()
    t0.record()
    if g is not None:
        g.replay()
    else:
        for _ in range(k):
            step(dh, dl)
    t1.record()
    barrier()
    return t0.elapsed_time(t1) / k


def strong_config3(rank, world, dev, dist, barrier, k=50):
    """BASELINE.json configs[2]: ONE global batch of 1024 YOLOv3-608 / 10-class images, first on rank 0 alone, then sharded
    1024/N per rank (the reference's only parallel mode is a scatter of the batch: nn.DataParallel,
    demos/yolov3_huaweiShip/train.py:104).  The 1024 images are 1024/128 copies of one seeded 128-image block (generation on
    the CPU is what bounds the bench's run time); every copy is a distinct tensor region, so nothing is served from cache."""
    from fastvision_b200.pipeline import ValStep
    cfg, total = synth.SHIP608, 1024
    if total % world:
        return {"skipped": "1024 images do not divide over %d ranks" % world}
    per = total // world
    block = 128 if per % 128 == 0 else per
    g = synth.make_generator(3)
    lab_b = synth.make_labels(cfg, block, g)
    hb = [h.to(dev) for h in synth.make_heads(cfg, block, lab_b, g)]

    def tiled(copies):
        heads = [h.repeat(copies, 1, 1, 1, 1).contiguous() for h in hb]
        labs = []
        for c in range(copies):
            t = lab_b.clone()
            t[:, 0] += c * block
            labs.append(t)
        return heads, torch.cat(labs).to(dev)

    res = {"workload": "YOLOv3-608, 10 classes, global batch 1024 (BASELINE.json configs[2])", "steps": k}
    full_loss = full_cnt = None
    if rank == 0:
        fh, fl_ = tiled(total // block)
        full = ValStep(cfg.anchors_levels(), cfg.strides, data_parallel=False)
        res["ms_full_1gpu"] = time_step_loop(full, fh, fl_, k, dev, lambda: torch.cuda.synchronize(), True, False)
        full_loss = float(full.out["loss"])
        full_cnt = full.out["cnt"][:per].clone()
        full_boxes = full.out["boxes"][:per].clone()
        del full, fh, fl_
        torch.cuda.empty_cache()
    barrier()
    sh, sl = tiled(per // block)
    step = ValStep(cfg.anchors_levels(), cfg.strides, batch_global=total)
    step(sh, sl)
    in_graph = step._peer() is not None
    ms = time_step_loop(step, sh, sl, k, dev, barrier, in_graph, True)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res["ms_sharded"] = float(t.item())
    res["images_per_rank"] = per
    res["launch"] = "one CUDA graph of %d steps (peer-memory reduce inside)" % k if in_graph else "eager launches (NCCL all-reduce)"
    if rank == 0:
        res["speedup"] = res["ms_full_1gpu"] / res["ms_sharded"]
        res["images_per_s"] = total / (res["ms_sharded"] * 1e-3)
        rel = abs(float(step.out["loss"]) - full_loss) / abs(full_loss)
        same = bool(torch.equal(step.out["cnt"], full_cnt))
        for i in range(per):
            c = int(full_cnt[i])
            same = same and bool(torch.equal(step.out["boxes"][i, :c], full_boxes[i, :c]))
        res["parity"] = {"loss_rel_diff_sharded_vs_full": rel, "detections_bit_equal": same}
        if rel > 1e-6 or not same:
            raise SystemExit("bench.py: config-3 sharded step differs from the full-batch step: %s" % res["parity"])
    del step, sh, sl
    torch.cuda.empty_cache()
    return res


def other_configs_1gpu(dev, peak):
    """BASELINE.json configs[2] and configs[3] on this one GPU, outside the timed region, so that a single-GPU bench record
    carries every config's number: YOLOv3-608 / 10 classes / B=1024 step (the 1024 images are 8 copies of one seeded 128-image
    block, as in strong_config3) and the Faster R-CNN RPN proposal filter at B=64 (demos/faster_rcnn/models/rpn.py:168-208)."""
    from fastvision_b200.pipeline import ValStep
    from fastvision_b200.detection import tools as ft
    res = {}
    cfg, total, block = synth.SHIP608, 1024, 128
    g = synth.make_generator(3)
    lab_b = synth.make_labels(cfg, block, g)
    hb = [h.to(dev) for h in synth.make_heads(cfg, block, lab_b, g)]
    heads = [h.repeat(total // block, 1, 1, 1, 1).contiguous() for h in hb]
    labs = []
    for c in range(total // block):
        t = lab_b.clone()
        t[:, 0] += c * block
        labs.append(t)
    labels = torch.cat(labs).to(dev)
    step = ValStep(cfg.anchors_levels(), cfg.strides, data_parallel=False)
    ms = time_step_loop(step, heads, labels, 30, dev, lambda: torch.cuda.synchronize(), True, False)
    alg = 2 * total * step.ctx.rows * step.ctx.k * 4
    res["config3_1gpu"] = {"workload": "YOLOv3-608, 10 classes, batch 1024 (BASELINE.json configs[2]) on one GPU", "ms_per_step": ms,
                           "images_per_s": total / (ms * 1e-3), "algorithmic_GBps": alg / (ms * 1e-3) / 1e9,
                           "frac_of_peak": alg / (ms * 1e-3) / 1e9 / peak, "nms_under_decode": bool(step.overlap_nms)}
    del step, heads, labels, hb
    torch.cuda.empty_cache()
    gen = synth.make_generator(4)
    cls, reg = synth.make_rpn_inputs(64, 50, 50, 9, gen)
    dc, dr = cls.to(dev), reg.to(dev)
    base = ft.get_base_anchor([128, 256, 512], [1, 0.5, 2]) / 16
    rpn = {}
    for pre, post in ((12000, 2000), (6000, 300), (2000, 2000)):
        for _ in range(3):
            ft.filter_proposals_batched(dc, dr, base, pre, post, 0.7)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            ft.filter_proposals_batched(dc, dr, base, pre, post, 0.7)
        b.record()
        torch.cuda.synchronize()
        t = a.elapsed_time(b) / 10
        rpn["pre%d_post%d" % (pre, post)] = {"ms": t, "images_per_s": 64 / (t * 1e-3)}
    res["config4_rpn"] = dict(rpn, workload="Faster R-CNN RPN proposal filter, B=64, 22 500 anchors/image, nms 0.7 (BASELINE.json configs[3])")
    return res


def dp_parity(step, dh, dl, batch, rank, world, dev, dist, cfg):
    """Outside the timed region: the sharded step against ONE single-GPU step over the gathered global batch (rank 0).
    loss/yolov3_loss.py:52,58,64 normalise by GLOBAL counts, so the sharded loss must equal the full-batch loss (rtol 1e-6) and
    every image's detections must be bit-equal.  Any difference fails the run."""
    from fastvision_b200.pipeline import ValStep
    out = step(dh, dl)
    torch.cuda.synchronize()
    gh = []
    for h in dh:
        buf = torch.empty((world * h.size(0),) + tuple(h.shape[1:]), dtype=h.dtype, device=dev)
        dist.all_gather_into_tensor(buf, h.contiguous())
        gh.append(buf)
    nlab = torch.tensor([dl.size(0)], dtype=torch.int64, device=dev)
    counts = [torch.zeros_like(nlab) for _ in range(world)]
    dist.all_gather(counts, nlab)
    counts = [int(c.item()) for c in counts]
    pad = torch.zeros(max(counts), 6, dtype=dl.dtype, device=dev)
    pad[:dl.size(0)] = dl
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    labs = []
    for r, (p_, c) in enumerate(zip(parts, counts)):
        t = p_[:c].clone()
        t[:, 0] += r * batch
        labs.append(t)
    keys = ("cnt", "boxes", "scores", "cls")
    gathered = {}
    for key in keys:
        t = out[key].contiguous()
        buf = torch.empty((world * t.size(0),) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
        dist.all_gather_into_tensor(buf, t)
        gathered[key] = buf
    res = None
    if rank == 0:
        full = ValStep(cfg.anchors_levels(), cfg.strides, data_parallel=False)
        fo = full(gh, torch.cat(labs))
        torch.cuda.synchronize()
        rel = abs(float(out["loss"]) - float(fo["loss"])) / abs(float(fo["loss"]))
        ok = bool(torch.equal(gathered["cnt"], fo["cnt"]))
        valid = torch.arange(fo["boxes"].size(1), device=dev)[None, :] < fo["cnt"][:, None]
        for key in ("boxes", "scores", "cls"):
            ok = ok and bool(torch.equal(gathered[key][valid], fo[key][valid]))
        res = {"status": "ok" if (ok and rel <= 1e-6) else "FAILED", "loss_rel_diff": rel, "detections_bit_equal": ok,
               "images_compared": world * batch,
               "how": "all ranks' heads/labels all-gathered to rank 0, one single-GPU step over the %d-image global batch" % (world * batch)}
        del full
    del gh, gathered
    torch.cuda.empty_cache()
    flag = torch.tensor([1 if (res is None or res["status"] == "ok") else 0], device=dev)
    dist.broadcast(flag, 0)
    if int(flag.item()) != 1:
        raise SystemExit("bench.py: dp_parity FAILED: %s" % (res,))
    return res


def config5_map(step, labels, cfg, batch, rank, world, dev, distributed, barrier):
    """BASELINE.json configs[4]: mAP@[.5:.95] over 5000 synthetic images, 5000/N per rank: per-rank matcher (one launch), the
    evidence gathered over NCCL (dist.gather_map_state), AP integration once on rank 0 -- utils/fit.py:94-103,
    metrics/map.py:120-141.  Detections are this rank's NMS output of the timed batch, image j re-using batch image j mod B with
    a seeded per-image box jitter; checked against a single-GPU evaluation of the gathered raw detections."""
    import numpy as np
    from fastvision_b200.metrics import CalculateMAP
    from fastvision_b200 import dist as fdist
    total = 5000
    lo, hi = fdist.shard_range(total, rank, world)
    n_img = hi - lo
    out = step.out
    gen = torch.Generator(device=dev)
    gen.manual_seed(20220504 + 5000)
    jitter_all = (torch.rand(total, 1, 4, generator=gen, device=dev) - 0.5) * 4.0       # +-2 px, keyed by GLOBAL image id
    src = torch.arange(lo, hi, device=dev) % batch
    cnt = out["cnt"][src].long()
    md = out["boxes"].size(1)
    valid = torch.arange(md, device=dev)[None, :] < cnt[:, None]
    boxes = out["boxes"][src] + jitter_all[lo:hi]
    dets = torch.cat([out["cls"][src].float().unsqueeze(-1), out["scores"][src].unsqueeze(-1), boxes], 2)[valid].contiguous()
    det_off = torch.zeros(n_img + 1, dtype=torch.int32, device=dev)
    det_off[1:] = torch.cumsum(cnt, 0)
    lab = labels.to(dev)
    per_img = torch.bincount(lab[:, 0].long(), minlength=batch)
    starts = torch.cumsum(per_img, 0) - per_img
    half = lab[:, 4:6] / 2
    gt_all = torch.cat([lab[:, 1:2], (lab[:, 2:4] - half) * cfg.img, (lab[:, 2:4] + half) * cfg.img], 1)
    g_cnt = per_img[src]
    gt_off = torch.zeros(n_img + 1, dtype=torch.int32, device=dev)
    gt_off[1:] = torch.cumsum(g_cnt, 0)
    idx = torch.repeat_interleave(starts[src], g_cnt) + (torch.arange(int(g_cnt.sum()), device=dev) - torch.repeat_interleave(gt_off[:-1].long(), g_cnt))
    gts = gt_all[idx].contiguous()
    thr = np.linspace(0.5, 0.95, 10)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    est = CalculateMAP(thr)
    est.process_batch(dets, det_off, gts, gt_off)                # warm-up (workspace)
    est = CalculateMAP(thr)
    barrier()
    ev[0].record()
    est.process_batch(dets, det_off, gts, gt_off)
    ev[1].record()
    rows, tcls = est.state()
    if distributed:
        rows, tcls = fdist.gather_map_state(rows, tcls)
    ev[2].record()
    result = None
    if rank == 0:
        est.load_state(rows, tcls)
        result = est.fetch()
    ev[3].record()
    torch.cuda.synchronize()
    times = torch.tensor([ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3])], dtype=torch.float64, device=dev)
    if distributed:
        import torch.distributed as dist
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
        all_dets = fdist.all_gather_rows(torch.cat([dets, torch.repeat_interleave(torch.arange(lo, hi, device=dev), cnt).float()[:, None]], 1))
        all_gts = fdist.all_gather_rows(torch.cat([gts, torch.repeat_interleave(torch.arange(lo, hi, device=dev), g_cnt).float()[:, None]], 1))
    else:
        all_dets = all_gts = None
    res = None
    if rank == 0:
        parity = "n/a (one rank)"
        if distributed:
            d_img, g_img = all_dets[:, 6].long(), all_gts[:, 5].long()
            doff = torch.zeros(total + 1, dtype=torch.int32, device=dev)
            doff[1:] = torch.cumsum(torch.bincount(d_img, minlength=total), 0)
            goff = torch.zeros(total + 1, dtype=torch.int32, device=dev)
            goff[1:] = torch.cumsum(torch.bincount(g_img, minlength=total), 0)
            one = CalculateMAP(thr)
            one.process_batch(all_dets[:, :6].contiguous(), doff, all_gts[:, :5].contiguous(), goff)
            want = one.fetch()
            same = bool(np.array_equal(want[0], result[0])) and bool(np.array_equal(want[1], result[1])) and want[2] == result[2]
            parity = "ok" if same else "FAILED"
            if not same:
                raise SystemExit("bench.py: config-5 gathered mAP differs from the single-GPU mAP")
        res = {"images": total, "images_per_rank": n_img, "detections": int(rows.size(0)), "match_ms": float(times[0]),
               "gather_ms": float(times[1]), "fetch_ms": float(times[2]), "images_per_s": total / (float(times.sum()) * 1e-3),
               "map50": float(result[0][0]), "map": float(result[0].mean()), "classes": len(result[2]),
               "parity_vs_single_gpu": parity}
    return res


# ------------------------------------------------------------------------------------------ CUDA leg
def run_cuda(args, cfg):
    rank, world, local = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback (use --impl reference)")
    from fastvision_b200 import _lib
    from fastvision_b200.pipeline import ValStep
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    affinity = bind_to_gpu_numa(local)      # before the pinned host buffers exist
    distributed = world > 1
    if distributed:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    batch = args.batch
    labels, heads = workload(cfg, batch, rank)
    host_heads = [h.pin_memory() for h in heads]
    host_labels = labels.pin_memory()
    dh = [h.to(dev, non_blocking=True) for h in host_heads]
    dl = host_labels.to(dev, non_blocking=True)
    torch.cuda.synchronize()

    step = ValStep(cfg.anchors_levels(), cfg.strides, batch_global=batch * world)
    lib = _lib.load()
    n0 = lib.fvb_launch_count()
    step(dh, dl)
    torch.cuda.synchronize()
    launches_per_step = int(lib.fvb_launch_count() - n0)
    decode_fn, tail_replay = step.capture(dh, dl, split_decode=True)

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- timed region: K steps, inputs resident in HBM; every decode launch bracketed by CUDA events ----
    for _ in range(max(args.warmup, 3)):
        decode_fn()
        tail_replay()
    barrier()
    k = args.steps
    # One CUDA graph holding the K steps back to back, each decode between its own pair of EXTERNAL timing events (event
    # record nodes): the K steps are timed with no host launch gaps, and every decode launch is still measured live inside
    # the timed region.  Not possible with the NCCL fallback (collectives stay outside graphs) or for very long runs.
    # (a pair of timing-event nodes costs ~10 us of the step it brackets -- plain graph replay: 0.374 ms, with events: 0.386 --
    # so only every `every`-th decode launch is bracketed; the samples are still taken live inside the timed region)
    every = max(1, args.decode_event_every)
    sampled = [i for i in range(k) if i % every == 0]
    unrolled = None
    if (not distributed or step._peer() is not None) and k <= 400 and not args.no_unrolled_graph:
        try:
            ev_d0 = {i: torch.cuda.Event(enable_timing=True, external=True) for i in sampled}
            ev_d1 = {i: torch.cuda.Event(enable_timing=True, external=True) for i in sampled}
            g_all = torch.cuda.CUDAGraph()
            cap_s = torch.cuda.Stream(device=dev)
            cap_s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.graph(g_all, stream=cap_s):
                for i in range(k):
                    step._head(dh, dl)                     # loss_match on the side stream, beside the decode
                    if i in ev_d0:
                        ev_d0[i].record()
                    step._decode(dh)
                    if i in ev_d1:
                        ev_d1[i].record()
                    step._tail(dh, dl, reduce_inside=distributed)
            torch.cuda.current_stream().wait_stream(cap_s)
            g_all.replay()                      # one untimed replay (also validates the graph)
            torch.cuda.synchronize()
            unrolled = g_all
        except Exception as exc:                # external events unsupported: fall back to per-step launches
            print("bench.py: unrolled graph unavailable (%s); timing per-step launches" % (exc,), file=sys.stderr)
            unrolled = None
    if unrolled is None:
        ev_d0 = {i: torch.cuda.Event(enable_timing=True) for i in sampled}
        ev_d1 = {i: torch.cuda.Event(enable_timing=True) for i in sampled}
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        barrier()
        t_begin.record()
        if unrolled is not None:
            unrolled.replay()
        else:
            for i in range(k):
                if i in ev_d0:
                    ev_d0[i].record()
                decode_fn()
                if i in ev_d1:
                    ev_d1[i].record()
                tail_replay()
        t_end.record()
        barrier()
    total_ms = t_begin.elapsed_time(t_end)
    step.check()                                   # no NMS CTA timed out, no failed peer reduction during the timed region
    decode_ms = [ev_d0[i].elapsed_time(ev_d1[i]) for i in sampled]
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if distributed:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    value = batch * world * k / (total_ms_max * 1e-3)
    # per-rank view (N > 1): every rank's own decode time and timed-region length.  The step ends with a reduce that aligns the
    # ranks, so the slowest GPU's decode sets everybody's step; these numbers say whether that or the reduce is the difference to N=1
    per_rank = None
    if distributed:
        mine = torch.tensor([sum(decode_ms) / len(decode_ms), total_ms / k], dtype=torch.float64, device=dev)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = {"decode_ms": [round(float(x[0]), 4) for x in allr], "ms_per_step": [round(float(x[1]), 4) for x in allr]}

    # ---- e2e: host buffers -> public API -> host results, copies inside the timed region --------------------
    out = step.out
    h_loss = torch.empty(1, dtype=torch.float32).pin_memory()
    h_boxes = torch.empty_like(out["boxes"], device="cpu").pin_memory()
    h_scores = torch.empty_like(out["scores"], device="cpu").pin_memory()
    h_cls = torch.empty_like(out["cls"], device="cpu").pin_memory()
    h_cnt = torch.empty_like(out["cnt"], device="cpu").pin_memory()
    h2d = sum(h.numel() * 4 for h in host_heads) + host_labels.numel() * 4
    d2h = 4 + h_boxes.numel() * 4 + h_scores.numel() * 4 + h_cls.numel() * 8 + h_cnt.numel() * 4

    def e2e_step():
        for dst, src in zip(dh, host_heads):
            dst.copy_(src, non_blocking=True)
        dl.copy_(host_labels, non_blocking=True)
        o = step(dh, dl)                      # the public API call (eager, 5 kernel launches)
        h_loss.copy_(o["loss"], non_blocking=True)
        h_boxes.copy_(o["boxes"], non_blocking=True)
        h_scores.copy_(o["scores"], non_blocking=True)
        h_cls.copy_(o["cls"], non_blocking=True)
        h_cnt.copy_(o["cnt"], non_blocking=True)
        torch.cuda.current_stream().synchronize()   # the caller reads the loss / detections every step
        return float(h_loss[0])

    ke = max(3, min(args.e2e_steps, k))
    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(ke):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if distributed:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_serial = batch * world * ke / float(te.item())

    # The same loop double-buffered: the H2D copy of step i+1 (copy stream, second device buffer set) runs under the kernels
    # of step i; the host reads step i's loss / detections (copied D2H every step) once step i+1 has been enqueued.  Every
    # byte still crosses PCIe inside the timed region; only the idle bubbles between copy and compute go away.
    copy_s = torch.cuda.Stream(device=dev)
    sets = [(dh, dl, step)]
    dh2 = [torch.empty_like(h) for h in dh]
    dl2 = torch.empty_like(dl)
    step2 = ValStep(cfg.anchors_levels(), cfg.strides, batch_global=batch * world)
    step2(dh2, dl2[:0])
    sets.append((dh2, dl2, step2))
    hosts = []
    for _ in range(2):
        hosts.append({"loss": torch.empty(1, dtype=torch.float32).pin_memory(), "boxes": torch.empty_like(h_boxes).pin_memory(),
                      "scores": torch.empty_like(h_scores).pin_memory(), "cls": torch.empty_like(h_cls).pin_memory(),
                      "cnt": torch.empty_like(h_cnt).pin_memory()})
    ev_copied = [torch.cuda.Event() for _ in range(2)]
    ev_done = [torch.cuda.Event() for _ in range(2)]
    main_s = torch.cuda.current_stream()

    def e2e_pipelined(n_steps):
        last = None
        for i in range(n_steps):
            sl = i & 1
            bh, bl, st = sets[sl]
            with torch.cuda.stream(copy_s):
                if i >= 2:
                    copy_s.wait_event(ev_done[sl])          # the set's previous step has consumed its inputs
                for dst, src in zip(bh, host_heads):
                    dst.copy_(src, non_blocking=True)
                bl.copy_(host_labels, non_blocking=True)
                ev_copied[sl].record(copy_s)
            main_s.wait_event(ev_copied[sl])
            o = st(bh, bl)
            hb = hosts[sl]
            for key in ("loss", "boxes", "scores", "cls", "cnt"):
                hb[key].copy_(o[key], non_blocking=True)
            ev_done[sl].record(main_s)
            if last is not None:                            # read the previous step's results while this one runs
                ev_done[last].synchronize()
                _ = float(hosts[last]["loss"][0])
            last = sl
        ev_done[last].synchronize()
        return float(hosts[last]["loss"][0])

    e2e_pipelined(4)
    barrier()
    t0 = time.perf_counter()
    e2e_pipelined(ke)
    barrier()
    tp2 = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if distributed:
        dist.all_reduce(tp2, op=dist.ReduceOp.MAX)
    e2e_value = batch * world * ke / float(tp2.item())
    e2e_seconds = float(tp2.item())
    del dh2, dl2, step2, sets
    # the host-side ceiling of that loop: the same pinned -> device copies with NO kernels, all ranks at once
    barrier()
    t0 = time.perf_counter()
    for _ in range(ke):
        for dst, src in zip(dh, host_heads):
            dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    barrier()
    tc = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if distributed:
        dist.all_reduce(tc, op=dist.ReduceOp.MAX)
    h2d_copy_only_gbps = h2d * ke / float(tc.item()) / 1e9

    # ---- extra (reported in config, not the headline): the cross-batch pipeline, tail of batch i under decode i+1 ----
    from fastvision_b200.pipeline import ValPipeline
    pipe = ValPipeline(cfg.anchors_levels(), cfg.strides, batch_global=batch * world)
    for _ in range(4):
        pipe.submit(dh, dl)
    pipe.flush()
    barrier()
    kp = max(20, min(k, 100))
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for _ in range(kp):
        pipe.submit(dh, dl)
    pipe.flush()
    p1.record()
    barrier()
    tp = torch.tensor([p0.elapsed_time(p1)], dtype=torch.float64, device=dev)
    if distributed:
        dist.all_reduce(tp, op=dist.ReduceOp.MAX)
    pipelined_ms = float(tp.item()) / kp
    del pipe

    # ---- the decode kernel by itself (outside the timed region): inside the step it shares every SM with an NMS CTA, so the live
    # figure above is the kernel UNDER that load; this is the same launch (all side outputs) with nothing beside it
    from fastvision_b200.detection.models import yolov3_decode
    def decode_alone(n):
        for _ in range(n):
            yolov3_decode(dh, cfg.anchors_levels(), cfg.strides, ctx=step.ctx, out=step.out["results"], conf_thres=step.conf_thres,
                          want_bce0=True)
    decode_alone(3)
    torch.cuda.synchronize()
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record()
    decode_alone(20)                              # back to back (candidate bits are OR-ed: re-setting them changes nothing)
    a1.record()
    torch.cuda.synchronize()
    step.ctx.bitmap().zero_()                     # (the NMS kernel would have consumed and cleared the candidate bits)
    decode_alone_ms = a0.elapsed_time(a1) / 20
    # ---- correctness and the other BASELINE configs, outside the timed region ---------------------------------------------
    extras = {}
    if not args.no_extras:
        extras["config5_map"] = config5_map(step, labels, cfg, batch, rank, world, dev, distributed, barrier)
        if distributed:
            extras["dp_parity"] = dp_parity(step, dh, dl, batch, rank, world, dev, dist, cfg)
            extras["strong_config3"] = strong_config3(rank, world, dev, dist, barrier)
        else:
            extras.update(other_configs_1gpu(dev, measured_peak_hbm()[0]))
    if unrolled is not None:
        step_launch = ("the K steps captured back to back in one CUDA graph (decode -> NMS branch || loss branch%s), every decode "
                       "kernel between its own pair of external timing-event nodes" %
                       (" + single-kernel all-reduce/combine of the 12 fp64 partials over NVLink peer memory" if distributed else ""))
    elif not distributed:
        step_launch = "decode launched eagerly between CUDA events, NMS + loss branches replayed as a CUDA graph"
    elif step._peer() is not None:
        step_launch = ("decode launched eagerly between CUDA events; NMS branch and loss branch (+ the single-kernel all-reduce + "
                       "combine of the 12 fp64 partials over NVLink peer memory) replayed as a CUDA graph")
    else:
        step_launch = ("decode launched eagerly between CUDA events; NMS branch and loss branch (+ NCCL all-reduce of the 12 fp64 "
                       "partials) launched eagerly on two streams")
    if rank == 0:
        peak, peak_src = measured_peak_hbm()
        rows = step.ctx.rows
        alg_bytes = 2 * batch * rows * step.ctx.k * 4          # one read of the heads + one write of results
        dec_avg = sum(decode_ms) / len(decode_ms)
        achieved = alg_bytes / (dec_avg * 1e-3) / 1e9
        traffic, traffic_src = measured_traffic()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": k, "warmup": max(args.warmup, 3),
            "ms_per_step": total_ms_max / k, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "batch_per_gpu": batch, "global_batch": batch * world, "rows_per_image": rows, "channels": step.ctx.k,
                       "conf_thres": 0.25, "iou_thres": 0.45, "max_det": 300, "loss_ratios": [0.05, 1.0, 0.5],
                       "labels": int(labels.size(0)), "parallelism": "per-image sharding, dp%d" % world,
                       "l2": "inputs (%.0f MB per step) larger than the 126 MB L2; no flush needed" % (alg_bytes / 2e6),
                       "step_launch": step_launch,
                       "pipelined_extra": {"ms_per_step": pipelined_ms, "images_per_s": batch * world / (pipelined_ms * 1e-3), "steps": kp,
                                           "note": "ValPipeline: two ValSteps on alternating streams -- batch i+1's decode waits only for batch i's "
                                                   "decode kernel, so the ~40 us left of batch i (NMS of its last images, loss finish) runs under it; "
                                                   "not the headline: the K steps of the headline do not overlap one another"}},
            "clocks": clocks.summary(),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": ke, "serial_value": e2e_serial,
                    "h2d_gbps_per_rank": h2d * ke / e2e_seconds / 1e9, "h2d_copy_only_gbps_per_rank": h2d_copy_only_gbps,
                    "cpu_affinity_cpus": affinity,
                    "note": "pinned host heads+labels copied H2D, ValStep public call, loss + padded detections copied D2H and read by the "
                            "host, every step; double-buffered (the copy of step i+1 overlaps the kernels of step i); serial_value = "
                            "the same loop with no overlap"},
            "gpu_launches": launches_per_step * k,
            "roofline": {"bound": "hbm", "kernel": "fvb::decode_kernel (decode + candidate bitmap/records + objectness-BCE partials)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "traffic_source": traffic_src,
                         "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": dec_avg,
                         "launches_timed": "%d of the %d decode launches of the timed region (every %d-th), CUDA events" % (len(sampled), k, every), "peak_source": peak_src,
                         "note": "launch_ms / achieved / frac are measured live inside the timed region, where every image's NMS runs UNDER "
                                 "this kernel (programmatic dependent launch) and takes issue slots from it; kernel_alone is the same "
                                 "launch with nothing beside it (20 back-to-back launches between one pair of events, after the timed region)",
                         "kernel_alone": {"launch_ms": decode_alone_ms, "achieved": alg_bytes / (decode_alone_ms * 1e-3) / 1e9,
                                          "frac": alg_bytes / (decode_alone_ms * 1e-3) / 1e9 / peak},
                         "step_achieved": alg_bytes * world / (total_ms_max / k * 1e-3) / 1e9 / world,
                         "step_frac": alg_bytes / (total_ms_max / k * 1e-3) / 1e9 / peak},
        }
        if per_rank is not None:
            line["config"]["per_rank"] = per_rank
        for key, val in extras.items():
            if val is not None:
                line["config"][key] = val
        if "dp_parity" in extras and extras["dp_parity"] is not None:
            line["dp_parity"] = extras["dp_parity"]["status"]
        if not args.no_cpu_baseline and world == 1:
            torch.set_num_threads(os.cpu_count() or 1)
            v, med, reps, backend = time_cpu_reference(cfg, args.cpu_sample, 0, args.cpu_budget, 3, 1)
            line["cpu_baseline"] = {
                "value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                "sample": "%d-image slice of the same workload (same generator/seed), %d passes, median %.3f s/pass; oracle port "
                          "(torch-CPU restatement of the reference, NMS backend '%s')" % (args.cpu_sample, reps, med, backend),
                "cpu_model": cpu_model_name(), "stages": cpu_stage_baseline()}
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line), flush=True)
    if distributed:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="images per GPU (BASELINE configs[1]: 256)")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--cpu-sample", type=int, default=16, help="images in the bounded CPU-baseline sample")
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of CPU-baseline timing")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--decode-event-every", type=int, default=16, help="bracket every n-th decode launch with timing events")
    ap.add_argument("--no-unrolled-graph", action="store_true", help="time per-step launches instead of one K-step CUDA graph")
    ap.add_argument("--no-extras", action="store_true", help="skip config-5 mAP, dp_parity and config-3 strong scaling")
    args = ap.parse_args()
    cfg = synth.COCO416
    if args.impl == "reference":
        run_reference(args, cfg)
    else:
        run_cuda(args, cfg)


if __name__ == "__main__":
    main()
