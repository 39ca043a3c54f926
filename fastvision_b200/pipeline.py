"""The fused validation step: decode -> NMS (all images) -> loss, i.e. the per-batch body of the
reference's ``Fit._val`` (utils/fit.py:86-95) without its per-image Python loop and host syncs.

Kernel sequence per step (5 launches, optionally replayed as one CUDA graph):
    side stream   loss_prep, loss_match (raw heads + labels only: beside the decode)             fvb_yolov3_loss_match_dense_f32
    main stream   decode (+ NMS candidate bitmap/records + zero-target objectness BCE partials,
                  publishing per-image progress)                                                 fvb_yolo_decode_sync_f32
    main stream   NMS for every image, a PROGRAMMATIC DEPENDENT of the decode kernel: image b's
                  CTA starts when image b is decoded, while later images still stream            fvb_yolo_nms_after_decode_f32
    side stream   after the decode: objectness-partials sum + reduce (+ data-parallel combine)    fvb_yolov3_loss_finish_f32
Under ``torch.distributed`` the batch is sharded per image: every rank runs the same step on its slice
and the only collective is an all-reduce of the L*4 fp64 loss partials; decode and NMS need none.
"""
from typing import List, Optional

import torch

from . import _lib
from .detection.models.yolov3 import DecodeContext, yolov3_decode
from .detection.tools.nms import non_max_suppression_batched
from .loss.yolov3_loss import Yolov3Loss


class _ModelStub:
    def __init__(self, anchors_per_level, strides):
        self.anchors_per_level = anchors_per_level
        self.backbone_strides_per_level = strides


class ValStep:
    def __init__(self, anchors_per_level, strides, conf_thres=0.25, iou_thres=0.45, max_det=300,
                 ratio_box=0.05, ratio_conf=1.0, ratio_cls=0.5, nms_flavour="lib", precise_decode=False,
                 process_group=None, batch_global: Optional[int] = None, overlap_nms: Optional[bool] = None,
                 data_parallel: Optional[bool] = None):
        self.anchors_per_level = anchors_per_level
        self.strides = strides
        self.conf_thres, self.iou_thres, self.max_det = conf_thres, iou_thres, max_det
        self.nms_flavour = nms_flavour
        self.precise = precise_decode
        self.loss_fn = Yolov3Loss(_ModelStub(anchors_per_level, strides), 0.5, ratio_box, ratio_conf, ratio_cls)
        self.pg = process_group
        self.batch_global = batch_global
        self.data_parallel = data_parallel     # None: follow torch.distributed; False: a local step even inside a process group
        self.ctx = None
        self.graph = None
        self.out = None
        self._side = None
        self._ev_decoded = None
        self._ev_loss = None
        self._peer_reducer = False       # not looked up yet
        # image b's NMS starts as soon as image b is decoded: the NMS kernel is launched as a programmatic dependent of the
        # decode kernel and follows its per-image progress counters (fvb_yolo_decode_sync_f32 / fvb_yolo_nms_after_decode_f32)
        # (None: decided per head geometry in _prepare -- on when an NMS CTA fits on an SM beside a decode CTA)
        self._overlap_request = overlap_nms
        self.overlap_nms = bool(overlap_nms)
        self._ws = _lib.Workspaces()     # owned by this step: a captured graph has these pointers baked in

    def _prepare(self, heads):
        ctx = DecodeContext(heads, self.anchors_per_level, self.strides)
        if self.ctx is not None and self.ctx.key == ctx.key:
            return
        self.ctx = ctx
        dev, b, md = ctx.device, ctx.batch, self.max_det
        self.out = {
            "results": torch.empty(b, ctx.rows, ctx.k, dtype=torch.float32, device=dev),
            "boxes": torch.empty(b, md, 4, dtype=torch.float32, device=dev),
            "scores": torch.empty(b, md, dtype=torch.float32, device=dev),
            "cls": torch.empty(b, md, dtype=torch.int64, device=dev),
            "cnt": torch.zeros(b, dtype=torch.int32, device=dev),
            "rows": torch.empty(b, md, dtype=torch.int32, device=dev),
            "loss": torch.empty(1, dtype=torch.float32, device=dev),
            "partials": torch.empty(ctx.geom.levels, 4, dtype=torch.float64, device=dev),
        }
        ctx.bitmap()
        ctx.records()
        ctx.bce0()
        ctx.tile_sync()
        room = _lib.load().fvb_yolo_decode_leaves_room_for_nms(ctx.geom) == 1
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        # no room (narrow rows): the pair still pays for at most one image per SM -- 1024-thread NMS CTAs that start, SM by SM, as
        # the decode CTAs leave instead of after the decode grid + a launch gap (608 / C=10 / 128 images: 0.1083 -> 0.1040 ms).
        # Not under data parallelism: whole-SM NMS CTAs that arrive that early leave the loss finish + peer reduce no SM to run
        # on until they are done (2 ranks x 128 images: 0.1115 vs 0.1035-0.1070 ms with the plain launch).
        self._wide_cta = not room
        if self._overlap_request is None:
            self.overlap_nms = room or (b <= sms and not self._distributed())
        self._room = room
        self._nms_ws = self._ws.get("yolo_nms", _lib.load().fvb_yolo_nms_workspace_bytes(b, ctx.rows), dev)
        self.graph = None
        # NMS (latency-bound, one CTA per image) and the loss kernels only depend on the decode: they run as
        # two branches (second stream; two parallel branches of the graph when captured)
        self._side = torch.cuda.Stream(device=dev)
        self._ev_decoded = torch.cuda.Event()
        self._ev_loss = torch.cuda.Event()
        self._ev_start = torch.cuda.Event()

    def _distributed(self):
        if self.data_parallel is not None:
            return bool(self.data_parallel)
        return self.pg is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()
                                       and torch.distributed.get_world_size() > 1)

    def _early_match(self):
        """The loss in two parts -- target assignment + matched-row terms beside the decode, the rest after it -- pays when the
        loss branch would otherwise be the step's tail or fight the NMS for the SMs: with the overlapped NMS (an NMS CTA beside
        every decode CTA leaves no registers for loss_match's CTAs) and under data parallelism (short NMS, reduce at the end).
        A single GPU with a geometry that cannot overlap (narrow rows) keeps the one-call loss after the decode, where the long
        NMS phase hides it completely; launched in front of the decode its CTAs only delay the decode's start (608 / C=10 /
        B=1024: 0.742 vs 0.72 ms)."""
        return (self.overlap_nms and self._room) or self._distributed()

    def _head(self, heads, labels):
        """First launches of a step: the loss's target assignment + matched-row terms read only the RAW heads and the labels,
        so they go onto the side stream BEFORE the decode is launched -- their few small CTAs take their SM slots first and
        hide under the decode kernel (launched after the decode + overlapped NMS pair they would wait for free registers)."""
        if not self._early_match():
            return
        self._ev_start.record(torch.cuda.current_stream())
        with torch.cuda.stream(self._side):
            self._side.wait_event(self._ev_start)
            self.loss_fn.match(heads, labels, self.ctx, conf_bce0_precise=self.precise)

    def _decode(self, heads):
        ctx, o = self.ctx, self.out
        yolov3_decode(heads, self.anchors_per_level, self.strides, precise=self.precise, ctx=ctx, out=o["results"],
                      conf_thres=self.conf_thres, want_bce0=True, tile_sync=ctx.tile_sync() if self.overlap_nms else None)

    def _nms(self):
        """NMS of every image (main stream).  With ``overlap_nms`` this launch must directly follow ``_decode``'s."""
        ctx, o = self.ctx, self.out
        sync = ctx.tile_sync() if self.overlap_nms else None
        non_max_suppression_batched(o["results"], self.conf_thres, self.iou_thres, self.max_det, self.nms_flavour,
                                    cand_bitmap=ctx.bitmap(), cand_records=ctx.records(), clear_bitmap=True,
                                    out=(o["boxes"], o["scores"], o["cls"], o["cnt"], o["rows"]),
                                    tile_sync=sync, tiles_per_image=ctx.tiles_per_image, ws=self._nms_ws,
                                    wide_cta=self._wide_cta and self.overlap_nms)

    def _tail(self, heads, labels, reduce_inside=False):
        """Everything after the decode launch (``_head`` and ``_decode`` came first): the NMS kernel directly behind the decode
        kernel on the main stream (its programmatic dependent: image b's NMS starts when image b is decoded), and on the side
        stream the second half of the two-part loss -- the sum of the decode's objectness partials + the reduction of the
        matched terms -- followed, with ``reduce_inside``, by the data-parallel all-reduce + combine.  Joined at the end."""
        ctx, o = self.ctx, self.out
        main = torch.cuda.current_stream()
        self._ev_decoded.record(main)       # (an event record is not a launch: the NMS kernel still follows the decode kernel)
        self._nms()
        with torch.cuda.stream(self._side):
            self._side.wait_event(self._ev_decoded)
            if self._early_match():
                self.loss_fn.finish(labels.size(0), ctx, ctx.bce0(), out=o["loss"], partials=o["partials"])
            else:
                self.loss_fn(heads, labels, conf_bce0=ctx.bce0(), ctx=ctx, out=o["loss"], partials=o["partials"],
                             conf_bce0_precise=self.precise)
            if reduce_inside:
                self._reduce()
            self._ev_loss.record(self._side)
        main.wait_event(self._ev_loss)

    def _reduce(self):
        """Data-parallel only: all-reduce the per-level partial sums (96 bytes, the ONLY collective of the step) and
        form the scalar with the global-batch normalisers.  Kept outside the captured graph (plain NCCL launch)."""
        if not self._distributed():
            return
        o, ctx = self.out, self.ctx
        bg = self.batch_global or ctx.batch * torch.distributed.get_world_size(self.pg)
        peer = self._peer()
        if peer is not None:
            # one single-CTA kernel over NVLink peer memory: publish, wait for the peers, sum in rank order, combine
            peer.reduce_combine(self.loss_fn, ctx.geom, o["partials"], bg, o["loss"])
            return
        torch.distributed.all_reduce(o["partials"], group=self.pg)
        self.loss_fn.combine(o["partials"], bg, ctx=ctx, out=o["loss"])

    def _peer(self):
        if self._peer_reducer is False:
            from .dist import peer_reducer
            self._peer_reducer = peer_reducer(self.ctx.device, self.pg) if torch.distributed.get_backend(self.pg) == "nccl" else None
        return self._peer_reducer

    def _run(self, heads, labels):
        self._head(heads, labels)
        self._decode(heads)
        self._tail(heads, labels, reduce_inside=True)
        return self.out

    def __call__(self, head_out: List[torch.Tensor], labels: torch.Tensor):
        """head_out: raw [B,A,H,W,K] per level (this rank's images); labels [T,6] with LOCAL batch indices.

        Returns the dict of persistent output buffers: results [B,N,K], padded detections
        (boxes/scores/cls/rows + cnt) and loss [1].  Buffers are reused by the next call.
        """
        heads = [_lib.require_cuda(h, "head_out[%d]" % i) for i, h in enumerate(head_out)]
        labels = _lib.require_cuda(labels, "labels").view(-1, 6)
        self._prepare(heads)
        return self._run(heads, labels)

    def capture(self, head_out, labels, split_decode=False):
        """Capture the step for these (static) input tensors into a CUDA graph; returns a replay callable.

        With ``split_decode`` only the part after the decode is captured and ``(decode_fn, tail_replay)`` is
        returned, so a caller can bracket the (eagerly launched) decode kernel with timing events (the NMS kernel is then
        not a programmatic dependent of the decode: it finds every image complete and runs as a plain launch would).
        """
        heads = [_lib.require_cuda(h, "head_out[%d]" % i) for i, h in enumerate(head_out)]
        labels = _lib.require_cuda(labels, "labels").view(-1, 6)
        self._prepare(heads)
        dist_graph = False
        if self._distributed():
            if self._peer() is None:
                # NCCL fallback: keep the all-reduce out of any graph and inside the loss branch (overlaps the NMS)
                if split_decode:
                    return (lambda: (self._head(heads, labels), self._decode(heads))), (lambda: self._tail(heads, labels, reduce_inside=True))
                return lambda: self._run(heads, labels)
            dist_graph = True    # the peer-memory reduce is an ordinary kernel: the whole step can be captured
        warm = torch.cuda.Stream(device=self.ctx.device)
        warm.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(warm):
            for _ in range(2):          # warm-up: sizes every cached workspace outside the capture
                self._run(heads, labels)
        torch.cuda.current_stream().wait_stream(warm)
        torch.cuda.synchronize(self.ctx.device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            if not split_decode:
                self._head(heads, labels)
                self._decode(heads)
            self._tail(heads, labels, reduce_inside=dist_graph)
        self.graph = g

        if split_decode:
            return (lambda: (self._head(heads, labels), self._decode(heads))), g.replay
        return g.replay

    def check(self):
        """Host-side health check for the places where the caller reads results anyway (it syncs): raises if an image's NMS
        gave up waiting for its decode (cnt < 0) or if a data-parallel peer reduction failed (loss is NaN then)."""
        if self.out is not None and int(self.out["cnt"].min()) < 0:
            raise RuntimeError("fastvision_b200: an overlapped NMS CTA timed out waiting for the decode kernel (cnt = -1)")
        if self._distributed() and self._peer_reducer not in (False, None):
            self._peer_reducer.raise_if_failed()

    def detections(self, out=None):
        """Ragged per-image detections [k,6] = [cls, conf, x1, y1, x2, y2] (the layout utils/fit.py:96 builds). Syncs."""
        o = out or self.out
        cnt = o["cnt"].cpu().tolist()
        dets = torch.cat([o["cls"].float().unsqueeze(-1), o["scores"].unsqueeze(-1), o["boxes"]], dim=2)
        return [dets[i, :c] for i, c in enumerate(cnt)]


class ValPipeline:
    """Software-pipelined validation loop: what is left of batch i after its decode overlaps the decode of batch i+1.

    ``Fit._val`` (utils/fit.py:86-105) walks the validation set batch by batch; nothing in batch i+1's decode depends on batch
    i's detections.  Every ``ValStep`` already runs each image's NMS under its own decode kernel; what remains at the end of a
    step is the NMS of the images decoded last (~40 us) and the loss finish.  Here ``depth`` ValSteps (own buffers, own
    streams) alternate: batch i+1's decode waits only for batch i's DECODE kernel (two persistent decode kernels would just
    fight for the SMs), not for its tail, which therefore runs under the next decode.  ``submit`` returns the slot's output
    buffers; they are valid after ``wait(slot)`` / ``flush()`` and stay so until the slot is submitted again.
    """

    def __init__(self, anchors_per_level, strides, depth=2, **kw):
        self.steps = [ValStep(anchors_per_level, strides, **kw) for _ in range(depth)]
        self.depth = depth
        self.count = 0
        self._streams = None
        self._done = None
        self._ev_in = None

    def _prepare(self, dev):
        if self._streams is None:
            self._streams = [torch.cuda.Stream(device=dev) for _ in range(self.depth)]
            self._done = [torch.cuda.Event() for _ in range(self.depth)]
            self._ev_in = [torch.cuda.Event() for _ in range(self.depth)]

    def submit(self, head_out: List[torch.Tensor], labels: torch.Tensor, decode_events=None):
        """Enqueue one batch; returns the dict of output buffers of this slot (valid after ``wait(slot)``/``flush``).

        ``decode_events``: optional (start, stop) CUDA events recorded around the decode kernel on its stream.
        """
        slot = self.count % self.depth
        st = self.steps[slot]
        heads = [_lib.require_cuda(h, "head_out[%d]" % i) for i, h in enumerate(head_out)]
        labels = _lib.require_cuda(labels, "labels").view(-1, 6)
        st._prepare(heads)
        self._prepare(st.ctx.device)
        stream, ev_in = self._streams[slot], self._ev_in[slot]
        ev_in.record(torch.cuda.current_stream())    # inputs are ready once the caller's stream gets here
        with torch.cuda.stream(stream):
            stream.wait_event(ev_in)
            # (the slot's previous batch is older work on this same stream: its buffers are free by stream order)
            if self.count > 0:
                stream.wait_event(self.steps[(self.count - 1) % self.depth]._ev_decoded)   # the previous batch's decode kernel
            st._head(heads, labels)
            if decode_events is not None:
                decode_events[0].record(stream)
            st._decode(heads)
            if decode_events is not None:
                decode_events[1].record(stream)
            st._tail(heads, labels, reduce_inside=True)
            self._done[slot].record(stream)
        self.count += 1
        return st.out

    def wait(self, slot=None):
        """Make the current stream wait for the batch in ``slot`` (default: the most recently submitted one)."""
        if self.count == 0:
            return
        slot = (self.count - 1) % self.depth if slot is None else slot
        torch.cuda.current_stream().wait_event(self._done[slot])

    def flush(self):
        """Join every outstanding batch into the current stream."""
        for slot in range(min(self.count, self.depth)):
            self.wait(slot)


from .dist import shard_labels  # noqa: E402,F401  (re-exported: tests and callers import it from here)
