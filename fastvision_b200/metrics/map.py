"""COCO-style mAP -- drop-in for CalculateMAP, metrics/map.py:6-141.

``process_one`` keeps the reference signature; ``process_batch`` matches every image of a batch in ONE
kernel launch (fvb_map_match_f32) without the per-image D2H + numpy argsort/unique, and keeps the evidence on
the device.  ``fetch`` runs the AP integration of map.py:85-141 on the device too (fvb_map_ap_f64: radix sort by
(class, -conf), TP prefix sums, precision envelope, 101-point np.interp + np.trapz in float64) and only brings
the [classes, thresholds] AP table back for the two final means.
"""
import ctypes as C

import numpy as np
import torch

from .. import _lib


class CalculateMAP:
    def __init__(self, map_iou_values):
        self.map_iou_values = np.asarray(map_iou_values, dtype=np.float64)
        if self.map_iou_values.size > 16:
            raise ValueError("at most 16 IoU thresholds")
        self._dets_list = []     # device [M,6] blocks = [cls, conf, x1,y1,x2,y2]
        self._correct_list = []  # device u8 [M, n_thr] blocks
        self._targets_list = []  # device f32 target classes
        self._pending = []       # capacity-shaped batches of process_padded: (dets, correct, det_off, gts, gt_off)

    # Batches recorded by ``process_padded`` keep capacity shapes (no host sync while the loader runs); their valid row counts
    # are read in ONE transfer the first time the evidence is needed (``fetch`` / ``state`` / the reference's accumulators).
    def _flush(self):
        if not self._pending:
            return
        totals = torch.stack([torch.stack([d_off[-1], g_off[-1]]) for _, _, d_off, _, g_off in self._pending]).cpu().tolist()
        for (dets, correct, _, gts, _), (nd, ng) in zip(self._pending, totals):
            if ng:
                self._targets_list.append(gts[:ng, 0].contiguous())
            if nd:
                self._dets_list.append(dets[:nd])
                self._correct_list.append(correct[:nd])
        self._pending = []

    @property
    def _dets(self):
        self._flush()
        return self._dets_list

    @_dets.setter
    def _dets(self, v):
        self._pending, self._dets_list = [], v

    @property
    def _correct(self):
        self._flush()
        return self._correct_list

    @_correct.setter
    def _correct(self, v):
        self._correct_list = v

    @property
    def _targets(self):
        self._flush()
        return self._targets_list

    @_targets.setter
    def _targets(self, v):
        self._targets_list = v

    # the reference's public accumulators (map.py:13-14), materialised on the host on demand
    @property
    def correct_all_images(self):
        """list of float64 [M, 2 + n_thr] blocks [conf, cls, correct...] (one per process_one / process_batch call)."""
        out = []
        for d, c in zip(self._dets, self._correct):
            block = np.zeros([d.size(0), 2 + self.map_iou_values.size], dtype=np.float64)
            host = d[:, :2].detach().cpu().numpy()
            block[:, 0] = host[:, 1]
            block[:, 1] = host[:, 0]
            block[:, 2:] = c.cpu().numpy()
            out.append(block)
        return out

    @property
    def seen_all_targets_cls(self):
        return [t.detach().cpu().numpy() for t in self._targets]

    # ---- device matcher -------------------------------------------------------------------------------
    def match(self, dets, det_off, gts, gt_off):
        """dets [sum M,6]=[cls,conf,x1,y1,x2,y2], gts [sum N,5]=[cls,x1,y1,x2,y2], int32 CSR offsets [I+1] -> u8 [sum M,n_thr]."""
        dets = _lib.require_cuda(dets, "dets")
        gts = _lib.require_cuda(gts, "gts")
        det_off = _lib.require_cuda(det_off, "det_off", torch.int32)
        gt_off = _lib.require_cuda(gt_off, "gt_off", torch.int32)
        images = det_off.numel() - 1
        total = dets.size(0)
        nthr = self.map_iou_values.size
        correct = torch.zeros(total, nthr, dtype=torch.uint8, device=dets.device)
        lib = _lib.load()
        ws = _lib.workspace(lib.fvb_map_match_workspace_bytes(total), dets.device, "map_match")
        thr = (C.c_double * nthr)(*self.map_iou_values.tolist())
        with torch.cuda.device(dets.device):
            _lib.check(lib.fvb_map_match_f32(_lib.dptr(dets), _lib.dptr(det_off), _lib.dptr(gts), _lib.dptr(gt_off),
                                             images, total, thr, nthr, _lib.dptr(correct), _lib.dptr(ws),
                                             _lib.stream()), "map_match")
        return correct

    def process_batch(self, dets, det_off, gts, gt_off):
        """Batched ``process_one``: records the detections, their ``correct`` bits and the target classes (all on the device)."""
        correct = self.match(dets, det_off, gts, gt_off)
        if gts.size(0):                                   # map.py:43-44
            self._targets.append(gts[:, 0].detach().contiguous())
        if dets.size(0) == 0:                             # map.py:46-47
            return
        self._dets.append(dets.detach())
        self._correct.append(correct)

    def process_padded(self, boxes, scores, cls, cnt, labels, img_w, img_h):
        """The whole per-image loop of utils/fit.py:94-101 for one batch, on the device and WITHOUT a host sync: the padded NMS
        outputs (boxes [B,max_det,4] xyxy, scores, cls int64, cnt int32) and the batch's labels [T,6] become compact detection
        rows + pixel-unit targets (``fvb_val_evidence_f32``), then one matcher launch.  Row counts stay on the device until
        ``fetch``."""
        boxes = _lib.require_cuda(boxes, "boxes")
        scores = _lib.require_cuda(scores, "scores")
        cls = _lib.require_cuda(cls, "cls", torch.int64)
        cnt = _lib.require_cuda(cnt, "cnt", torch.int32)
        labels = _lib.require_cuda(labels, "labels").view(-1, 6)
        b, md = boxes.size(0), boxes.size(1)
        t, dev = labels.size(0), boxes.device
        dets = torch.empty(b * md, 6, dtype=torch.float32, device=dev)
        gts = torch.empty(max(t, 1), 5, dtype=torch.float32, device=dev)
        det_off = torch.empty(b + 1, dtype=torch.int32, device=dev)
        gt_off = torch.empty(b + 1, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.load().fvb_val_evidence_f32(_lib.dptr(boxes), _lib.dptr(scores), _lib.dptr(cls), _lib.dptr(cnt), b, md,
                                                        _lib.dptr(labels), t, float(img_w), float(img_h), _lib.dptr(dets),
                                                        _lib.dptr(det_off), _lib.dptr(gts), _lib.dptr(gt_off), _lib.stream()),
                       "val_evidence")
        correct = self.match(dets, det_off, gts, gt_off)
        self._pending.append((dets, correct, det_off, gts, gt_off))

    def process_one(self, y_pred, y_true):
        """metrics/map.py:16-83.  y_pred[M,6]=[cls,conf,x1,y1,x2,y2]; y_true[N,5]=[cls,x1,y1,x2,y2]."""
        y_pred = _lib.require_cuda(y_pred, "y_pred").view(-1, 6)
        y_true = _lib.require_cuda(y_true, "y_true").view(-1, 5)
        dev = y_pred.device
        det_off = torch.tensor([0, y_pred.size(0)], dtype=torch.int32, device=dev)
        gt_off = torch.tensor([0, y_true.size(0)], dtype=torch.int32, device=dev)
        self.process_batch(y_pred, det_off, y_true, gt_off)

    # ---- evidence exchange for data-parallel evaluation (dist.gather_map_state) -----------------------------------------
    def state(self):
        """-> (rows [sum M, 2 + n_thr] f32 = [conf, cls, correct...], target classes [sum N] f32), both on the device."""
        nthr = self.map_iou_values.size
        if self._dets:
            d = torch.cat(self._dets)
            rows = torch.cat([d[:, 1:2], d[:, 0:1], torch.cat(self._correct).float()], dim=1)
        else:
            rows = torch.zeros(0, 2 + nthr, device=self._targets[0].device if self._targets else "cuda")
        tg = torch.cat(self._targets) if self._targets else torch.zeros(0, device=rows.device)
        return rows, tg

    def load_state(self, rows, target_cls):
        """Replace the accumulated evidence by gathered rows / target classes (see ``state``)."""
        rows = _lib.require_cuda(rows, "rows")
        dets = torch.zeros(rows.size(0), 6, device=rows.device)
        dets[:, 0], dets[:, 1] = rows[:, 1], rows[:, 0]
        self._dets, self._correct = [dets], [rows[:, 2:].to(torch.uint8).contiguous()]
        self._targets = [_lib.require_cuda(target_cls, "target_cls").contiguous()]

    # ---- AP integration on the device (map.py:85-141) ----------------------------------------------------------------------
    def ap_table(self):
        """-> (ap [max_class+1, n_thr] f64 device tensor, NaN rows for unseen classes; pos_count [max_class+2] i32)."""
        if not self._targets:
            raise ValueError("need at least one array to concatenate")      # np.concatenate([]) at map.py:124
        targets = torch.cat(self._targets)
        dev = targets.device
        nthr = self.map_iou_values.size
        if self._dets:
            dets, correct = torch.cat(self._dets).contiguous(), torch.cat(self._correct).contiguous()
        else:
            raise ValueError("need at least one array to concatenate")      # map.py:122
        max_class = int(targets.max().item())                               # the one sync of fetch (it returns host arrays anyway)
        if max_class < 0 or max_class >= 65535:
            raise ValueError("class ids must be integers in [0, 65535)")
        ap = torch.empty(max_class + 1, nthr, dtype=torch.float64, device=dev)
        pos = torch.empty(max_class + 2, dtype=torch.int32, device=dev)
        lib = _lib.load()
        ws = _lib.workspace(lib.fvb_map_ap_workspace_bytes(dets.size(0), nthr, max_class), dev, "map_ap")
        with torch.cuda.device(dev):
            _lib.check(lib.fvb_map_ap_f64(_lib.dptr(dets), _lib.dptr(correct), dets.size(0), _lib.dptr(targets), targets.numel(),
                                          nthr, max_class, _lib.dptr(ap), _lib.dptr(pos), _lib.dptr(ws), _lib.stream()), "map_ap")
        return ap, pos

    def fetch(self):
        """-> (map_each_iou[n_thr], map_each_cls[n_cls], cls_ids); a class with targets but no detections scores 0.5."""
        ap, pos = self.ap_table()
        ap, pos = ap.cpu().numpy(), pos.cpu().numpy()
        uniq = [int(c) for c in np.nonzero(pos[:-1] > 0)[0]]                # np.unique(seen) for integer class ids (map.py:125)
        table = ap[uniq]
        return np.mean(table, axis=0), np.mean(table, axis=1), uniq         # map.py:138-141
