"""COCO-style mAP -- drop-in for CalculateMAP, metrics/map.py:6-141.

``process_one`` keeps the reference signature; ``process_batch`` matches every image of a batch in ONE
kernel launch (fvb_map_match_f32) without the per-image D2H + numpy argsort/unique.  ``fetch`` is the
reference's host-side AP integration (float64 numpy, map.py:85-141), restated: per seen class, rows sorted
by confidence, cumulative TP/FP, precision envelope, 101-point interpolation.
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
        self.correct_all_images = []     # float64 [M, 2 + n_thr] blocks: [conf, cls, correct...]
        self.seen_all_targets_cls = []

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
        """Batched ``process_one``: appends one [sum M, 2+n_thr] block and the target classes of the batch."""
        correct = self.match(dets, det_off, gts, gt_off)
        if gts.size(0):
            self.seen_all_targets_cls.append(gts[:, 0].detach().cpu().numpy())
        if dets.size(0) == 0:
            return
        block = np.zeros([dets.size(0), 2 + self.map_iou_values.size], dtype=np.float64)
        host = dets[:, :2].detach().cpu().numpy()
        block[:, 0] = host[:, 1]
        block[:, 1] = host[:, 0]
        block[:, 2:] = correct.cpu().numpy()
        self.correct_all_images.append(block)

    def process_one(self, y_pred, y_true):
        """metrics/map.py:16-83.  y_pred[M,6]=[cls,conf,x1,y1,x2,y2]; y_true[N,5]=[cls,x1,y1,x2,y2]."""
        y_pred = _lib.require_cuda(y_pred, "y_pred").view(-1, 6)
        y_true = _lib.require_cuda(y_true, "y_true").view(-1, 5)
        dev = y_pred.device
        det_off = torch.tensor([0, y_pred.size(0)], dtype=torch.int32, device=dev)
        gt_off = torch.tensor([0, y_true.size(0)], dtype=torch.int32, device=dev)
        self.process_batch(y_pred, det_off, y_true, gt_off)

    # ---- host AP integration (float64, identical arithmetic to map.py:85-141) -------------------------------
    @staticmethod
    def compute_ap(recall, precision, method='coco'):
        m_recall = np.concatenate(([0.0], recall, [1.0]))
        m_precision = np.concatenate(([1.0], precision, [0.0]))
        envelope = np.flip(np.maximum.accumulate(m_precision[::-1]))
        if method == 'coco':
            x = np.linspace(0, 1, 101)
            integrate = getattr(np, "trapezoid", None) or np.trapz
            return integrate(np.interp(x, m_recall, envelope), x)
        if method == 'voc2009':
            i = np.where(m_recall[1:] != m_recall[:-1])[0]
            return np.sum((m_recall[i + 1] - m_recall[i]) * envelope[i + 1])
        raise Exception('Not complete')

    def _ap_per_class(self, total_positive, correct):
        ap = np.zeros((len(self.map_iou_values),), dtype=np.float64)
        tp = np.cumsum(correct, axis=0)
        fn = total_positive - tp
        fp = np.cumsum(1 - correct, axis=0)
        recall = tp / (tp + fn + 1e-16)
        precision = tp / (tp + fp + 1e-16)
        for k in range(correct.shape[1]):
            ap[k] = self.compute_ap(recall[:, k], precision[:, k])
        return ap

    def fetch(self):
        """-> (map_each_iou[n_thr], map_each_cls[n_cls], cls_ids); a class with targets but no detections scores 0.5."""
        correct = np.concatenate(self.correct_all_images, axis=0)
        seen = np.concatenate(self.seen_all_targets_cls, axis=0)
        uniq = np.unique(seen).tolist()
        table = np.zeros((len(uniq), len(self.map_iou_values)), dtype=np.float64)
        for i, c in enumerate(uniq):
            cur = correct[correct[:, 1] == c, ...]
            cur = cur[np.argsort(-cur[:, 0]), ...]
            table[i] = self._ap_per_class(np.sum(seen == c), cur[:, 2:])
        return np.mean(table, axis=0), np.mean(table, axis=1), [int(c) for c in uniq]
