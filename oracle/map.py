"""Oracle: COCO-style mAP matcher and AP integration.  TEST INFRASTRUCTURE ONLY.

``MapOracle`` restates metrics/map.py:16-141 with the same torch/numpy calls in the same order
(np.float/np.long spelled float/np.int64), so it is bit-identical to the reference
including the order-dependent dedupe (:74-76) and the AP=0.5 quirk for classes with targets
but no detections.  ``match_rule`` is the equivalent deterministic rule the CUDA
matcher implements:
  for each prediction p: t*(p) = argmax-IoU target among (IoU > thr[0] and same class)
  (ties: the first in argsort(-iou) order, reproduced by evaluating on the same sorted list);
  for each target t the winner is the lowest-index p with t*(p) = t;
  correct[p,k] = float32 IoU > thr_k (float64).
"""
import numpy as np
import torch

from .iou import cal_iou_batch


def match_rule(y_pred, y_true, thresholds):
    """Loop form of the matcher; returns the [M, len(thr)] boolean block of ``correct``."""
    thr = np.asarray(thresholds, dtype=np.float64)
    m = y_pred.size(0)
    out = np.zeros((m, thr.size), dtype=bool)
    if m == 0 or y_true.size(0) == 0:
        return out
    iou = cal_iou_batch(y_true[:, 1:], y_pred[:, 2:], mode="xyxy").numpy()   # [N,M] fp32
    tc = y_true[:, 0].numpy()
    pc = y_pred[:, 0].numpy()
    n = iou.shape[0]
    best_t = np.full(m, -1, dtype=np.int64)
    for p in range(m):
        best = -1.0
        for t in range(n):
            if tc[t] == pc[p] and float(iou[t, p]) > thr[0]:
                if float(iou[t, p]) > best:      # strict: first (lowest t) wins ties
                    best, best_t[p] = float(iou[t, p]), t
    taken = set()
    for p in range(m):                            # lowest-index p claims its target
        t = best_t[p]
        if t >= 0 and t not in taken:
            taken.add(t)
            out[p] = iou[t, p].astype(np.float64) > thr
    return out


class MapOracle:
    """metrics/map.py:6-141."""

    def __init__(self, map_iou_values):
        self.map_iou_values = map_iou_values
        self.correct_all_images = []
        self.seen_all_targets_cls = []

    def process_one(self, y_pred, y_true):
        """map.py:16-83.  y_pred[M,6]=[cls,conf,x1,y1,x2,y2]; y_true[N,5]=[cls,x1,y1,x2,y2]."""
        nthr = len(self.map_iou_values)
        correct = np.zeros([y_pred.size(0), 2 + nthr], dtype=float)              # :34
        p_cls, p_conf, p_box = y_pred[:, 0], y_pred[:, 1], y_pred[:, 2:]
        t_cls, t_box = y_true[:, 0], y_true[:, 1:]
        if t_cls.size(0) != 0:                                                    # :43-44
            self.seen_all_targets_cls.append(t_cls.detach().cpu().numpy())
        if y_pred.size(0) == 0:                                                   # :46-47
            return
        iou = cal_iou_batch(t_box, p_box, mode="xyxy")                            # :50 [N,M]
        hit = ((iou > self.map_iou_values[0]) & (t_cls[:, None] == p_cls))        # :51-57
        ti, pi = np.where(hit.numpy())                                            # :60 row-major
        rows = torch.cat([torch.from_numpy(ti).float().view(-1, 1), torch.from_numpy(pi).float().view(-1, 1),
                          iou[ti, pi].view(-1, 1), t_cls[ti].view(-1, 1), p_conf[pi].view(-1, 1)], dim=1)
        rows = rows.view(-1, 5).numpy()                                           # :71-72
        rows = rows[np.argsort(-rows[:, 2]), ...]                                 # :74
        rows = rows[np.unique(rows[:, 1], return_index=True)[1], ...]             # :75
        rows = rows[np.unique(rows[:, 0], return_index=True)[1], ...]             # :76
        correct[:, 0] = p_conf.numpy()                                            # :79
        correct[:, 1] = p_cls.numpy()                                             # :80
        correct[rows[:, 1].astype(np.int64), 2:] = rows[:, 2:3] > self.map_iou_values   # :81
        self.correct_all_images.append(correct)

    @staticmethod
    def compute_ap(recall, precision):
        """map.py:85-94 ('coco' method)."""
        r = np.concatenate(([0.0], recall, [1.0]))
        p = np.concatenate(([1.0], precision, [0.0]))
        env = np.flip(np.maximum.accumulate(p[::-1]))
        x = np.linspace(0, 1, 101)
        trapz = getattr(np, "trapezoid", None) or np.trapz
        return trapz(np.interp(x, r, env), x)

    def _ap_per_class(self, total_positive, correct):
        """map.py:104-118."""
        ap = np.zeros((len(self.map_iou_values),), dtype=float)
        tp = np.cumsum(correct, axis=0)
        fn = total_positive - tp
        fp = np.cumsum(1 - correct, axis=0)
        recall = tp / (tp + fn + 1e-16)
        precision = tp / (tp + fp + 1e-16)
        for k in range(correct.shape[1]):
            ap[k] = self.compute_ap(recall[:, k], precision[:, k])
        return ap

    def fetch(self):
        """map.py:120-141 -> (map_each_iou[n_thr], map_each_cls[n_cls], cls_ids)."""
        correct = np.concatenate(self.correct_all_images, axis=0)
        seen = np.concatenate(self.seen_all_targets_cls, axis=0)
        uniq = np.unique(seen).tolist()
        table = np.zeros((len(uniq), len(self.map_iou_values)), dtype=float)
        for c in uniq:
            cur = correct[correct[:, 1] == c, ...]
            cur = cur[np.argsort(-cur[:, 0]), ...]                                # :131
            table[uniq.index(c)] = self._ap_per_class(np.sum(seen == c), cur[:, 2:])
        return np.mean(table, axis=0), np.mean(table, axis=1), [int(c) for c in uniq]
