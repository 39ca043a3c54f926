"""Oracle: anchor k-means.  TEST INFRASTRUCTURE ONLY.

Follows KMeans of detection/tools/ANCHOR.py:11-46 with numpy float32 in the reference's order (np.random.shuffle in place,
first k rows as initial centres, distance 1 - wh_iou_batch numpy branch (IOU.py:166-175), np.argmin + 1, np.mean per cluster,
an empty cluster keeps its centre)."""
import numpy as np


def wh_iou_batch_numpy(wh1, wh2, eps=1e-7):
    """detection/tools/IOU.py:166-175 (numpy branch)."""
    area1 = wh1[:, 0] * wh1[:, 1]
    area2 = wh2[:, 0] * wh2[:, 1]
    inter = np.minimum(wh1[:, None, 0], wh2[:, 0]) * np.minimum(wh1[:, None, 1], wh2[:, 1])
    union = area1[:, None] + area2 - inter + eps
    return inter / union


class KMeans:
    def __init__(self, xs, k=9):
        self.num_samples = len(xs)
        self.samples = xs
        self.k = k
        np.random.shuffle(self.samples)                                     # :17
        self.centers = self.samples[:k, :]                                  # :19

    def fit(self, iters):
        for _ in range(iters):
            self._fit()
        return self.centers, self.categories

    def _fit(self):
        distance = 1 - wh_iou_batch_numpy(self.samples, self.centers)      # :22-24, :34
        self.categories = np.argmin(distance, axis=1) + 1                   # :35
        new_centers = []
        for category_id in range(1, self.k + 1):
            sel = self.samples[self.categories == category_id]
            if sel.shape[0] == 0:
                new_centers.append(self.centers[category_id - 1, :])        # :39-40
            else:
                new_centers.append([np.mean(sel[:, 0]), np.mean(sel[:, 1])])   # :42-44
        self.centers = np.array(new_centers).reshape([-1, 2])               # :46
