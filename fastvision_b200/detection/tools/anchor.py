"""Anchor k-means -- drop-in for ``KMeans`` / ``AnchorGenerator`` of detection/tools/ANCHOR.py:11-120.

``KMeans(xs, k).fit(iters)`` keeps the reference's contract: ``xs`` is a float32 numpy array [n,2] of normalised (w, h), it
is shuffled IN PLACE with ``np.random.shuffle`` (so a seeded run picks the same initial centres as the reference), the first
``k`` rows are the initial centres, and ``fit`` returns ``(centers [k,2] float32 numpy, categories [n] int64 numpy, 1-based)``.
Each Lloyd iteration is three launches on the device (fvb_kmeans_step_f32) instead of an [n,k] numpy distance matrix and
k boolean-mask means.  ``AnchorGenerator`` walks the data loaders exactly like the reference, then sorts the centres by area
and scales them to the input size; plotting and the cache file are kept (the plot only when matplotlib is importable).
"""
import os

import numpy as np
import torch

from ... import _lib


class KMeans():
    def __init__(self, xs, k=9, device=None):
        self.num_samples = len(xs)
        self.samples = xs
        self.k = k
        np.random.shuffle(self.samples)                      # ANCHOR.py:17 (in place, global numpy RNG)
        self.centers = self.samples[:k, :]
        self.device = torch.device(device or ("cuda:%d" % torch.cuda.current_device()))
        if self.device.type != "cuda":
            raise RuntimeError("KMeans runs on CUDA only (no CPU fallback)")
        self._d_samples = torch.from_numpy(np.ascontiguousarray(self.samples, dtype=np.float32)).to(self.device)
        self._d_centers = torch.from_numpy(np.ascontiguousarray(self.centers, dtype=np.float32)).to(self.device)
        self._d_next = torch.empty_like(self._d_centers)
        self._d_cat = torch.empty(self.num_samples, dtype=torch.int64, device=self.device)
        self.categories = None

    def cal_distance(self, xs, centers):
        """1 - iou(x, center) (ANCHOR.py:21-24), [n,k] on the device."""
        from .iou import wh_iou_batch
        xs = torch.as_tensor(xs, dtype=torch.float32, device=self.device)
        centers = torch.as_tensor(centers, dtype=torch.float32, device=self.device)
        return 1 - wh_iou_batch(xs, centers)

    def _fit(self):
        lib = _lib.load()
        ws = _lib.workspace(lib.fvb_kmeans_workspace_bytes(self.k), self.device, "kmeans")
        with torch.cuda.device(self.device):
            _lib.check(lib.fvb_kmeans_step_f32(_lib.dptr(self._d_samples), self.num_samples, _lib.dptr(self._d_centers), self.k,
                                               1e-7, _lib.dptr(self._d_cat), _lib.dptr(self._d_next), _lib.dptr(ws),
                                               _lib.stream()), "kmeans_step")
        self._d_centers, self._d_next = self._d_next, self._d_centers

    def fit(self, iters):
        for _ in range(iters):
            self._fit()
        self.centers = self._d_centers.cpu().numpy().reshape([-1, 2])
        self.categories = self._d_cat.cpu().numpy()
        return self.centers, self.categories


class AnchorGenerator():
    """Same constructor and ``get_anchors()`` result as ANCHOR.py:50-120: collects the normalised (w, h) of every label the data
    loaders yield, clusters them on the device, orders the centres by decreasing area and scales them to the input size.  The
    cache file keeps the reference's format (``str`` of a nested list in ``<cache>/anchor.txt``)."""

    def __init__(self, data_loaders: list, k=9, iters=100, num_workers=1, plot=True, cache='./cache', use_cache=False):
        self.data_loaders, self.k, self.iters, self.num_workers = data_loaders, k, iters, num_workers
        self.cache_dir = cache
        self.cache = os.path.join(cache, 'anchor.txt')
        self.use_cache, self.plot = use_cache, plot
        self.input_height = self.input_widht = None          # (attribute spelled as in the reference, ANCHOR.py:77)

    def load_data(self):
        """[sum T, 2] float array of label (w, h); remembers the input size of the last batch seen (ANCHOR.py:62-86)."""
        chunks = []
        for loader in self.data_loaders:
            for images, labels in loader:
                self.input_height, self.input_widht = int(images.shape[2]), int(images.shape[3])
                chunks.append(labels[:, 4:].detach().cpu().numpy())
        return np.concatenate(chunks, axis=0)

    def load_cache(self):
        import ast
        with open(self.cache) as fh:
            return ast.literal_eval(fh.read())

    def _plot(self, wh, categories, centers):
        try:
            from matplotlib import pyplot as plt
        except ImportError:
            return
        for cid in range(1, self.k + 1):
            sel = categories == cid
            plt.scatter(wh[sel, 0], wh[sel, 1], alpha=0.8)
        plt.scatter(centers[:, 0], centers[:, 1], c='black', marker='x')
        plt.savefig(os.path.join(self.cache_dir, 'anchor.png'))

    def get_anchors(self):
        if self.use_cache:
            print(f'Use anchor from cache {self.cache}')
            return np.asarray(self.load_cache(), dtype=float).reshape(-1, 2)
        wh = np.asarray(self.load_data(), dtype=np.float32).reshape(-1, 2)
        centers, categories = KMeans(xs=wh, k=self.k).fit(iters=self.iters)
        order = sorted(range(len(centers)), key=lambda i: -float(centers[i][0]) * float(centers[i][1]))   # ANCHOR.py:106, stable
        centers = np.asarray([centers[i].tolist() for i in order], dtype=float).reshape(-1, 2)
        os.makedirs(self.cache_dir, exist_ok=True)
        if self.plot:
            self._plot(wh, categories, centers)
        pixels = centers * np.array([self.input_widht, self.input_height])                                 # ANCHOR.py:116
        with open(self.cache, 'w') as fh:
            fh.write(str(pixels.tolist()))
        return pixels
