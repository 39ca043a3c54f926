"""Cell-offset lattice -- drop-in for detection/tools/GRID.py:4-31.

Host-side helper only: inside the decode kernel the lattice is index arithmetic, never a tensor.
``offset`` is the function detection/models/yolov3.py:4 imports but the reference never defines;
per SURVEY F1/F2 it is ``grid(h, w, mode, dtype='numpy')``.
"""
import numpy as np
import torch


def grid(height, width, mode='xy', dtype='torch'):
    """Last dim is (x, y).  torch: 'xy' -> [H,W,2], 'yx' -> [W,H,2]; numpy: the opposite (GRID.py:6-29)."""
    ys, xs = np.meshgrid(np.arange(height), np.arange(width), indexing="ij")
    hw = np.stack([xs, ys], axis=-1)            # [H,W,2]
    wh = np.ascontiguousarray(hw.transpose(1, 0, 2))
    if dtype == 'torch':
        return torch.from_numpy(hw if mode == 'xy' else wh).long()
    return wh if mode == 'xy' else hw


def offset(height, width, mode='yx'):
    return grid(height, width, mode=mode, dtype='numpy')
