"""Oracle: box conversions and grid lattice.  TEST INFRASTRUCTURE ONLY.

Follows detection/tools/BOX.py:4-26 and detection/tools/GRID.py:4-31 of the reference.
torch fp32 on CPU; same arithmetic order (divide by 2, then add/subtract).
"""
import numpy as np
import torch


def xywh2xyxy(b):
    """detection/tools/BOX.py:4-10 -- x1=x-w/2, y1=y-h/2, x2=x+w/2, y2=y+h/2 (copy, never in place)."""
    if isinstance(b, np.ndarray):
        half = b[:, 2:4] / 2
        out = b.copy()
        out[:, 0:2] = b[:, 0:2] - half
        out[:, 2:4] = b[:, 0:2] + half
        return out
    half = b[:, 2:4] / 2
    out = b.clone()
    out[:, 0:2] = b[:, 0:2] - half
    out[:, 2:4] = b[:, 0:2] + half
    return out


def xyxy2xywh(b):
    """detection/tools/BOX.py:12-18."""
    out = b.copy() if isinstance(b, np.ndarray) else b.clone()
    out[:, 0:2] = (b[:, 0:2] + b[:, 2:4]) / 2
    out[:, 2:4] = b[:, 2:4] - b[:, 0:2]
    return out


def xyxy2xywhn(b, heigth, width):
    """detection/tools/BOX.py:20-26 (argument spelled ``heigth`` in the reference)."""
    out = b.copy() if isinstance(b, np.ndarray) else b.clone()
    out[:, 0] = ((b[:, 0] + b[:, 2]) / 2) / width
    out[:, 1] = ((b[:, 1] + b[:, 3]) / 2) / heigth
    out[:, 2] = (b[:, 2] - b[:, 0]) / width
    out[:, 3] = (b[:, 3] - b[:, 1]) / heigth
    return out


def grid(height, width, mode="xy", dtype="torch"):
    """detection/tools/GRID.py:4-31.

    numpy branch (``np.meshgrid`` is 'xy'-indexed): mode='yx' -> [H,W,2] with [...,0]=col, [...,1]=row;
    mode='xy' -> its transpose [W,H,2].  torch branch (``torch.meshgrid`` is 'ij'-indexed) has the
    opposite meaning: mode='yx' -> [W,H,2], mode='xy' -> [H,W,2].  Last dim is (x, y).
    """
    ys = np.arange(height)
    xs = np.arange(width)
    hw = np.stack(np.broadcast_arrays(xs[None, :], ys[:, None]), axis=-1)  # [H,W,2] = (x,y)
    wh = hw.transpose(1, 0, 2)
    if dtype == "torch":
        out = hw if mode == "xy" else wh
        return torch.from_numpy(np.ascontiguousarray(out)).long()
    return wh if mode == "xy" else hw
