"""Box conversions -- drop-in for detection/tools/BOX.py:4-26 (CUDA tensors only; one kernel each)."""
import torch

from ... import _lib


def _convert(x, op, height=1.0, width=1.0):
    x = _lib.require_cuda(x, "boxes")
    if x.dim() != 2 or x.size(1) != 4:
        raise ValueError("boxes must be [n,4], got %s" % (tuple(x.shape),))
    out = torch.empty_like(x)
    lib = _lib.load()
    with torch.cuda.device(x.device):
        _lib.check(lib.fvb_box_convert_f32(_lib.dptr(x), x.size(0), op, float(height), float(width), _lib.dptr(out),
                                           _lib.stream()), "box_convert")
    return out


def xywh2xyxy(xywh):
    """detection/tools/BOX.py:4-10.  Returns a new tensor; the input is never modified."""
    return _convert(xywh, 0)


def xyxy2xywh(xyxy):
    """detection/tools/BOX.py:12-18."""
    return _convert(xyxy, 1)


def xyxy2xywhn(xyxy, heigth, width):
    """detection/tools/BOX.py:20-26 (the reference spells the argument ``heigth``)."""
    return _convert(xyxy, 2, heigth, width)
