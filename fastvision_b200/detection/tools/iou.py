"""IoU family -- drop-in for detection/tools/IOU.py (torch branches, bug-compatible).

Every function is ONE fused kernel launch (fvb_iou_elementwise_f32 / fvb_iou_pairwise_f32).
Shapes follow the reference: element-wise functions return [n,1] except ``GIOU`` which returns [n]
(IOU.py:239); pairwise ones return [N,M].  ``variant='demo'`` selects demos/*/utils/iou.py arithmetic.
"""
import torch

from ... import _lib


def _check_pair(a, b, cols):
    a = _lib.require_cuda(a, "box1")
    b = _lib.require_cuda(b, "box2")
    if a.dim() != 2 or b.dim() != 2 or a.size(1) != cols or b.size(1) != cols:
        raise ValueError("expected [n,%d] boxes, got %s and %s" % (cols, tuple(a.shape), tuple(b.shape)))
    if a.device != b.device:
        raise ValueError("boxes on different devices")
    return a, b


def _elementwise(a, b, mode, kind, eps, variant="lib"):
    if mode not in _lib.BOX_MODES:
        raise Exception('mode must be xyxy or xywh or wh')
    a, b = _check_pair(a, b, 2 if mode == "wh" else 4)
    if a.size(0) != b.size(0):
        raise ValueError("element-wise IoU needs equal lengths, got %d and %d" % (a.size(0), b.size(0)))
    out = torch.empty(a.size(0), dtype=torch.float32, device=a.device)
    lib = _lib.load()
    with torch.cuda.device(a.device):
        _lib.check(lib.fvb_iou_elementwise_f32(_lib.dptr(a), _lib.dptr(b), a.size(0), _lib.BOX_MODES[mode],
                                               _lib.IOU_KINDS[kind], _lib.VARIANTS[variant], float(eps),
                                               _lib.dptr(out), _lib.stream()), "iou_elementwise")
    return out


def _pairwise(a, b, mode, kind, eps, variant="lib"):
    if mode not in _lib.BOX_MODES:
        raise Exception('mode must be xyxy or xywh or wh')
    a, b = _check_pair(a, b, 2 if mode == "wh" else 4)
    out = torch.empty(a.size(0), b.size(0), dtype=torch.float32, device=a.device)
    lib = _lib.load()
    with torch.cuda.device(a.device):
        _lib.check(lib.fvb_iou_pairwise_f32(_lib.dptr(a), a.size(0), _lib.dptr(b), b.size(0), _lib.BOX_MODES[mode],
                                            _lib.IOU_KINDS[kind], _lib.VARIANTS[variant], float(eps),
                                            _lib.dptr(out), _lib.stream()), "iou_pairwise")
    return out


def cal_iou(box1, box2, mode='xyxy', eps=1e-7):
    """IOU.py:7-15 -> [n,1]."""
    return _elementwise(box1, box2, mode, "iou", eps).view(-1, 1)


def cal_iou_batch(box1, box2, mode='xyxy', eps=1e-7):
    """IOU.py:17-25 -> [N,M]."""
    return _pairwise(box1, box2, mode, "iou", eps)


def xyxy_iou(xyxy1, xyxy2, eps=1e-7):
    return cal_iou(xyxy1, xyxy2, 'xyxy', eps)


def xywh_iou(xywh1, xywh2, eps=1e-7):
    return cal_iou(xywh1, xywh2, 'xywh', eps)


def wh_iou(wh1, wh2, eps=1e-7):
    return cal_iou(wh1, wh2, 'wh', eps)


def xyxy_iou_batch(xyxy1, xyxy2, eps=1e-7):
    return cal_iou_batch(xyxy1, xyxy2, 'xyxy', eps)


def xywh_iou_batch(xywh1, xywh2, eps=1e-7):
    return cal_iou_batch(xywh1, xywh2, 'xywh', eps)


def wh_iou_batch(wh1, wh2, eps=1e-7):
    return cal_iou_batch(wh1, wh2, 'wh', eps)


def _no_wh(mode):
    if mode not in ('xyxy', 'xywh'):
        raise Exception('mode must be xyxy or xywh')


def GIOU(box1, box2, mode='xyxy', eps=1e-7):
    """IOU.py:193-241 -> [n] (sic)."""
    _no_wh(mode)
    return _elementwise(box1, box2, mode, "giou", eps)


def GIOU_batch(box1, box2, mode='xyxy', eps=1e-7):
    """IOU.py:243-292 -> [N,M]."""
    _no_wh(mode)
    return _pairwise(box1, box2, mode, "giou", eps)


def DIOU(box1, box2, mode='xyxy', eps=1e-7, variant='lib'):
    """IOU.py:294-343 -> [n,1]."""
    _no_wh(mode)
    return _elementwise(box1, box2, mode, "diou", eps, variant).view(-1, 1)


def DIOU_batch(box1, box2, mode='xyxy', eps=1e-7, variant='lib'):
    """IOU.py:345-395 -> [N,M]."""
    _no_wh(mode)
    return _pairwise(box1, box2, mode, "diou", eps, variant)


def CIOU(box1, box2, mode='xyxy', eps=1e-7, variant='lib'):
    """IOU.py:397-440 -> [n,1]."""
    _no_wh(mode)
    return _elementwise(box1, box2, mode, "ciou", eps, variant).view(-1, 1)


def CIOU_batch(box1, box2, mode='xyxy', eps=1e-7, variant='lib'):
    """IOU.py:442-482 -> [N,M]."""
    _no_wh(mode)
    return _pairwise(box1, box2, mode, "ciou", eps, variant)
