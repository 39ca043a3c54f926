"""RPN proposal filter -- drop-in for ``RPN.filter_proposals`` (demos/faster_rcnn/models/rpn.py:168-208).

The reference decodes every anchor with ~25 ATen launches and then loops over the images in Python
(topk -> torchvision.ops.nms -> slice -> xyxy2xywh, one host sync per image).  Here the whole batch is
two launches (``fvb_rpn_proposals_f32``): padded outputs + counts; the ragged list is only built at
this API edge.  ``get_base_anchor`` / ``make_anchors_xywh`` mirror the small host helpers the demo
uses to describe the anchors (demos/faster_rcnn/utils/anchor_generator.py:4-14, rpn.py:160-166).
"""
import ctypes as C
import math

import torch

from ... import _lib


def get_base_anchor(scales, ratios):
    """demos/faster_rcnn/utils/anchor_generator.py:4-14 -> float32 [len(ratios)*len(scales), 2] (w,h) pixels."""
    out = []
    for r in ratios:
        for s in scales:
            w = math.sqrt(s ** 2 / r)
            out.append((w, s ** 2 / w))
    return torch.tensor(out, dtype=torch.float32).view(-1, 2)


def make_anchors_xywh(base_anchors, feature_height, feature_width, device="cpu"):
    """rpn.py:160-166 -> [1,fh,fw,A,4] = (x = column, y = row, w, h), feature units."""
    base = base_anchors.to(device=device, dtype=torch.float32).view(-1, 2)
    a = base.size(0)
    wh = base.view(1, 1, 1, a, 2).expand(1, feature_height, feature_width, a, 2)
    ys = torch.arange(feature_height, device=device, dtype=torch.float32).view(1, -1, 1, 1, 1)
    xs = torch.arange(feature_width, device=device, dtype=torch.float32).view(1, 1, -1, 1, 1)
    xy = torch.cat([xs.expand(1, feature_height, feature_width, a, 1), ys.expand(1, feature_height, feature_width, a, 1)], 4)
    return torch.cat([xy, wh], dim=4)


def _base_from_anchor_xywh(anchor_xywh):
    """The kernel regenerates cell centres from indices; only the A (w,h) pairs are read from ``anchor_xywh``."""
    if anchor_xywh.dim() == 2:
        return anchor_xywh.detach().float().cpu().reshape(-1, 2)
    if anchor_xywh.dim() != 5 or anchor_xywh.size(-1) != 4:
        raise ValueError("anchor_xywh must be [1,H,W,A,4] (rpn.py:160-166) or base anchors [A,2]")
    return anchor_xywh[0, 0, 0, :, 2:4].detach().float().cpu()


def filter_proposals_batched(cls, dxdydwdh, anchor_xywh, rpn_pre_nms_top_n=2000, rpn_post_nms_top_n=2000,
                             rpn_nms_thresh=0.7, want_idx=False, want_decoded=False):
    """cls [B,H,W,A,2], dxdydwdh [B,H,W,A,4] -> (xywh [B,post_n,4], cnt [B] int32[, anchor idx [B,post_n] int32]).

    No host sync; entries past ``cnt[b]`` are undefined.  ``want_decoded`` (parity tests) appends the kernel's own
    per-anchor results, copied out of the workspace: clamped xyxy boxes [B,n,4] and foreground scores [B,n] -- the
    tensors the reference's ``topk -> nms`` (rpn.py:193-198) would be run on.
    """
    cls = _lib.require_cuda(cls, "cls")
    reg = _lib.require_cuda(dxdydwdh, "dxdydwdh")
    if reg.dim() != 5 or reg.size(-1) != 4 or cls.shape[:4] != reg.shape[:4] or cls.size(-1) != 2:
        raise ValueError("cls must be [B,H,W,A,2] and dxdydwdh [B,H,W,A,4]; got %s / %s" % (tuple(cls.shape), tuple(reg.shape)))
    b, fh, fw, a, _ = reg.shape
    base = _base_from_anchor_xywh(anchor_xywh).contiguous()
    if base.size(0) != a:
        raise ValueError("anchor_xywh describes %d anchors per cell, tensors have %d" % (base.size(0), a))
    post = int(rpn_post_nms_top_n)
    dev = reg.device
    out = torch.empty(b, post, 4, dtype=torch.float32, device=dev)
    idx = torch.empty(b, post, dtype=torch.int32, device=dev) if want_idx else None
    cnt = torch.empty(b, dtype=torch.int32, device=dev)
    lib = _lib.load()
    ws = _lib.workspace(lib.fvb_rpn_workspace_bytes(b, fh, fw, a), dev, "rpn")
    base_c = (C.c_float * (2 * a))(*[float(v) for v in base.view(-1).tolist()])
    with torch.cuda.device(dev):
        _lib.check(lib.fvb_rpn_proposals_f32(_lib.dptr(cls), _lib.dptr(reg), base_c, b, fh, fw, a, int(rpn_pre_nms_top_n),
                                             post, float(rpn_nms_thresh), _lib.dptr(out), _lib.dptr(idx), _lib.dptr(cnt),
                                             _lib.dptr(ws), _lib.stream()), "rpn_proposals")
    res = (out, cnt, idx) if want_idx else (out, cnt)
    if want_decoded:
        res = res + _decoded_from_workspace(ws, b, fh * fw * a)
    return res


def _decoded_from_workspace(ws, b, n):
    """Workspace layout of fvb_rpn_proposals_f32 (csrc/rpn.cu): 64-bit sort keys [B][2][n] (after the 4-pass sort the
    ordered keys are back in the first half: high word = order-reversing image of the score bits, low word = anchor index),
    then the clamped xyxy boxes [B][n] float4."""
    align = lambda x: (x + 255) // 256 * 256
    keys = ws[:b * 2 * n * 8].view(torch.int64).view(b, 2, n)[:, 0]
    boxes = ws[align(b * 2 * n * 8):align(b * 2 * n * 8) + b * n * 16].view(torch.float32).view(b, n, 4).clone()
    idx = keys & 0xffffffff
    u = (~(keys >> 32)) & 0xffffffff                        # ascending-orderable image of the float (nms.cuh desc_key)
    bits = torch.where((u & 0x80000000) != 0, u & 0x7fffffff, (~u) & 0xffffffff)
    sc_sorted = torch.where(bits >= 2 ** 31, bits - 2 ** 32, bits).to(torch.int32).view(torch.float32)
    scores = torch.empty(b, n, dtype=torch.float32, device=ws.device)
    scores.scatter_(1, idx, sc_sorted)
    return boxes, scores


def filter_proposals(cls, dxdydwdh, anchor_xywh, feature_height=None, feature_width=None, rpn_pre_nms_top_n=2000,
                     rpn_post_nms_top_n=2000, rpn_nms_thresh=0.7):
    """Reference signature (rpn.py:168) + the three thresholds the RPN module holds (rpn.py:75-77).

    Returns the reference's list of ``[k_i, 4]`` xywh tensors (feature units), one per image.
    """
    if feature_height is not None and (int(feature_height) != dxdydwdh.size(1) or int(feature_width) != dxdydwdh.size(2)):
        raise ValueError("feature_height/width do not match the tensors")
    out, cnt = filter_proposals_batched(cls, dxdydwdh, anchor_xywh, rpn_pre_nms_top_n, rpn_post_nms_top_n, rpn_nms_thresh)
    cnt = cnt.cpu().tolist()  # the one sync: ragged outputs
    return [out[i, :k] for i, k in enumerate(cnt)]
