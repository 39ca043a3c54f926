"""Confidence filter + NMS -- drop-in for detection/tools/NMS.py:5-23 and the demo flavours.

``non_max_suppression`` keeps the reference signature and return shapes (one image per call).
``non_max_suppression_batched`` is what a B200 wants: every image of a decoded [B,N,K] tensor in ONE
launch, padded outputs + counts, no host sync.  ``nms`` is the batched ``torchvision.ops.nms``
equivalent.  No torchvision at run time.
"""
import torch

from ... import _lib


def non_max_suppression_batched(results, conf_thres=0.25, iou_thres=0.45, max_det=300, flavour="lib",
                                cand_bitmap=None, cand_records=None, clear_bitmap=True, max_wh=4096.0,
                                want_rows=False, out=None, tile_sync=None, tiles_per_image=0, ws=None, wide_cta=False):
    """results [B,N,K] decoded -> (boxes[B,max_det,4] xyxy, scores[B,max_det], cls[B,max_det] i64, cnt[B] i32[, rows]).

    Entries past cnt[b] are undefined.  ``cand_bitmap`` / ``cand_records`` are the [B, ceil(N/32)] int32 bitmap
    and [B,N,8] records written by the decode kernel for the same ``conf_thres`` (``DecodeContext.bitmap()`` /
    ``.records()``); without them one extra scoring launch derives both from ``results``.
    ``tile_sync`` / ``tiles_per_image`` (``DecodeContext.tile_sync()`` / ``.tiles_per_image``): this call follows the
    ``yolov3_decode(..., tile_sync=...)`` launch of the same tensors DIRECTLY on the current stream and is launched as its
    programmatic dependent -- image b's NMS starts as soon as image b is decoded (``fvb_yolo_nms_after_decode_f32``).
    ``wide_cta`` (with ``tile_sync``): the decode geometry leaves no room for an NMS CTA beside a decode CTA (FVB_NMS_WIDE_CTA).
    ``ws``: caller-owned workspace (uint8, >= ``fvb_yolo_nms_workspace_bytes``) instead of the process-wide scratch.
    """
    results = _lib.require_cuda(results, "results")
    if results.dim() != 3:
        raise ValueError("results must be [B,N,K], got %s" % (tuple(results.shape),))
    b, n, k = results.shape
    dev = results.device
    if out is None:
        boxes = torch.empty(b, max_det, 4, dtype=torch.float32, device=dev)
        scores = torch.empty(b, max_det, dtype=torch.float32, device=dev)
        cls = torch.empty(b, max_det, dtype=torch.int64, device=dev)
        cnt = torch.empty(b, dtype=torch.int32, device=dev)
        rows = torch.empty(b, max_det, dtype=torch.int32, device=dev) if want_rows else None
    else:
        boxes, scores, cls, cnt, rows = out
    lib = _lib.load()
    need = lib.fvb_yolo_nms_workspace_bytes(b, n)
    if ws is None:
        ws = _lib.workspace(need, dev, "yolo_nms")
    elif ws.numel() < need:
        raise ValueError("yolo_nms workspace of %d bytes, need %d" % (ws.numel(), need))
    with torch.cuda.device(dev):
        _lib.check(lib.fvb_yolo_nms_after_decode_f32(_lib.dptr(results), b, n, k, float(conf_thres), float(iou_thres), int(max_det),
                                                     _lib.NMS_FLAVOURS[flavour], float(max_wh), _lib.dptr(cand_bitmap),
                                                     _lib.dptr(cand_records), (1 if clear_bitmap else 0) | (2 if wide_cta else 0),
                                                     _lib.dptr(boxes),
                                                     _lib.dptr(scores), _lib.dptr(cls), _lib.dptr(rows), _lib.dptr(cnt),
                                                     _lib.dptr(tile_sync), int(tiles_per_image) if tile_sync is not None else 0,
                                                     _lib.dptr(ws), _lib.stream()), "yolo_nms")
    if want_rows:
        return boxes, scores, cls, cnt, rows
    return boxes, scores, cls, cnt


def non_max_suppression(prediction, conf_thres=0.25, iou_thres=0.45, max_det=300):
    """detection/tools/NMS.py:5-23 -> (scores[k,1], categories[k,1] int64, boxes[k,4] xyxy), score-descending.

    As in the reference a [B,N,K] input is flattened across images (NMS.py:7-8) and an input with no
    candidate returns three empty CPU float tensors (NMS.py:10-11).
    """
    pred = _lib.require_cuda(prediction, "prediction")
    k = pred.size(-1)
    flat = pred.reshape(1, -1, k)
    boxes, scores, cls, cnt = non_max_suppression_batched(flat, conf_thres, iou_thres, max_det, "lib")
    kept = int(cnt[0])  # the reference syncs here too (boolean-mask indexing)
    if kept == 0:
        return torch.Tensor().view(-1, 1), torch.Tensor().view(-1, 1), torch.Tensor().view(-1, 4)
    return scores[0, :kept].view(-1, 1), cls[0, :kept].view(-1, 1), boxes[0, :kept].view(-1, 4)


def non_max_suppression_demo(prediction, conf_thres=0.25, iou_thres=0.45, max_det=300):
    """demos/yolov3_u/utils/nms.py:5-53 -> [k,6] = [x1,y1,x2,y2,obj,cat]; class-aware, boxes already xyxy."""
    pred = _lib.require_cuda(prediction, "prediction")
    boxes, scores, cls, cnt = non_max_suppression_batched(pred.unsqueeze(0), conf_thres, iou_thres, max_det, "demo")
    kept = int(cnt[0])
    if kept == 0:
        return torch.zeros((0, 6), device=pred.device)
    return torch.cat([boxes[0, :kept], scores[0, :kept, None], cls[0, :kept, None].float()], dim=1)


def non_max_suppression_batch(prediction_batch, conf_thres=0.25, iou_thres=0.45, max_det=300):
    """demos/yolov3_u/utils/nms.py:55-98 -> list of CPU [k,6] = [x1,y1,x2,y2,score,cat] (one launch per call)."""
    stacked = torch.stack([_lib.require_cuda(p, "prediction") for p in prediction_batch], 0)
    boxes, scores, cls, cnt = non_max_suppression_batched(stacked, conf_thres, iou_thres, max_det, "demo_batch")
    out6 = torch.cat([boxes, scores[..., None], cls[..., None].float()], dim=2).cpu()
    cnt = cnt.cpu().tolist()
    return [out6[i, :cnt[i]] for i in range(len(cnt))]


def non_max_suppression_frcnn(prediction, conf_thres=0.25, iou_thres=0.45, max_det=300, max_wh=4096, max_nms=30000):
    """demos/faster_rcnn/utils/nms.py:5-39: rows [x1,y1,x2,y2,cat,score] -> the kept rows, score-descending.

    The filter / gather glue stays torch indexing (as in the reference, one sync at the boolean mask); the
    suppression itself is the segmented NMS kernel on gap-offset boxes (``box + cat*max_wh`` in fp32, :30-31).
    """
    pred = _lib.require_cuda(prediction, "prediction")
    pred = pred[pred[:, 5] > conf_thres]                                    # :18
    if pred.size(0) == 0:
        return torch.zeros((0, 6), device=pred.device)
    if pred.size(0) > max_nms:                                              # :23-24 (top by column 0, sic)
        pred = pred[pred[:, 0].argsort(descending=True)[:max_nms]]
    boxes = (pred[:, :4] + pred[:, 4:5] * max_wh).contiguous()
    keep = nms(boxes, pred[:, 5].contiguous(), iou_thres, max_keep=max_det)
    return pred[keep]


def nms(boxes, scores, iou_threshold, seg_offsets=None, max_keep=None):
    """Batched ``torchvision.ops.nms``: kept indices, score-descending (ties: lower index first).

    Without ``seg_offsets`` behaves like the torchvision call (one segment, int64 indices).  With
    ``seg_offsets`` ([S+1] int32) returns (keep_idx[S,max_keep] int32 relative to each segment, keep_cnt[S]).
    """
    boxes = _lib.require_cuda(boxes, "boxes")
    scores = _lib.require_cuda(scores, "scores")
    n = boxes.size(0)
    dev = boxes.device
    single = seg_offsets is None
    if single:
        seg_offsets = torch.tensor([0, n], dtype=torch.int32, device=dev)
    segs = seg_offsets.numel() - 1
    if max_keep is None:
        max_keep = max(n, 1) if single else 2000
    lib = _lib.load()
    # never truncates: a single segment may keep all n boxes (as torchvision.ops.nms does); kept lists longer than what fits
    # in shared memory live in the workspace (csrc/nms.cu: kKeepSmem)
    max_keep = max(1, int(max_keep))
    keep = torch.empty(segs, max_keep, dtype=torch.int32, device=dev)
    cnt = torch.empty(segs, dtype=torch.int32, device=dev)
    ws = _lib.workspace(lib.fvb_nms_segmented_workspace_bytes(n, segs), dev, "seg_nms")
    with torch.cuda.device(dev):
        _lib.check(lib.fvb_nms_segmented_f32(_lib.dptr(boxes), _lib.dptr(scores), _lib.dptr(seg_offsets), segs, n,
                                             float(iou_threshold), max_keep, _lib.dptr(keep), _lib.dptr(cnt),
                                             _lib.dptr(ws), _lib.stream()), "nms_segmented")
    if single:
        return keep[0, :int(cnt[0])].long()
    return keep, cnt
