"""ctypes binding of libfvb200.so (include/fvb200.h).  No fallback: a missing library or device raises."""
import ctypes as C
import os
import threading

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libfvb200.so")

MAX_LEVELS = 4
MAX_ANCHORS = 16

BOX_MODES = {"xyxy": 0, "xywh": 1, "wh": 2}
IOU_KINDS = {"iou": 0, "giou": 1, "diou": 2, "ciou": 3}
VARIANTS = {"lib": 0, "demo": 1}
DECODE_FORMS = {"v3": 0, "v5": 1}
HEAD_LAYOUTS = {"bahwk": 0, "nchw": 1}
NMS_FLAVOURS = {"lib": 0, "demo": 1, "demo_batch": 2}
REDUCTIONS = {"mean": 0, "sum": 1}
DEMO_LOSS_FLAVOURS = {"ship": 0, "u": 1}


class Geom(C.Structure):
    _fields_ = [
        ("levels", C.c_int32), ("batch", C.c_int32), ("anchors", C.c_int32), ("channels", C.c_int32),
        ("height", C.c_int32 * MAX_LEVELS), ("width", C.c_int32 * MAX_LEVELS),
        ("stride", C.c_float * MAX_LEVELS),
        ("anchor_w", (C.c_float * MAX_ANCHORS) * MAX_LEVELS),
        ("anchor_h", (C.c_float * MAX_ANCHORS) * MAX_LEVELS),
        ("head_layout", C.c_int32),
    ]


_P = C.c_void_p
_SIGS = {
    "fvb_abi_version": (C.c_int, []),
    "fvb_last_error": (C.c_char_p, []),
    "fvb_launch_count": (C.c_uint64, []),
    "fvb_yolo_rows_per_image": (C.c_int, [C.POINTER(Geom)]),
    "fvb_yolo_bitmap_words": (C.c_int, [C.POINTER(Geom)]),
    "fvb_yolo_decode_partials": (C.c_int, [C.POINTER(Geom)]),
    "fvb_yolo_decode_workspace_bytes": (C.c_size_t, []),
    "fvb_yolo_decode_f32": (C.c_int, [C.POINTER(Geom), C.POINTER(_P), C.c_int, C.c_int, _P, C.c_float, _P, _P, _P, _P, _P]),
    "fvb_yolo_decode_tiles_per_image": (C.c_int, [C.POINTER(Geom)]),
    "fvb_yolo_decode_leaves_room_for_nms": (C.c_int, [C.POINTER(Geom)]),
    "fvb_yolo_decode_sync_f32": (C.c_int, [C.POINTER(Geom), C.POINTER(_P), C.c_int, C.c_int, _P, C.c_float, _P, _P, _P, _P, _P, _P]),
    "fvb_box_convert_f32": (C.c_int, [_P, C.c_int64, C.c_int, C.c_float, C.c_float, _P, _P]),
    "fvb_iou_elementwise_f32": (C.c_int, [_P, _P, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_float, _P, _P]),
    "fvb_iou_pairwise_f32": (C.c_int, [_P, C.c_int64, _P, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_float, _P, _P]),
    "fvb_reduce_workspace_bytes": (C.c_size_t, [C.c_int64]),
    "fvb_iou_loss_f32": (C.c_int, [_P, _P, _P, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, _P, _P, _P]),
    "fvb_bce_loss_f32": (C.c_int, [_P, C.c_int64, C.c_int, _P, _P, C.c_int, _P, C.c_int, _P, _P, _P]),
    "fvb_nms_segmented_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int]),
    "fvb_nms_segmented_f32": (C.c_int, [_P, _P, _P, C.c_int, C.c_int64, C.c_double, C.c_int, _P, _P, _P, _P]),
    "fvb_yolo_nms_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "fvb_yolo_nms_f32": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_float, C.c_double, C.c_int, C.c_int, C.c_float,
                                   _P, _P, C.c_int, _P, _P, _P, _P, _P, _P, _P]),
    "fvb_yolo_nms_after_decode_f32": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_float, C.c_double, C.c_int, C.c_int, C.c_float,
                                                _P, _P, C.c_int, _P, _P, _P, _P, _P, _P, C.c_int, _P, _P]),
    "fvb_rpn_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "fvb_rpn_proposals_f32": (C.c_int, [_P, _P, C.POINTER(C.c_float), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_double, _P, _P, _P, _P, _P]),
    "fvb_yolov3_loss_workspace_bytes": (C.c_size_t, [C.POINTER(Geom), C.c_int64]),
    "fvb_yolov3_loss_f32": (C.c_int, [C.POINTER(Geom), C.POINTER(_P), _P, C.c_int64, C.c_float, C.c_float, C.c_float,
                                      _P, _P, _P, _P, _P]),
    "fvb_yolov3_loss_match_f32": (C.c_int, [C.POINTER(Geom), C.POINTER(_P), _P, C.c_int64, _P, _P]),
    "fvb_yolov3_loss_dense_f32": (C.c_int, [C.POINTER(Geom), C.POINTER(_P), _P, C.c_int64, C.c_float, C.c_float, C.c_float,
                                            _P, C.c_int, _P, _P, _P, _P]),
    "fvb_yolov3_loss_match_dense_f32": (C.c_int, [C.POINTER(Geom), C.POINTER(_P), _P, C.c_int64, C.c_int, _P, _P]),
    "fvb_yolov3_loss_finish_f32": (C.c_int, [C.POINTER(Geom), C.c_int64, C.c_float, C.c_float, C.c_float, _P, _P, _P, _P, _P]),
    "fvb_yolov3_saved_conf_floats": (C.c_int64, [C.POINTER(Geom)]),
    "fvb_yolov3_loss_train_f32": (C.c_int, [C.POINTER(Geom), C.POINTER(_P), _P, C.c_int64, C.c_float, C.c_float, C.c_float,
                                            _P, _P, _P, _P, _P, _P]),
    "fvb_yolov3_loss_combine_f32": (C.c_int, [C.POINTER(Geom), C.c_int64, _P, C.c_float, C.c_float, C.c_float, _P, _P]),
    "fvb_peer_buffer_bytes": (C.c_size_t, []),
    "fvb_yolov3_loss_peer_combine_f32": (C.c_int, [C.POINTER(Geom), C.c_int64, _P, _P, C.c_int, C.c_int, C.c_float, C.c_float,
                                                   C.c_float, _P, _P, _P, _P]),
    "fvb_yolov3_build_target_f32": (C.c_int, [C.POINTER(Geom), C.c_int, _P, C.c_int64, C.c_int, _P, _P, _P, _P, _P, _P,
                                              _P, _P, _P]),
    "fvb_yolov3_loss_backward_workspace_bytes": (C.c_size_t, [C.POINTER(Geom), C.c_int64]),
    "fvb_yolov3_loss_backward_f32": (C.c_int, [C.POINTER(Geom), C.POINTER(_P), _P, C.c_int64, C.c_float, C.c_float, C.c_float,
                                               C.c_int64, _P, _P, _P, C.POINTER(_P), _P, _P]),
    "fvb_iou_loss_backward_f32": (C.c_int, [_P, _P, _P, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, _P, _P, _P,
                                            _P, _P]),
    "fvb_bce_loss_backward_f32": (C.c_int, [_P, C.c_int64, C.c_int, _P, _P, C.c_int, _P, C.c_int, _P, _P, _P]),
    "fvb_demo_loss_workspace_bytes": (C.c_size_t, [C.POINTER(Geom), C.c_int64]),
    "fvb_demo_loss_mask_bytes": (C.c_int64, [C.POINTER(Geom)]),
    "fvb_demo_loss_f32": (C.c_int, [C.POINTER(Geom), C.POINTER(_P), _P, C.c_int64, C.c_int, _P, _P, _P, _P, _P]),
    "fvb_demo_loss_combine_f32": (C.c_int, [C.POINTER(Geom), _P, C.c_int, _P, _P]),
    "fvb_demo_loss_backward_f32": (C.c_int, [C.POINTER(Geom), C.POINTER(_P), _P, C.c_int64, C.c_int, _P, _P, _P,
                                             C.POINTER(_P), _P, _P]),
    "fvb_map_match_workspace_bytes": (C.c_size_t, [C.c_int64]),
    "fvb_map_match_f32": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int64, C.POINTER(C.c_double), C.c_int, _P, _P, _P]),
    "fvb_demo_boxes_postprocess_f32": (C.c_int, [_P, C.c_int64, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                                                 C.c_float, _P]),
    "fvb_map_ap_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int, C.c_int]),
    "fvb_map_ap_f64": (C.c_int, [_P, _P, C.c_int64, _P, C.c_int64, C.c_int, C.c_int, _P, _P, _P, _P]),
    "fvb_kmeans_workspace_bytes": (C.c_size_t, [C.c_int]),
    "fvb_kmeans_step_f32": (C.c_int, [_P, C.c_int64, _P, C.c_int, C.c_float, _P, _P, _P, _P]),
    "fvb_val_evidence_f32": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, _P, C.c_int64, C.c_float, C.c_float, _P, _P, _P, _P, _P]),
    "fvb_debug_set_nms_trace": (None, [_P]),
    "fvb_debug_reload_knobs": (None, []),
}
ABI_VERSION = 2

EXPORTS = tuple(_SIGS)

_lock = threading.Lock()
_lib = None


def load():
    """Return the loaded library; raises RuntimeError when libfvb200.so has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "fastvision_b200: %s is missing -- build it with `python -m fastvision_b200._build` "
                "(nvcc, sm_100a).  There is no CPU fallback." % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(lib, name)  # AttributeError if the .so is stale: loud by design
            fn.restype = res
            fn.argtypes = args
        if lib.fvb_abi_version() != ABI_VERSION:
            raise RuntimeError("libfvb200.so ABI version %d, expected %d" % (lib.fvb_abi_version(), ABI_VERSION))
        _lib = lib
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = load().fvb_last_error().decode("utf-8", "replace")
        exc = ValueError if rc == -1 else RuntimeError
        raise exc("libfvb200 %s failed (%d): %s" % (what, rc, msg))


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def dptr(t):
    """Device pointer of a CUDA tensor (or NULL for None)."""
    if t is None:
        return C.c_void_p(0)
    return C.c_void_p(t.data_ptr())


def require_cuda(t, name, dtype=torch.float32):
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor, got %s" % (name, type(t).__name__))
    if not t.is_cuda:
        raise RuntimeError("%s is on %s: fastvision_b200 runs on CUDA only (no CPU fallback)" % (name, t.device))
    if dtype is not None and t.dtype != dtype:
        raise TypeError("%s must be %s, got %s" % (name, dtype, t.dtype))
    return t if t.is_contiguous() else t.contiguous()


def make_geom(batch, channels, heights, widths, strides, anchors_per_level, head_layout="bahwk"):
    """anchors_per_level: list (per level) of [A,2]-shaped (w,h) pixel anchors (tensor / list)."""
    g = Geom()
    levels = len(heights)
    if levels > MAX_LEVELS:
        raise ValueError("at most %d levels" % MAX_LEVELS)
    g.levels, g.batch, g.channels = levels, int(batch), int(channels)
    na = None
    for l in range(levels):
        a = anchors_per_level[l]
        a = a.detach().cpu().reshape(-1, 2).tolist() if isinstance(a, torch.Tensor) else [list(x) for x in a]
        if na is None:
            na = len(a)
        if len(a) != na:
            raise ValueError("every level must have the same number of anchors")
        if na > MAX_ANCHORS:
            raise ValueError("at most %d anchors per level" % MAX_ANCHORS)
        g.height[l], g.width[l] = int(heights[l]), int(widths[l])
        g.stride[l] = float(strides[l])
        for i, (w, h) in enumerate(a):
            g.anchor_w[l][i] = float(w)
            g.anchor_h[l][i] = float(h)
    g.anchors = na or 0
    g.head_layout = HEAD_LAYOUTS[head_layout]
    return g


def head_ptrs(heads):
    arr = (C.c_void_p * len(heads))()
    for i, h in enumerate(heads):
        arr[i] = h.data_ptr()
    return arr


class Workspaces:
    """Scratch buffers owned by ONE object (a ValStep, a Yolov3Loss, a DecodeContext ...), one per tag and device.

    A buffer that has to grow is RETIRED, not freed: a captured CUDA graph has the old pointer baked in and another
    stream may still be using it, so handing the block back to the caching allocator would let a later tensor alias
    memory that a ``graph.replay()`` still writes.  Retired buffers live as long as their owner; growth is geometric
    (x1.5), so the memory held is bounded by ~3x the largest request.
    """

    def __init__(self):
        self._bufs = {}
        self._retired = []

    def get(self, tag, nbytes, device, zero=False):
        key = (tag, device.index if device.index is not None else torch.cuda.current_device())
        buf = self._bufs.get(key)
        nbytes = max(int(nbytes), 256)
        if buf is None or buf.numel() < nbytes:
            if buf is not None:
                self._retired.append(buf)
                nbytes = max(nbytes, buf.numel() * 3 // 2)
            buf = (torch.zeros if zero else torch.empty)(nbytes, dtype=torch.uint8, device=device)
            self._bufs[key] = buf
        return buf


_shared = Workspaces()


def workspace(nbytes, device, tag="default"):
    """Process-wide scratch for the one-shot API calls (``non_max_suppression(x)``, ``cal_iou_batch`` ...): at least
    ``nbytes`` of uint8 on ``device`` per tag.  Objects that capture CUDA graphs or run on their own streams (ValStep,
    Yolov3Loss) own a private ``Workspaces`` instead of sharing these."""
    return _shared.get(tag, nbytes, device)
