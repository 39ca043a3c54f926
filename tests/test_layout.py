"""Host-side checks that need no GPU: the C-ABI library loads and exports every declared symbol, the product
package never touches the oracle, and the drop-in modules expose the reference's names."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from fastvision_b200 import _build, _lib
    _build.build()
    header = open(os.path.join(ROOT, "include", "fvb200.h")).read()
    declared = set(re.findall(r"\b(fvb_[a-z0-9_]+)\s*\(", header))
    declared -= {"fvb_yolo_geom"}
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), "libfvb200.so does not export %s" % name
    assert declared == set(_lib.EXPORTS), (declared ^ set(_lib.EXPORTS))
    assert _lib.load().fvb_abi_version() == 2


def test_product_never_imports_the_oracle():
    # the package AND the timing / profiling tools: only tests/, __graft_entry__.smoke() and bench.py's CPU legs may touch oracle/
    bad = []
    for dirpath, _, files in list(os.walk(os.path.join(ROOT, "fastvision_b200"))) + list(os.walk(os.path.join(ROOT, "tools"))):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", text, re.M) or "torchvision" in text and f.endswith(".py") and "import torchvision" in text:
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_drop_in_names_match_the_reference():
    import fastvision_b200.detection.tools as t
    import fastvision_b200.loss as l
    import fastvision_b200.metrics as m
    from fastvision_b200.detection.models import Yolov3, yolov3  # noqa: F401
    for name in ["xywh2xyxy", "xyxy2xywh", "xyxy2xywhn", "grid", "cal_iou", "cal_iou_batch", "xyxy_iou", "xywh_iou", "wh_iou",
                 "xyxy_iou_batch", "xywh_iou_batch", "wh_iou_batch", "GIOU", "GIOU_batch", "DIOU", "DIOU_batch", "CIOU",
                 "CIOU_batch", "non_max_suppression"]:
        assert callable(getattr(t, name)), name
    for name in ["Yolov3Loss", "BiCrossEntropyLoss", "IOULoss", "GIOULoss", "DIOULoss", "CIOULoss"]:
        assert callable(getattr(l, name)), name
    assert callable(m.CalculateMAP)
    from fastvision_b200.utils import Fit
    from fastvision_b200.detection.tools import KMeans, AnchorGenerator  # noqa: F401
    from fastvision_b200.loss import ComputeLoss, ComputeLossU  # noqa: F401
    import inspect
    assert list(inspect.signature(Fit.__init__).parameters)[:12] == ['self', 'model', 'device', 'optimizer', 'scheduler', 'loss', 'end_epoch',
                                                                     'start_epoch', 'train_loader', 'val_loader', 'test_loader', 'data_dict']


def test_cpu_tensors_are_rejected_loudly():
    import torch
    import fastvision_b200.detection.tools as t
    with pytest.raises(RuntimeError, match="CUDA only"):
        t.xywh2xyxy(torch.zeros(3, 4))
    with pytest.raises(RuntimeError, match="CUDA only"):
        t.non_max_suppression(torch.zeros(10, 9))


def test_grid_matches_oracle():
    import numpy as np
    from fastvision_b200.detection.tools import grid, offset
    from oracle.boxes import grid as ogrid
    for mode in ("xy", "yx"):
        assert np.array_equal(grid(3, 5, mode, "torch").numpy(), ogrid(3, 5, mode, "torch").numpy())
        assert np.array_equal(grid(3, 5, mode, "numpy"), ogrid(3, 5, mode, "numpy"))
    assert np.array_equal(offset(2, 3), ogrid(2, 3, "yx", "numpy"))


def test_geometry_helpers_without_gpu():
    from fastvision_b200 import _lib, synth
    lib = _lib.load()
    cfg = synth.COCO416
    g = _lib.make_geom(8, cfg.k, cfg.feat, cfg.feat, cfg.strides, cfg.anchors_levels())
    assert lib.fvb_yolo_rows_per_image(g) == 10647
    assert lib.fvb_yolo_bitmap_words(g) == 333
    assert lib.fvb_yolo_decode_partials(g) == 8 * (32 + 127 + 507)     # one objectness partial per 16-row tile
    bad = _lib.make_geom(8, 3, cfg.feat, cfg.feat, cfg.strides, cfg.anchors_levels())
    assert lib.fvb_yolo_rows_per_image(bad) == -1
    assert b"channels" in lib.fvb_last_error()


def test_header_is_plain_c_and_links_from_gcc(tmp_path):
    """include/fvb200.h compiles as C (not only C++) and a gcc-built program links against libfvb200.so and calls it."""
    import shutil
    import subprocess
    from fastvision_b200 import _build, _lib
    if shutil.which("gcc") is None:
        pytest.skip("gcc not installed")
    _build.build()
    exe = str(tmp_path / "abi_smoke")
    libdir = os.path.dirname(_lib.LIB_PATH)
    cuda_inc = "/usr/local/cuda/include"
    cmd = ["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), "-I", cuda_inc,
           os.path.join(ROOT, "tests", "abi_smoke.c"), "-o", exe, "-L", libdir, "-l:libfvb200.so", "-Wl,-rpath," + libdir]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "abi ok" in r.stdout, (r.returncode, r.stdout, r.stderr)
