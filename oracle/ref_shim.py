"""Import shim for the REAL reference (ielym/fastvision) -- container-only, test infrastructure.

TEST INFRASTRUCTURE ONLY.  Nothing under ``fastvision_b200/`` may import this.

``/root/reference`` exists only in the build container, never on the GPU box, so this
module is used solely by ``oracle/make_golden.py`` (to produce ``tests/golden/*.npz``)
and by ``tests/test_oracle_vs_reference.py`` (skipped when the reference is absent).
It makes the reference importable *unmodified* (SURVEY.md Appendix B):

  * ``fastvision`` alias: a temp dir holding a symlink ``fastvision -> /root/reference``
    is put on ``sys.path`` (all intra-repo imports are absolute, e.g. loss/iou_loss.py:3);
  * stub ``matplotlib`` / ``matplotlib.pyplot`` (detection/tools/ANCHOR.py:2 imports it);
  * ``np.float`` / ``np.long`` aliases (metrics/map.py:34,81,106,127 need numpy<1.24);
  * ``Yolov3Loss.build_target`` is re-bound to a body identical to
    loss/yolov3_loss.py:75-124 except the clamp bounds are cast to int
    (lines 116-117 crash on torch>=2: Long.clamp_(Float tensor)).
"""
import os
import sys
import tempfile
import types

REFERENCE_ROOT = os.environ.get("FASTVISION_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "detection", "tools", "IOU.py"))


_loaded = None


def load():
    """Returns a namespace with the reference's hot-path callables."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    import numpy as np
    import torch

    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        plt.rcParams = {}
        mpl.pyplot = plt
        mpl.rcParams = {}
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
    if not hasattr(np, "float"):
        np.float = float
    if not hasattr(np, "long"):
        np.long = np.int64

    alias_dir = tempfile.mkdtemp(prefix="fv_ref_alias_")
    os.symlink(REFERENCE_ROOT, os.path.join(alias_dir, "fastvision"))
    sys.path.insert(0, alias_dir)

    import fastvision.detection.tools as tools  # noqa
    import fastvision.loss as loss  # noqa
    import fastvision.metrics as metrics  # noqa

    ns = types.SimpleNamespace()
    ns.tools = tools
    ns.loss = loss
    ns.metrics = metrics

    def _build_target(self, y_pred, y_true):
        # same statements as loss/yolov3_loss.py:75-124; only :116-117 differ (int() cast)
        gt_locations, gt_categories, gt_xywh, matched_anchors = [], [], [], []
        for layer_idx, pre in enumerate(y_pred):
            anchors = self.anchor_levels[layer_idx].squeeze()
            anchors = anchors / self.backbone_stride_levels[layer_idx]
            num_anchors = anchors.size(0)
            feature_whwh = torch.tensor(pre.size()).to(pre)[[3, 2, 3, 2]]
            target = y_true.clone()
            target[:, 2:] = y_true[:, 2:] * feature_whwh
            num_targets = target.size(0)
            wh_similarity = target[:, None, 4:] / anchors
            similarity_mask = torch.max(wh_similarity, 1 / wh_similarity).max(2)[0] < 4
            target_like_anchors = target.unsqueeze(1).repeat(1, num_anchors, 1)
            anchors_idxs = torch.arange(num_anchors).unsqueeze(0).repeat(num_targets, 1).to(pre)
            target_with_anchors = torch.cat([target_like_anchors, anchors_idxs[:, :, None]], dim=2)
            m = target_with_anchors[similarity_mask, ...]
            b = m[:, 0].long()
            c = m[:, 1].long()
            xy = m[:, 2:4]
            wh = m[:, 4:6]
            a = m[:, 6].long()
            gxy = torch.floor(xy).long()
            off = xy - gxy.float()
            gxy[:, 0] = gxy[:, 0].clamp_(0, int(feature_whwh[0]) - 1)
            gxy[:, 1] = gxy[:, 1].clamp_(0, int(feature_whwh[1]) - 1)
            gt_locations.append((b, gxy, a))
            gt_categories.append(c)
            gt_xywh.append(torch.cat([off, wh], dim=1))
            matched_anchors.append(anchors[a])
        return gt_locations, gt_categories, gt_xywh, matched_anchors

    class Yolov3LossPatched(loss.Yolov3Loss):
        build_target = _build_target

    ns.Yolov3Loss = Yolov3LossPatched

    def decode(head_out, anchors_per_level, strides, num_classes):
        # detection/models/yolov3.py:33-53 with offset := grid(h, w, 'yx', 'numpy');
        # the module itself cannot be imported (yolov3.py:4 imports a missing symbol).
        results = []
        for i in range(len(head_out)):
            out = head_out[i]
            bs, num_anchors, height, width, _ = out.size()
            offset_level = torch.tensor(tools.grid(height, width, mode="yx", dtype="numpy")).to(out)
            offset_level = offset_level.expand_as(out[..., 0:2])
            xy = (out[..., 0:2].sigmoid() + offset_level) * strides[i]
            wh = torch.exp(out[..., 2:4]) * anchors_per_level[i].expand_as(out[..., 2:4]).to(out)
            out = torch.cat((xy, wh, out[..., 4:].sigmoid()), -1)
            results.append(out.view(bs, -1, num_classes + 5))
        return torch.cat(results, 1)

    ns.decode = decode

    def load_demo(name, module):
        """Load demos/<name>/utils/<module>.py under a private package name."""
        import importlib.util
        pkg_name = "_fvref_%s_utils" % name
        pkg_dir = os.path.join(REFERENCE_ROOT, "demos", name, "utils")
        if pkg_name not in sys.modules:
            pkg = types.ModuleType(pkg_name)
            pkg.__path__ = [pkg_dir]
            sys.modules[pkg_name] = pkg
        full = pkg_name + "." + module
        if full in sys.modules:
            return sys.modules[full]
        spec = importlib.util.spec_from_file_location(full, os.path.join(pkg_dir, module + ".py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[full] = mod
        spec.loader.exec_module(mod)
        return mod

    ns.load_demo = load_demo

    def load_rpn():
        import importlib.util
        path = os.path.join(REFERENCE_ROOT, "demos", "faster_rcnn", "models", "rpn.py")
        spec = importlib.util.spec_from_file_location("_fvref_rpn", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod

    ns.load_rpn = load_rpn
    _loaded = ns
    return ns
