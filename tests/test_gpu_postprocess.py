"""GPU parity of the demos' postProcess drop-ins (decode of the raw conv outputs, un-letterbox / clamp / 5-px filter /
xyxy, class-aware NMS) against golden vectors recorded from the reference's own function and against the oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import oracle
from conftest import T
from gpu_util import cuda, close
from small_cfg import SMALL
from fastvision_b200 import synth
from fastvision_b200.demos.yolov3_u.inference import postProcess as post_u
from fastvision_b200.demos.yolov3_huaweiShip.inference import postProcess as post_ship


@pytest.mark.parametrize("demo,fn", [("yolov3_u", post_u), ("yolov3_huaweiShip", post_ship)])
def test_postprocess_golden(golden_postprocess, demo, fn):
    g = golden_postprocess
    heads = [cuda(g["head%d" % i]) for i in range(3)]
    anchors = [cuda(g["anchors%d" % i]) for i in range(3)]
    ct, it, rr, pl, pt, ow, oh = g["args"].tolist()
    keep = [h.clone() for h in heads]
    s, c, b = fn(heads, SMALL.strides, anchors, ct, it, rr, int(pl), int(pt), int(ow), int(oh))
    assert s.shape == g[demo + "_scores"].shape and b.shape == g[demo + "_boxes"].shape
    close(s, g[demo + "_scores"]); close(b, g[demo + "_boxes"], atol=2e-5)
    assert np.array_equal(c.cpu().numpy(), g[demo + "_cats"])
    for h, k in zip(heads, keep):
        assert torch.equal(h, k)                 # the inputs are not modified (the reference writes into its own view-copy)


@pytest.mark.parametrize("form,fn", [("v5", post_u), ("v3", post_ship)])
def test_postprocess_vs_oracle_608(form, fn):
    cfg = synth.SHIP608
    g = synth.make_generator(6)
    labels = synth.make_labels(cfg, 1, g)
    heads = [h.permute(0, 1, 4, 2, 3).reshape(1, -1, h.size(2), h.size(3)).contiguous() for h in synth.make_heads(cfg, 1, labels, g)]
    anchors = [a.reshape(-1, 2) / s for a, s in zip(cfg.anchors_levels(), cfg.strides)]
    args = (0.25, 0.45, 0.76, 0, 76, 800, 600)
    ws, wc, wb, rows = oracle.postprocess.post_process(heads, cfg.strides, anchors, *args, form=form)
    s, c, b = fn([h.cuda() for h in heads], cfg.strides, [a.cuda() for a in anchors], *args)
    if s.shape == ws.shape and np.array_equal(c.cpu().numpy(), wc.numpy()):
        close(s, ws); close(b, wb, atol=1e-4)
    else:
        # a keep set may only differ through a pair whose IoU sits within 1e-6 of the threshold (north-star slack)
        from oracle.nms import nms_greedy
        cand = rows[rows[:, 4] > args[0]]
        cat = cand[:, 5:].argmax(1).float()
        _, margin = nms_greedy(cand[:, :4] + cat[:, None] * 4096, cand[:, 4], args[1], return_iou_margin=True)
        assert margin < 1e-6, (s.shape, ws.shape, margin)


def test_postprocess_nothing_survives(golden_postprocess):
    g = golden_postprocess
    heads = [cuda(g["head%d" % i]) for i in range(3)]
    anchors = [cuda(g["anchors%d" % i]) for i in range(3)]
    s, c, b = post_u(heads, SMALL.strides, anchors, 1.1, 0.4, 0.8, 3, 5, 72, 66)         # no objectness exceeds 1.1
    assert s.shape == (0, 1) and c.shape == (0, 1) and b.shape == (0, 4)
    s, c, b = post_ship(heads, SMALL.strides, anchors, 0.2, 0.4, 0.8, 3, 5, 4, 4)       # a 4x4 original image: every box <= 5 px
    assert s.shape == (0, 1) and b.shape == (0, 4)
    with pytest.raises(ValueError):
        post_u([torch.cat([h, h]) for h in heads], SMALL.strides, anchors, 0.2, 0.4, 0.8, 3, 5, 72, 66)
