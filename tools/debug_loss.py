import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastvision_b200 import synth
from fastvision_b200.pipeline import ValStep, ValPipeline
from fastvision_b200.loss import Yolov3Loss
cfg = synth.COCO416
B = 256
g = synth.make_generator(2)
labels = synth.make_labels(cfg, B, g)
heads = synth.make_heads(cfg, B, labels, g)
dh, dl = [h.cuda() for h in heads], labels.cuda()
anc, st = cfg.anchors_levels(), cfg.strides
class M:
    anchors_per_level = anc
    backbone_strides_per_level = st
lossf = Yolov3Loss(M(), 0.5, 0.05, 1.0, 0.5)
for i in range(3):
    l = lossf(dh, dl)
    print("standalone", l.item(), lossf.partials.cpu().tolist())
step = ValStep(anc, st)
for i in range(4):
    o = step(dh, dl)
    torch.cuda.synchronize()
    print("valstep", o["loss"].item(), o["partials"].cpu().tolist())
replay = step.capture(dh, dl)
for i in range(3):
    replay(); torch.cuda.synchronize()
    print("graph", step.out["loss"].item(), step.out["partials"].cpu()[:, 2].tolist())
pipe = ValPipeline(anc, st)
for i in range(6):
    o = pipe.submit(dh, dl)
    pipe.flush(); torch.cuda.synchronize()
    print("pipe", o["loss"].item(), o["partials"].cpu()[:, 2].tolist())
