"""The tiny YOLO configuration the golden fixtures were generated with (oracle/make_golden.py: SMALL)."""
from fastvision_b200 import synth

SMALL = synth.YoloConfig("tiny", 64, 4, [[40, 30], [50, 60], [30, 50], [20, 24], [16, 10], [12, 22], [4, 6], [8, 5], [7, 9]],
                         labels_per_img=4.0, max_labels=9)
