"""Drop-in for demos/yolov3_u/utils/lossv3.py: ``from utils.lossv3 import ComputeLoss`` (train.py:15)."""
from ....loss.demo_loss import ComputeLossU as ComputeLoss  # noqa: F401
