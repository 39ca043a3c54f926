"""Drop-in for demos/yolov3_huaweiShip/utils/lossv3.py: ``from utils.lossv3 import ComputeLoss`` (train.py:16)."""
from ....loss.demo_loss import ComputeLoss  # noqa: F401
