"""Drop-in for ``fastvision.loss`` (loss/__init__.py:1-3)."""
from .yolov3_loss import Yolov3Loss
from .classification_loss import BiCrossEntropyLoss
from .iou_loss import IOULoss, GIOULoss, DIOULoss, CIOULoss
from .demo_loss import ComputeLoss, ComputeLossU

__all__ = ["Yolov3Loss", "BiCrossEntropyLoss", "IOULoss", "GIOULoss", "DIOULoss", "CIOULoss", "ComputeLoss", "ComputeLossU"]
