from .yolov3 import Yolov3, yolov3, yolov3_decode, DecodeContext  # noqa: F401
