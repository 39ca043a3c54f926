"""fastvision_b200 -- B200-native (sm_100a) drop-in for fastvision's detection hot path.

Only the path decode -> NMS -> target assignment + loss -> mAP matching lives
here: hand-written CUDA kernels behind a C ABI (``include/fvb200.h``, ``csrc/``) and a Python host
side that mirrors the reference's module layout and call signatures:

    fastvision.detection.tools   -> fastvision_b200.detection.tools
    fastvision.detection.models  -> fastvision_b200.detection.models
    fastvision.loss              -> fastvision_b200.loss
    fastvision.metrics           -> fastvision_b200.metrics

There is no CPU fallback: calling any op without ``libfvb200.so`` + a CUDA device raises.
"""
__version__ = "0.1.0"
