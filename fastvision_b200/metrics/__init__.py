"""Drop-in for ``fastvision.metrics`` (the mAP part; metrics/__init__.py:1-2)."""
from .map import CalculateMAP

__all__ = ["CalculateMAP"]
