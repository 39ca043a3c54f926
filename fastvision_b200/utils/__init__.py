"""Drop-in for the hot-path caller of ``fastvision.utils``: ``Fit`` (utils/fit.py)."""
from .fit import Fit

__all__ = ["Fit"]
