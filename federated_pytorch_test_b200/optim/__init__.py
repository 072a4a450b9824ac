"""Optimizers: ``LBFGSNew`` (reference-compatible) and the fused block Adam."""
from .lbfgsnew import LBFGSNew

__all__ = ["LBFGSNew"]
