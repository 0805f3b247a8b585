"""`gymnasium.spaces` stand-in: shape/nvec holders only (see ../__init__.py)."""
import numpy as np

from .multi_discrete import MultiDiscrete  # noqa: F401


class Box:
    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), dtype
