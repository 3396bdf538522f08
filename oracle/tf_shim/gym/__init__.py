"""TEST INFRASTRUCTURE ONLY -- minimal stand-in for the `gym` names the reference
imports (gym.spaces.Box, gym.core.Env, gym.logger).  See ../tensorflow/__init__.py."""
import sys as _sys
import types as _types

import numpy as _np


class Env:
    pass


class Box:
    """gym.spaces.Box(low, high, shape=None, dtype=float32) -- bounds broadcast to `shape`."""

    def __init__(self, low, high, shape=None, dtype=_np.float32):
        if shape is not None:
            shape = tuple(int(s) for s in (shape.numpy().tolist() if hasattr(shape, "numpy") else shape))
            low = _np.full(shape, low, dtype=dtype) if _np.isscalar(low) else _np.broadcast_to(_np.asarray(low, dtype=dtype), shape).copy()
            high = _np.full(shape, high, dtype=dtype) if _np.isscalar(high) else _np.broadcast_to(_np.asarray(high, dtype=dtype), shape).copy()
        else:
            low = _np.asarray(low, dtype=dtype)
            high = _np.asarray(high, dtype=dtype)
        self.low, self.high = low, high
        self.shape = low.shape
        self.dtype = _np.dtype(dtype)
        self.bounded_below = -_np.inf < low
        self.bounded_above = _np.inf > high

    def is_bounded(self, manner="both"):
        below = bool(_np.all(self.bounded_below))
        above = bool(_np.all(self.bounded_above))
        if manner == "both":
            return below and above
        return below if manner == "below" else above


core = _types.ModuleType("gym.core")
core.Env = Env
spaces = _types.ModuleType("gym.spaces")
spaces.Box = Box
logger = _types.ModuleType("gym.logger")
logger.ERROR = 40
logger.set_level = lambda level: None
_sys.modules["gym.core"] = core
_sys.modules["gym.spaces"] = spaces
_sys.modules["gym.logger"] = logger
