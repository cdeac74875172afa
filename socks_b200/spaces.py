"""Box set used for constraint/target tubes.

The reference takes `gym.spaces.Box` objects and reads only `.low` / `.high`
(`gym_socks/utils/__init__.py:42-47`).  gym is not a dependency here; any object exposing
`.low` and `.high` arrays works with this package, and this class is a minimal one with
gym's defaults (float32 bounds: the float32 rounding of the bounds is part of the
reference's semantics).
"""

import numpy as np

__all__ = ["Box"]


class Box:
    def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
        self.dtype = np.dtype(dtype)
        if shape is None:
            shape = np.broadcast(np.asarray(low), np.asarray(high)).shape
        self.shape = tuple(shape)
        self.low = np.broadcast_to(np.asarray(low, dtype=self.dtype), self.shape).copy()
        self.high = np.broadcast_to(np.asarray(high, dtype=self.dtype), self.shape).copy()

    def contains(self, x):
        x = np.asarray(x)
        return bool(x.shape == self.shape and np.all(x >= self.low) and np.all(x <= self.high))

    def __repr__(self):
        return f"Box({self.low}, {self.high}, {self.shape}, {self.dtype})"
