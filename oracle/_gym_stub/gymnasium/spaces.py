"""`Box` / `Discrete` that only remember their arguments (see package docstring)."""
import numpy as np


class Box:
    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.low, self.high, self.shape, self.dtype = low, high, shape, dtype


class Discrete:
    def __init__(self, n):
        self.n = n
        self._rng = np.random.default_rng()

    def sample(self):
        return int(self._rng.integers(self.n))
