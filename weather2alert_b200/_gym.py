"""`gymnasium` when it is installed, otherwise the few names this package needs from it."""
from __future__ import annotations

import numpy as np

try:  # pragma: no cover - depends on the image
    import gymnasium as _gym
    from gymnasium import spaces

    Env = _gym.Env
    VectorEnv = _gym.vector.VectorEnv
    HAVE_GYMNASIUM = True
except ImportError:  # the build image has no gymnasium wheel
    HAVE_GYMNASIUM = False

    class Env:  # noqa: D401 - same surface as gymnasium.Env for what we use
        metadata: dict = {}

        def __init__(self):
            pass

        def close(self):
            pass

    class VectorEnv:  # noqa: D401 - the attributes gymnasium.vector.VectorEnv declares, nothing more
        metadata: dict = {}
        num_envs: int
        closed: bool = False

        def close(self, **kwargs):
            self.closed = True

    class _Spaces:
        class Box:
            def __init__(self, low, high, shape=None, dtype=np.float32):
                self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), dtype

            def __repr__(self):
                return f"Box({self.low}, {self.high}, {self.shape}, {np.dtype(self.dtype).name})"

        class Discrete:
            def __init__(self, n, seed=None):
                self.n = int(n)
                self._rng = np.random.default_rng(seed)

            def sample(self):
                return int(self._rng.integers(self.n))

            def contains(self, x):
                return 0 <= int(x) < self.n

            def __repr__(self):
                return f"Discrete({self.n})"

    spaces = _Spaces()
