"""``Box`` space: gym's when gym / gymnasium is installed, else a minimal stand-in with the
attributes the reference and SB3's ``check_env`` read (low, high, shape, dtype, sample,
contains)."""
from __future__ import annotations

import numpy as np

try:                                    # pragma: no cover - neither package is in the build image
    from gym.spaces import Box          # type: ignore
    GYM_FLAVOUR = "gym"
except Exception:                       # noqa: BLE001
    try:
        from gymnasium.spaces import Box  # type: ignore
        GYM_FLAVOUR = "gymnasium"
    except Exception:                   # noqa: BLE001
        GYM_FLAVOUR = None

        class Box:  # type: ignore
            def __init__(self, low, high, shape=None, dtype=np.float32):
                self.dtype = np.dtype(dtype)
                if shape is None:
                    shape = np.broadcast(np.asarray(low), np.asarray(high)).shape
                self.shape = tuple(int(s) for s in shape)
                self.low = np.broadcast_to(np.asarray(low, dtype=self.dtype), self.shape).copy()
                self.high = np.broadcast_to(np.asarray(high, dtype=self.dtype), self.shape).copy()
                self._rng = np.random.default_rng()

            def seed(self, seed=None):
                self._rng = np.random.default_rng(seed)
                return [seed]

            def sample(self):
                return self._rng.uniform(self.low, self.high).astype(self.dtype)

            def contains(self, x) -> bool:
                x = np.asarray(x)
                return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

            def __repr__(self):
                return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"


def observation_box(n_traffic: int, dtype=np.float64) -> "Box":
    """reference environment.py:18-21: low [0,0,-1,0,0] + [0,-1,-1]*N, high 1."""
    lo = np.array([0, 0, -1, 0, 0] + [0, -1, -1] * n_traffic).astype(dtype)
    hi = np.ones([5 + 3 * n_traffic]).astype(dtype)
    return Box(low=lo, high=hi, dtype=dtype)


def action_box(dtype=np.float64) -> "Box":
    """reference environment.py:27: lateral acceleration scaled to [-1, 1]."""
    return Box(low=-1, high=1, shape=(1,), dtype=dtype)
