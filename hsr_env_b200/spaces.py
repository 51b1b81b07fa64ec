"""Minimal stand-in for the two gym names the reference uses (``gym.Space`` / ``gym.spaces.Box``); gym is not
installed here.  Only what /root/reference/hsr and /root/reference/rl_utils/argparse.py:57-73 touch."""
from __future__ import annotations

import numpy as np


class Space:
    def sample(self):
        raise NotImplementedError


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        low = np.asarray(low, dtype=dtype)
        high = np.asarray(high, dtype=dtype)
        if shape is not None:
            low = np.broadcast_to(low, shape).copy()
            high = np.broadcast_to(high, shape).copy()
        assert low.shape == high.shape
        self.low, self.high, self.dtype, self.shape = low, high, np.dtype(dtype), low.shape
        self._rng = np.random.default_rng()

    def seed(self, seed=None):
        self._rng = np.random.default_rng(seed)
        return [seed]

    def sample(self):
        lo = np.where(np.isfinite(self.low), self.low, -1e6)
        hi = np.where(np.isfinite(self.high), self.high, 1e6)
        return self._rng.uniform(lo, hi).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    def __repr__(self):
        return f"Box({self.low.tolist()}, {self.high.tolist()}, {self.dtype})"

    def __eq__(self, other):
        return isinstance(other, Box) and np.array_equal(self.low, other.low) and np.array_equal(self.high, other.high)


def space_to_size(space) -> int:
    """rl_utils.gym.space_to_size for Box (used at /root/reference/hsr/control.py:49,70)."""
    return int(np.prod(space.shape))
