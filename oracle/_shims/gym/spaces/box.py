"""`gym.spaces.box` stand-in: Box + get_inf (test infrastructure only)."""
import numpy as np

from .space import Space


def get_inf(dtype, sign):
    """Largest representable magnitude of `dtype` with the given sign ('+'/'-')."""
    dt = np.dtype(dtype)
    if dt.kind == 'f':
        return np.inf if sign == '+' else -np.inf
    if dt.kind in 'iu':
        info = np.iinfo(dt)
        return info.max - 2 if sign == '+' else info.min + 2
    raise ValueError(f"Unknown dtype {dtype} for infinite bounds")


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
        dtype = np.dtype(dtype)
        if shape is None:
            if np.isscalar(low) and np.isscalar(high):
                shape = (1,)
            else:
                shape = np.broadcast(np.asarray(low), np.asarray(high)).shape
        shape = tuple(int(s) for s in shape)
        self.low = np.broadcast_to(np.asarray(low), shape).astype(dtype)
        self.high = np.broadcast_to(np.asarray(high), shape).astype(dtype)
        self.bounded_below = np.broadcast_to(np.asarray(low), shape) > -np.inf
        self.bounded_above = np.broadcast_to(np.asarray(high), shape) < np.inf
        super().__init__(shape, dtype, seed)

    def is_bounded(self, manner="both"):
        below = bool(np.all(self.bounded_below))
        above = bool(np.all(self.bounded_above))
        if manner == "both":
            return below and above
        if manner == "below":
            return below
        if manner == "above":
            return above
        raise ValueError("manner is not in {'below', 'above', 'both'}")

    def sample(self):
        if self.dtype.kind == 'f':
            return self.np_random.uniform(self.low, self.high).astype(self.dtype)
        return self.np_random.integers(self.low, self.high + 1).astype(self.dtype)

    def contains(self, x):
        if not isinstance(x, np.ndarray):
            try:
                x = np.asarray(x, dtype=self.dtype)
            except (ValueError, TypeError):
                return False
        return bool(
            np.can_cast(x.dtype, self.dtype) and x.shape == self.shape
            and np.all(x >= self.low) and np.all(x <= self.high)
        )

    def __eq__(self, other):
        return isinstance(other, Box) and self.shape == other.shape and \
            np.allclose(self.low, other.low) and np.allclose(self.high, other.high)

    def __repr__(self):
        return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"
