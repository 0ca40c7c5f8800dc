"""Space descriptors (metadata only).

The reference attaches `gym.spaces` objects to every agent (actor.py:61-66, observer.py:166-174); the batched
engine emits tensors, so spaces are plain descriptors here -- enough for `null_action in action_space`
checks (agent_based_simulation.py:112-117) and for users that read shapes/bounds.
"""
import numpy as np


class Space:
    def contains(self, x):
        raise NotImplementedError

    def __contains__(self, x):
        return self.contains(x)

    def seed(self, seed=None):
        self._rng = np.random.default_rng(seed)

    @property
    def rng(self):
        if not hasattr(self, '_rng'):
            self.seed(None)
        return self._rng


class Box(Space):
    """Box(low, high, shape, dtype) that also accepts python scalars (abmarl/tools/gym_utils.py:6-24)."""

    def __init__(self, low, high, shape=None, dtype=int):
        self.dtype = np.dtype(dtype)
        if shape is None:
            shape = np.broadcast(np.asarray(low), np.asarray(high)).shape or (1,)
        self.shape = tuple(shape)
        self.low = np.broadcast_to(np.asarray(low), self.shape).astype(self.dtype)
        self.high = np.broadcast_to(np.asarray(high), self.shape).astype(self.dtype)

    def contains(self, x):
        if type(x) is int:
            x = np.array([x], dtype=int)
        elif type(x) is float:
            x = np.array([x], dtype=float)
        elif not isinstance(x, np.ndarray):
            x = np.asarray(x, dtype=self.dtype)
        return bool(np.can_cast(x.dtype, self.dtype) and x.shape == self.shape
                    and np.all(x >= self.low) and np.all(x <= self.high))

    def sample(self):
        if self.dtype.kind == 'f':
            return self.rng.uniform(self.low, self.high).astype(self.dtype)
        return self.rng.integers(self.low, self.high + 1).astype(self.dtype)

    def __eq__(self, other):
        return isinstance(other, Box) and self.shape == other.shape and \
            np.array_equal(self.low, other.low) and np.array_equal(self.high, other.high)

    def __repr__(self):
        return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"


class Discrete(Space):
    def __init__(self, n):
        assert n > 0
        self.n = int(n)
        self.shape = ()
        self.dtype = np.dtype(np.int64)

    def contains(self, x):
        if isinstance(x, (int, np.integer)) or (isinstance(x, np.ndarray) and x.shape == () and x.dtype.kind in 'iu'):
            return 0 <= int(x) < self.n
        return False

    def sample(self):
        return int(self.rng.integers(self.n))

    def __eq__(self, other):
        return isinstance(other, Discrete) and self.n == other.n

    def __repr__(self):
        return f"Discrete({self.n})"


class MultiDiscrete(Space):
    """gym.spaces.MultiDiscrete(nvec): a vector whose j-th entry is in [0, nvec[j])."""

    def __init__(self, nvec):
        self.nvec = np.asarray(nvec, dtype=int)
        self.shape, self.dtype = self.nvec.shape, np.dtype(int)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.nvec.shape and bool(np.all(x >= 0) and np.all(x < self.nvec))

    def sample(self):
        return np.array([self.rng.integers(0, n) for n in self.nvec], dtype=int)

    def __len__(self):
        return len(self.nvec)

    def __eq__(self, other):
        return isinstance(other, MultiDiscrete) and np.array_equal(self.nvec, other.nvec)

    def __repr__(self):
        return f"MultiDiscrete({self.nvec.tolist()})"


class Dict(Space):
    def __init__(self, spaces):
        self.spaces = dict(spaces)

    def contains(self, x):
        return isinstance(x, dict) and x.keys() == self.spaces.keys() and \
            all(s.contains(x[k]) for k, s in self.spaces.items())

    def sample(self):
        return {k: s.sample() for k, s in self.spaces.items()}

    def seed(self, seed=None):
        for i, s in enumerate(self.spaces.values()):
            s.seed(None if seed is None else seed + i)

    def __getitem__(self, k):
        return self.spaces[k]

    def __setitem__(self, k, v):
        self.spaces[k] = v

    def __iter__(self):
        return iter(self.spaces)

    def __len__(self):
        return len(self.spaces)

    def keys(self):
        return self.spaces.keys()

    def values(self):
        return self.spaces.values()

    def items(self):
        return self.spaces.items()

    def __eq__(self, other):
        return isinstance(other, Dict) and self.spaces == other.spaces

    def __repr__(self):
        return "Dict(" + ", ".join(f"{k}: {v}" for k, v in self.spaces.items()) + ")"


def check_space(space, strict=False):
    """abmarl/tools/gym_utils.py:27-55."""
    if isinstance(space, (Box, Discrete, MultiDiscrete)):
        return True
    if isinstance(space, Dict):
        return all(check_space(s) for s in space.spaces.values())
    if not strict and isinstance(space, dict):
        return all(check_space(s) for s in space.values())
    return False


def make_dict(space):
    """abmarl/tools/gym_utils.py:58-71."""
    for key, sub in space.items():
        if isinstance(sub, dict):
            space[key] = make_dict(sub)
    return Dict(space) if type(space) is dict else space
