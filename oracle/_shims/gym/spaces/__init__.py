"""`gym.spaces` stand-in (test infrastructure only; see ../__init__.py)."""
from collections import OrderedDict

import numpy as np

from .space import Space
from .box import Box
from . import box  # noqa: F401


class Discrete(Space):
    def __init__(self, n, seed=None, start=0):
        assert n > 0
        self.n = int(n)
        self.start = int(start)
        super().__init__((), np.int64, seed)

    def sample(self):
        return int(self.start + self.np_random.integers(self.n))

    def contains(self, x):
        if isinstance(x, (int, np.integer)):
            v = int(x)
        elif isinstance(x, np.ndarray) and x.dtype.kind in 'iu' and x.shape == ():
            v = int(x)
        else:
            return False
        return self.start <= v < self.start + self.n

    def __eq__(self, other):
        return isinstance(other, Discrete) and self.n == other.n and self.start == other.start

    def __repr__(self):
        return f"Discrete({self.n})"


class MultiBinary(Space):
    def __init__(self, n, seed=None):
        self.n = n
        shape = (n,) if isinstance(n, (int, np.integer)) else tuple(n)
        super().__init__(shape, np.int8, seed)

    def sample(self):
        return self.np_random.integers(0, 2, size=self.shape).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return bool(x.shape == self.shape and np.all((x == 0) | (x == 1)))

    def __eq__(self, other):
        return isinstance(other, MultiBinary) and self.shape == other.shape


class MultiDiscrete(Space):
    def __init__(self, nvec, dtype=np.int64, seed=None):
        self.nvec = np.array(nvec, dtype=dtype, copy=True)
        assert (self.nvec > 0).all()
        super().__init__(self.nvec.shape, dtype, seed)

    def sample(self):
        return (self.np_random.random(self.nvec.shape) * self.nvec).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return bool(x.shape == self.shape and x.dtype.kind in 'iu'
                    and np.all(x >= 0) and np.all(x < self.nvec))

    def __len__(self):
        return len(self.nvec)

    def __eq__(self, other):
        return isinstance(other, MultiDiscrete) and np.array_equal(self.nvec, other.nvec)


class Dict(Space):
    def __init__(self, spaces=None, seed=None, **spaces_kwargs):
        if spaces is None:
            spaces = spaces_kwargs
        if isinstance(spaces, Dict):
            spaces = spaces.spaces
        self.spaces = OrderedDict(spaces)
        for s in self.spaces.values():
            assert isinstance(s, Space), "Values of the dict should be instances of gym.Space"
        super().__init__(None, None, seed)

    def seed(self, seed=None):
        out = []
        for i, s in enumerate(self.spaces.values()):
            out += s.seed(None if seed is None else seed + i)
        return out

    def sample(self):
        return OrderedDict((k, s.sample()) for k, s in self.spaces.items())

    def contains(self, x):
        if not isinstance(x, dict) or len(x) != len(self.spaces):
            return False
        for k, s in self.spaces.items():
            if k not in x or not s.contains(x[k]):
                return False
        return True

    def __getitem__(self, key):
        return self.spaces[key]

    def __setitem__(self, key, value):
        self.spaces[key] = value

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
        return "Dict(" + ", ".join(f"{k}:{s}" for k, s in self.spaces.items()) + ")"


class Tuple(Space):
    def __init__(self, spaces, seed=None):
        self.spaces = tuple(spaces)
        super().__init__(None, None, seed)

    def seed(self, seed=None):
        out = []
        for i, s in enumerate(self.spaces):
            out += s.seed(None if seed is None else seed + i)
        return out

    def sample(self):
        return tuple(s.sample() for s in self.spaces)

    def contains(self, x):
        if isinstance(x, (list, np.ndarray)):
            x = tuple(x)
        return isinstance(x, tuple) and len(x) == len(self.spaces) and \
            all(s.contains(p) for s, p in zip(self.spaces, x))

    def __getitem__(self, i):
        return self.spaces[i]

    def __len__(self):
        return len(self.spaces)

    def __eq__(self, other):
        return isinstance(other, Tuple) and self.spaces == other.spaces
