"""Space containers with the semantics of gym 0.21-0.26 that the reference relies on."""
from collections import OrderedDict
from collections.abc import Mapping

import numpy as np


class Space:
    def __init__(self, shape=None, dtype=None, seed=None):
        self._shape = None if shape is None else tuple(shape)
        self.dtype = None if dtype is None else np.dtype(dtype)
        self._np_random = np.random.default_rng(seed)

    @property
    def shape(self):
        return self._shape

    @shape.setter
    def shape(self, value):
        self._shape = value

    def sample(self):
        raise NotImplementedError

    def contains(self, x):
        raise NotImplementedError

    def __contains__(self, x):
        return self.contains(x)

    def seed(self, seed=None):
        self._np_random = np.random.default_rng(seed)
        return [seed]


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
        dtype = np.dtype(dtype)
        if shape is None:
            if np.isscalar(low) and np.isscalar(high):
                shape = (1,)
            elif np.isscalar(low):
                shape = np.asarray(high).shape
            else:
                shape = np.asarray(low).shape
        shape = tuple(shape)
        low = np.full(shape, low, dtype=dtype) if np.isscalar(low) else np.asarray(low).astype(dtype)
        high = np.full(shape, high, dtype=dtype) if np.isscalar(high) else np.asarray(high).astype(dtype)
        assert low.shape == shape and high.shape == shape, (low.shape, high.shape, shape)
        self.low, self.high = low, high
        super().__init__(shape, dtype, seed)

    def sample(self):
        low = np.where(np.isfinite(self.low), self.low, -1e6)
        high = np.where(np.isfinite(self.high), self.high, 1e6)
        return self._np_random.uniform(low, high, size=self.shape).astype(self.dtype)

    def contains(self, x):
        try:
            x = np.asarray(x, dtype=self.dtype)
        except (TypeError, ValueError):
            return False
        return bool(x.shape == self.shape and np.all(x >= self.low) and np.all(x <= self.high))

    def __eq__(self, other):
        return (isinstance(other, Box) and self.shape == other.shape
                and np.allclose(self.low, other.low) and np.allclose(self.high, other.high))

    def __repr__(self):
        return f"Box({self.low}, {self.high}, {self.shape}, {self.dtype})"


class Discrete(Space):
    def __init__(self, n, seed=None, start=0):
        self.n = int(n)
        self.start = int(start)
        super().__init__((), np.int64, seed)

    def sample(self):
        return int(self.start + self._np_random.integers(self.n))

    def contains(self, x):
        if isinstance(x, (int, np.integer)):
            as_int = int(x)
        elif isinstance(x, np.ndarray) and x.dtype.kind in "iu" and x.shape == ():
            as_int = int(x)
        else:
            return False
        return self.start <= as_int < self.start + self.n

    def __eq__(self, other):
        return isinstance(other, Discrete) and self.n == other.n and self.start == other.start

    def __repr__(self):
        return f"Discrete({self.n})"


class Tuple(Space):
    def __init__(self, spaces, seed=None):
        self.spaces = tuple(spaces)
        super().__init__(None, None, seed)

    def sample(self):
        return tuple(s.sample() for s in self.spaces)

    def contains(self, x):
        if isinstance(x, (list, np.ndarray)):
            x = tuple(x)
        return (isinstance(x, tuple) and len(x) == len(self.spaces)
                and all(s.contains(p) for s, p in zip(self.spaces, x)))

    def __getitem__(self, index):
        return self.spaces[index]

    def __len__(self):
        return len(self.spaces)

    def __iter__(self):
        return iter(self.spaces)

    def __eq__(self, other):
        return isinstance(other, Tuple) and self.spaces == other.spaces

    def __repr__(self):
        return "Tuple(" + ", ".join(repr(s) for s in self.spaces) + ")"


class Dict(Space, Mapping):
    def __init__(self, spaces=None, seed=None, **spaces_kwargs):
        assert spaces is None or not spaces_kwargs
        if spaces is None:
            spaces = spaces_kwargs
        # gym sorts the keys of a plain mapping; an OrderedDict keeps its order.
        if isinstance(spaces, Mapping) and not isinstance(spaces, OrderedDict):
            try:
                spaces = OrderedDict(sorted(spaces.items()))
            except TypeError:
                spaces = OrderedDict(spaces.items())
        elif isinstance(spaces, (list, tuple)):
            spaces = OrderedDict(spaces)
        self.spaces = spaces
        for s in spaces.values():
            assert isinstance(s, Space), "Values of the dict should be instances of gym.Space"
        Space.__init__(self, None, None, seed)

    def sample(self):
        return OrderedDict((k, s.sample()) for k, s in self.spaces.items())

    def contains(self, x):
        if not isinstance(x, Mapping) or len(x) != len(self.spaces):
            return False
        return all(k in x and s.contains(x[k]) for k, s in self.spaces.items())

    def __getitem__(self, key):
        return self.spaces[key]

    def __setitem__(self, key, value):
        self.spaces[key] = value

    def __iter__(self):
        return iter(self.spaces)

    def __len__(self):
        return len(self.spaces)

    def __eq__(self, other):
        return isinstance(other, Dict) and self.spaces == other.spaces

    def __repr__(self):
        return "Dict(" + ", ".join(f"{k}:{s!r}" for k, s in self.spaces.items()) + ")"


def flatdim(space):
    if isinstance(space, Box):
        return int(np.prod(space.shape))
    if isinstance(space, Discrete):
        return int(space.n)
    if isinstance(space, Tuple):
        return sum(flatdim(s) for s in space.spaces)
    if isinstance(space, Dict):
        return sum(flatdim(s) for s in space.spaces.values())
    raise NotImplementedError(type(space))


def flatten(space, x):
    if isinstance(space, Box):
        return np.asarray(x, dtype=space.dtype).flatten()
    if isinstance(space, Discrete):
        onehot = np.zeros(space.n, dtype=space.dtype)
        onehot[x - space.start] = 1
        return onehot
    if isinstance(space, Tuple):
        parts = [flatten(s, x_part) for x_part, s in zip(x, space.spaces)]
        return np.concatenate(parts) if parts else np.array([])
    if isinstance(space, Dict):
        parts = [flatten(s, x[key]) for key, s in space.spaces.items()]
        return np.concatenate(parts) if parts else np.array([])
    raise NotImplementedError(type(space))


def flatten_space(space):
    if isinstance(space, Box):
        return Box(space.low.flatten(), space.high.flatten(), dtype=space.dtype)
    if isinstance(space, Discrete):
        return Box(low=0, high=1, shape=(space.n,), dtype=space.dtype)
    if isinstance(space, (Tuple, Dict)):
        subs = space.spaces if isinstance(space, Tuple) else space.spaces.values()
        flat = [flatten_space(s) for s in subs]
        if not flat:
            return Box(np.array([]), np.array([]), shape=(0,), dtype=np.float64)
        return Box(low=np.concatenate([s.low for s in flat]),
                   high=np.concatenate([s.high for s in flat]),
                   dtype=np.result_type(*[s.dtype for s in flat]))
    raise NotImplementedError(type(space))
