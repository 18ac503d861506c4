"""Observation / action space declarations.

The reference declares spaces with gymnasium (`gym.spaces.Box` ...).  gymnasium is used when
it is importable; otherwise these structural stand-ins (same constructor signatures,
`shape`, `sample()`, `==`) keep env definitions working unchanged.
"""
from __future__ import annotations

import numpy as np

try:  # pragma: no cover - gymnasium is absent from the build image
    from gymnasium import Space  # type: ignore
    from gymnasium.spaces import Box, Dict, Discrete, Tuple  # type: ignore

    HAVE_GYMNASIUM = True
except ImportError:
    HAVE_GYMNASIUM = False

    class Space:  # type: ignore[no-redef]
        shape = None
        dtype = None

        def contains(self, x) -> bool:
            return True

        def __contains__(self, x) -> bool:
            return self.contains(x)

    class Box(Space):  # type: ignore[no-redef]
        def __init__(self, low, high, shape=None, dtype=np.float32):
            lo, hi = np.asarray(low), np.asarray(high)
            if shape is None:
                shape = lo.shape if lo.shape else hi.shape
            self.shape = tuple(shape)
            self.dtype = np.dtype(dtype)
            self.low = np.broadcast_to(lo, self.shape).astype(np.float64)
            self.high = np.broadcast_to(hi, self.shape).astype(np.float64)

        def sample(self):
            lo = np.where(np.isfinite(self.low), self.low, -1.0)
            hi = np.where(np.isfinite(self.high), self.high, 1.0)
            return np.random.uniform(lo, hi).astype(self.dtype)

        def contains(self, x) -> bool:
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

        def __eq__(self, other):
            return (isinstance(other, Box) and self.shape == other.shape
                    and self.dtype == other.dtype and np.array_equal(self.low, other.low)
                    and np.array_equal(self.high, other.high))

        def __repr__(self):
            return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"

    class Discrete(Space):  # type: ignore[no-redef]
        def __init__(self, n, start=0):
            self.n, self.start = int(n), int(start)
            self.shape, self.dtype = (), np.dtype(np.int64)

        def sample(self):
            return self.start + int(np.random.randint(self.n))

        def contains(self, x) -> bool:
            return self.start <= int(x) < self.start + self.n

        def __eq__(self, other):
            return isinstance(other, Discrete) and (self.n, self.start) == (other.n, other.start)

    class Dict(Space):  # type: ignore[no-redef]
        def __init__(self, spaces=None, **kw):
            self.spaces = dict(spaces or {}, **kw)

        def sample(self):
            return {k: s.sample() for k, s in self.spaces.items()}

        def __getitem__(self, k):
            return self.spaces[k]

        def __eq__(self, other):
            return isinstance(other, Dict) and self.spaces == other.spaces

    class Tuple(Space):  # type: ignore[no-redef]
        def __init__(self, spaces):
            self.spaces = tuple(spaces)

        def sample(self):
            return tuple(s.sample() for s in self.spaces)

        def __getitem__(self, i):
            return self.spaces[i]

        def __len__(self):
            return len(self.spaces)

        def __eq__(self, other):
            return isinstance(other, Tuple) and self.spaces == other.spaces
