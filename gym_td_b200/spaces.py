"""Action / observation spaces with the gym<=0.21 surface the reference uses (`contains`, `sample`,
`shape`, `n`).  When a real `gym` is importable its classes are used instead, so that the envs plug
into gym wrappers unchanged."""
import numpy as np

try:  # pragma: no cover - gym is not installed in the build image
    from gym.spaces import Box, Dict, Discrete  # noqa: F401
    HAVE_GYM = True
except Exception:
    HAVE_GYM = False

    class _Space(object):
        def __init__(self, shape, dtype):
            self.shape = None if shape is None else tuple(shape)
            self.dtype = None if dtype is None else np.dtype(dtype)
            self.np_random = np.random.RandomState()

        def seed(self, seed=None):
            self.np_random = np.random.RandomState(seed)
            return [seed]

        def __contains__(self, x):
            return self.contains(x)

    class Box(_Space):
        def __init__(self, low, high, shape=None, dtype=np.float32):
            super().__init__(np.shape(low) if shape is None else shape, dtype)
            self.low = np.full(self.shape, low, dtype=self.dtype)
            self.high = np.full(self.shape, high, dtype=self.dtype)

        def contains(self, x):
            if not isinstance(x, np.ndarray):
                try:
                    x = np.asarray(x, dtype=self.dtype)
                except (TypeError, ValueError):
                    return False
            return bool(np.can_cast(x.dtype, self.dtype) and x.shape == self.shape
                        and np.all(x >= self.low) and np.all(x <= self.high))

        def sample(self):
            if self.dtype.kind == "f":
                return self.np_random.uniform(self.low, self.high, self.shape).astype(self.dtype)
            return self.np_random.randint(self.low, self.high + 1, self.shape).astype(self.dtype)

        def __repr__(self):
            return "Box(%s, %s, %s, %s)" % (self.low.min(), self.high.max(), self.shape, self.dtype)

    class Discrete(_Space):
        def __init__(self, n):
            super().__init__((), np.int64)
            self.n = int(n)

        def contains(self, x):
            if isinstance(x, (int, np.integer)) and not isinstance(x, bool):
                v = int(x)
            elif isinstance(x, np.ndarray) and x.dtype.kind in "iu" and x.shape == ():
                v = int(x)
            else:
                return False
            return 0 <= v < self.n

        def sample(self):
            return int(self.np_random.randint(self.n))

        def __repr__(self):
            return "Discrete(%d)" % self.n

    class Dict(_Space):
        def __init__(self, spaces=None, **kw):
            super().__init__(None, None)
            self.spaces = dict(spaces or {}, **kw)

        def contains(self, x):
            return isinstance(x, dict) and len(x) == len(self.spaces) and \
                all(k in x and s.contains(x[k]) for k, s in self.spaces.items())

        def sample(self):
            return {k: s.sample() for k, s in self.spaces.items()}

        def __getitem__(self, key):
            return self.spaces[key]

        def __repr__(self):
            return "Dict(%r)" % (self.spaces,)
