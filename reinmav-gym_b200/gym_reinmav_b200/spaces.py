"""``Box`` space.  Uses gym's / gymnasium's class when one of them is importable, else a
minimal stand-in with the attributes the reference's consumers read (``low``, ``high``,
``shape``, ``dtype``, ``sample()``, ``contains()``)."""
import numpy as np


class _Box:
    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.dtype = np.dtype(dtype)
        self.shape = tuple(shape) if shape is not None else np.shape(low)
        self.low = np.full(self.shape, low, dtype=self.dtype)
        self.high = np.full(self.shape, high, dtype=self.dtype)
        self._rng = np.random.RandomState()

    def seed(self, seed=None):
        self._rng = np.random.RandomState(seed)
        return [seed]

    def sample(self):
        return self._rng.uniform(self.low, self.high).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    def __repr__(self):
        return "Box(%s, %s, %s, %s)" % (self.low.min(), self.high.max(), self.shape, self.dtype)

    def __eq__(self, other):
        return (hasattr(other, "low") and self.shape == tuple(other.shape)
                and np.allclose(self.low, other.low) and np.allclose(self.high, other.high))


def _find_box():
    for mod in ("gym", "gymnasium"):
        try:
            return __import__(mod + ".spaces", fromlist=["Box"]).Box
        except Exception:
            continue
    return _Box


Box = _find_box()
