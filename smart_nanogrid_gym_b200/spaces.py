"""Minimal `Box` space, used when neither `gymnasium` nor `gym` is installed (they are not in the
build image).  Same attributes the reference relies on (envs/smart_nanogrid_environment.py:98-120)."""
import numpy as np


def _find_gym_box():
    for mod in ("gymnasium", "gym"):
        try:
            return __import__(mod).spaces.Box
        except Exception:  # noqa: BLE001 - any import problem means "not available"
            continue
    return None


class _Box:
    def __init__(self, low, high, shape=None, dtype=np.float32):
        shp = tuple(shape) if shape is not None else np.shape(low)
        self.low = np.broadcast_to(np.asarray(low, dtype=dtype), shp).copy()
        self.high = np.broadcast_to(np.asarray(high, dtype=dtype), shp).copy()
        self.shape = self.low.shape
        self.dtype = np.dtype(dtype)
        self._rng = np.random.default_rng()

    def seed(self, seed=None):
        self._rng = np.random.default_rng(seed)
        return [seed]

    def sample(self):
        return self._rng.uniform(self.low, self.high).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    def __repr__(self):
        return "Box(%s, %s, %s, %s)" % (self.low.min(), self.high.max(), self.shape, self.dtype)


Box = _find_gym_box() or _Box
