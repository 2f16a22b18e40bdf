"""``jaxmarl.environments.spaces`` stand-in (structure only)."""
import numpy as _np


class Space:
    pass


class Discrete(Space):
    def __init__(self, num_categories, dtype=_np.int32):
        self.n = num_categories
        self.shape = ()
        self.dtype = dtype


class MultiDiscrete(Space):
    def __init__(self, num_categories):
        self.num_categories = _np.asarray(num_categories)
        self.shape = (len(num_categories),)
        self.dtype = _np.int32


class Box(Space):
    def __init__(self, low, high, shape, dtype=_np.float32):
        self.low, self.high, self.shape, self.dtype = low, high, shape, dtype
