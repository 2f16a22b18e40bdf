"""``flax`` stand-in: ``struct.dataclass`` is real (pytree dataclass); ``linen`` and ``training`` exist so
that the reference learner module imports (its networks are never instantiated by the fixtures)."""
from . import struct          # noqa: F401
