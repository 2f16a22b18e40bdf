"""``chex`` stand-in (test infrastructure only; see ../README.md)."""
import numpy as _np

from _dataclass import dataclass          # noqa: F401

Array = _np.ndarray
PRNGKey = _np.ndarray
