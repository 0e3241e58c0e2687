"""gpflow.config: dtype defaults (the reference forces float64, romcomma/user/contexts.py:67)."""
import numpy as _np

_float, _int = _np.float64, _np.int32


def default_float():
    return _float


def default_int():
    return _int


def default_jitter():
    return 1e-6


def set_default_float(value):
    global _float
    _float = value


def set_default_int(value):
    global _int
    _int = value
