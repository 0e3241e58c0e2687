"""gpflow.utilities: positive() = tfp Softplus bijector, optionally Shift(lower) o Softplus (gpflow/utilities/bijectors.py)."""
import torch as _torch

from .base import set_trainable  # noqa: F401


class _Softplus:
    def __init__(self, lower=None):
        self.lower = lower

    def forward(self, u):
        y = _torch.nn.functional.softplus(u, threshold=1e9)
        return y if self.lower is None else y + self.lower

    def inverse(self, y):
        if self.lower is not None:
            y = y - self.lower
        return y + _torch.log(-_torch.expm1(-y))      # tfp softplus_inverse


def positive(lower=None, base=None):
    return _Softplus(lower)


def to_default_float(x):
    import tensorflow as tf
    return tf.cast(x, tf.float64)
