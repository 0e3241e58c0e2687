"""gpflow.mean_functions: MeanFunction base and Zero (gpflow/mean_functions.py)."""
import tensorflow as tf

from .base import Module


class MeanFunction(Module):
    def __init__(self, name=None):
        super().__init__(name=name)

    def __call__(self, X):
        raise NotImplementedError


class Zero(MeanFunction):
    def __init__(self, output_dim=1):
        super().__init__()
        self.output_dim = output_dim

    def __call__(self, X):
        X = tf.convert_to_tensor(X)
        return tf.zeros(tuple(X.shape[:-1]) + (self.output_dim,), dtype=tf.float64)
