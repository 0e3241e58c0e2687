"""gpflow.kernels: Kernel, Stationary family, SquaredExponential/RBF (gpflow/kernels/base.py, stationaries.py, 2.5.2)."""
import numpy as _np

import tensorflow as tf

from .base import Module, Parameter
from .utilities import positive


class Kernel(Module):
    def __init__(self, active_dims=None, name=None):
        super().__init__(name=name)
        self._active_dims = active_dims if (active_dims is None or isinstance(active_dims, slice)) else _np.array(active_dims, dtype=int)

    @property
    def active_dims(self):
        return self._active_dims

    def _validate_ard_active_dims(self, ard_parameter):
        if self.active_dims is None or isinstance(self.active_dims, slice):
            return
        if tf.rank(ard_parameter) > 0 and ard_parameter.shape[0] != len(self.active_dims):
            raise ValueError(f'Size of `active_dims` {self.active_dims} does not match size of ard parameter ({ard_parameter.shape[0]})')

    def slice(self, X, X2=None):
        dims = self.active_dims
        if isinstance(dims, slice):
            X = X[..., dims]
            if X2 is not None:
                X2 = X2[..., dims]
        elif dims is not None:
            X = tf.gather(X, dims.tolist(), axis=-1)
            if X2 is not None:
                X2 = tf.gather(X2, dims.tolist(), axis=-1)
        return X, X2

    def __call__(self, X, X2=None, *, full_cov=True, presliced=False):
        if (not full_cov) and (X2 is not None):
            raise ValueError('Ambiguous inputs: `not full_cov` and `X2` are not compatible.')
        X = tf.convert_to_tensor(X, tf.float64)
        X2 = None if X2 is None else tf.convert_to_tensor(X2, tf.float64)
        if not presliced:
            X, X2 = self.slice(X, X2)
        if not full_cov:
            return self.K_diag(X)
        return self.K(X, X2)


def square_distance(X, X2):
    """gpflow/utilities/ops.py: -2 X X2^T + |X|^2 + |X2|^2 (NOT clamped at zero)."""
    if X2 is None:
        Xs = tf.reduce_sum(tf.square(X), axis=-1, keepdims=True)
        dist = -2 * tf.matmul(X, X, transpose_b=True)
        dist = dist + Xs + tf.linalg.adjoint(Xs)
        return dist
    Xs = tf.reduce_sum(tf.square(X), axis=-1)
    X2s = tf.reduce_sum(tf.square(X2), axis=-1)
    dist = -2 * tf.tensordot(X, X2, [[-1], [-1]])
    dist = dist + Xs[..., :, None] + X2s[..., None, :]      # broadcasting_elementwise(tf.add, Xs, X2s) for rank-2 inputs
    return dist


def difference_matrix(X, X2):
    """gpflow/utilities/ops.py: [N..., N2..., D] pairwise differences; leading dims of X and X2 are flattened and restored."""
    if X2 is None:
        X2 = X
        return X[..., :, None, :] - X2[..., None, :, :]
    Xshape, X2shape = X.shape, X2.shape
    Xf = tf.reshape(X, (-1, Xshape[-1]))
    X2f = tf.reshape(X2, (-1, X2shape[-1]))
    diff = Xf[:, None, :] - X2f[None, :, :]
    return tf.reshape(diff, tuple(Xshape[:-1]) + tuple(X2shape[:-1]) + (Xshape[-1],))


class Stationary(Kernel):
    def __init__(self, variance=1.0, lengthscales=1.0, **kwargs):
        for kwarg in kwargs:
            if kwarg not in {'name', 'active_dims'}:
                raise TypeError(f'Unknown keyword argument: {kwarg}')
        super().__init__(**kwargs)
        self.variance = Parameter(variance, transform=positive())
        self.lengthscales = Parameter(lengthscales, transform=positive())
        self._validate_ard_active_dims(self.lengthscales)

    @property
    def ard(self):
        return self.lengthscales.shape.ndims > 0 if hasattr(self.lengthscales.shape, 'ndims') else len(self.lengthscales.shape) > 0

    def scale(self, X):
        return X / self.lengthscales if X is not None else X

    def K_diag(self, X):
        return tf.fill(X.shape[:-1], tf.squeeze(self.variance.__tf_tensor__()))


class IsotropicStationary(Stationary):
    def K(self, X, X2=None):
        return self.K_r2(self.scaled_squared_euclid_dist(X, X2))

    def scaled_squared_euclid_dist(self, X, X2=None):
        return square_distance(self.scale(X), self.scale(X2))


class AnisotropicStationary(Stationary):
    def K(self, X, X2=None):
        return self.K_d(self.scaled_difference_matrix(X, X2))

    def scaled_difference_matrix(self, X, X2=None):
        return difference_matrix(self.scale(X), self.scale(X2))


class SquaredExponential(IsotropicStationary):
    def K_r2(self, r2):
        return self.variance * tf.exp(-0.5 * r2)


RBF = SquaredExponential
