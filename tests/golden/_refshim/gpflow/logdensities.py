"""gpflow.logdensities.multivariate_normal (gpflow/logdensities.py, 2.5.2)."""
import numpy as _np

import tensorflow as tf


def multivariate_normal(x, mu, L):
    """x, mu: [N, D] (D independent columns) ; L: [N, N] lower Cholesky factor of the covariance.  Returns the [D] log densities."""
    d = x - mu
    alpha = tf.linalg.triangular_solve(L, d, lower=True)
    num_dims = tf.cast(tf.shape(d)[0], tf.float64)
    p = -0.5 * tf.reduce_sum(tf.square(alpha), 0)
    p = p - 0.5 * num_dims * _np.log(2 * _np.pi)
    p = p - tf.reduce_sum(tf.math.log(tf.linalg.diag_part(L)))
    return p
