"""gpflow.conditionals.base_conditional (gpflow/conditionals/util.py, 2.5.2), the q_sqrt=None branch that GPR-style models use."""
import tensorflow as tf


def base_conditional(Kmn, Kmm, Knn, f, *, full_cov=False, q_sqrt=None, white=False):
    Lm = tf.linalg.cholesky(Kmm)
    return base_conditional_with_lm(Kmn=Kmn, Lm=Lm, Knn=Knn, f=f, full_cov=full_cov, q_sqrt=q_sqrt, white=white)


def base_conditional_with_lm(Kmn, Lm, Knn, f, *, full_cov=False, q_sqrt=None, white=False):
    assert q_sqrt is None, 'the reference never passes q_sqrt'
    num_func = f.shape[-1]                                  # R
    N = Kmn.shape[-1]
    A = tf.linalg.triangular_solve(Lm, Kmn, lower=True)     # [M, N]
    if full_cov:
        fvar = Knn - tf.linalg.matmul(A, A, transpose_a=True)                    # [N, N]
        fvar = tf.broadcast_to(tf.expand_dims(fvar, -3), [num_func, N, N])       # [R, N, N]
    else:
        fvar = Knn - tf.reduce_sum(tf.square(A), -2)                             # [N]
        fvar = tf.broadcast_to(tf.expand_dims(fvar, -2), [num_func, N])          # [R, N]
    if not white:
        A = tf.linalg.triangular_solve(tf.linalg.adjoint(Lm), A, lower=False)
    fmean = tf.linalg.matmul(A, f, transpose_a=True)                             # [N, R]
    if not full_cov:
        fvar = tf.linalg.adjoint(fvar)                                           # [N, R]
    return fmean, fvar
