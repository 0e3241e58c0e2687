"""tf.linalg stand-ins (torch CPU float64)."""
from __future__ import annotations

import torch as _torch

from ._core import Tensor, as_t as _t


def cholesky(x):
    """tf.linalg.cholesky raises InvalidArgumentError on a non-PD input; torch raises LinAlgError - both propagate as exceptions."""
    return Tensor.wrap(_torch.linalg.cholesky(_t(x)))


def triangular_solve(matrix, rhs, lower=True, adjoint=False):
    """Batch dimensions broadcast (TF >= 2.2)."""
    A, B = _t(matrix), _t(rhs)
    if adjoint:
        A, lower = A.transpose(-1, -2), not lower
    return Tensor.wrap(_torch.linalg.solve_triangular(A, B, upper=not lower))


def cholesky_solve(chol, rhs):
    L, B = _t(chol), _t(rhs)
    y = _torch.linalg.solve_triangular(L, B, upper=False)
    return Tensor.wrap(_torch.linalg.solve_triangular(L.transpose(-1, -2), y, upper=True))


def diag_part(x):
    return Tensor.wrap(_torch.diagonal(_t(x), dim1=-2, dim2=-1))


def diag(x):
    return Tensor.wrap(_torch.diag_embed(_t(x)))


def set_diag(x, diagonal):
    x, d = _t(x), _t(diagonal)
    n = min(x.shape[-2], x.shape[-1])
    mask = _torch.eye(x.shape[-2], x.shape[-1], dtype=_torch.bool)
    return Tensor.wrap(_torch.where(mask, _torch.diag_embed(d.broadcast_to(tuple(x.shape[:-2]) + (n,))), x))


def band_part(x, num_lower, num_upper):
    x = _t(x)
    m, n = x.shape[-2], x.shape[-1]
    i = _torch.arange(m)[:, None]
    j = _torch.arange(n)[None, :]
    keep = _torch.ones((m, n), dtype=_torch.bool)
    if num_lower >= 0:
        keep &= (i - j) <= num_lower
    if num_upper >= 0:
        keep &= (j - i) <= num_upper
    return Tensor.wrap(_torch.where(keep, x, _torch.zeros((), dtype=x.dtype)))


def trace(x):
    return Tensor.wrap(_torch.diagonal(_t(x), dim1=-2, dim2=-1).sum(-1))


def adjoint(x):
    return _t(x).transpose(-1, -2)


def matmul(a, b, transpose_a=False, transpose_b=False):
    a, b = _t(a), _t(b)
    if transpose_a:
        a = a.transpose(-1, -2)
    if transpose_b:
        b = b.transpose(-1, -2)
    return a @ b
