"""A torch-CPU (float64) stand-in for the handful of ``tensorflow`` symbols rom-comma's hot path calls.  TEST INFRASTRUCTURE ONLY.

Purpose: let ``tests/golden/make_golden_from_reference.py`` import and EXECUTE the reference's own, unmodified source files
(/root/reference/romcomma/{gpf,gpr,gsa,base,data}/*.py) in a container that has no TensorFlow, so that the golden vectors under
tests/golden/ come from the reference's code rather than from a transliteration of it.  Each function below has TensorFlow's
documented semantics for the argument patterns the reference uses (cited per function); autograd stands in for tf.GradientTape.
This package is never importable from the product (it lives outside every import path except the generator's own sys.path).
"""
from __future__ import annotations

import math as _math
from typing import Any, Sequence

import numpy as _np
import torch as _torch

from . import linalg  # noqa: E402,F401  (defined below via submodule)
from ._core import (Tensor, TensorShape, Variable, as_t as _t, float64, int32, int64, bool_ as bool, newaxis,  # noqa: F401
                    DType)

__version__ = '0.0-torch-shim'


# ---- construction ------------------------------------------------------------------------------------------------------
def constant(value, dtype=None, shape=None, name=None):
    t = _t(value, dtype)
    return t.reshape(tuple(shape)) if shape is not None else t


def convert_to_tensor(value, dtype=None, name=None):
    return _t(value, dtype)


def identity(x, name=None):
    return _t(x).clone()


def cast(x, dtype):
    return _t(x).to(DType.torch(dtype))


def eye(num_rows, num_columns=None, dtype=float64):
    return Tensor.wrap(_torch.eye(int(num_rows), int(num_columns or num_rows), dtype=DType.torch(dtype)))


def fill(dims, value):
    """Differentiable in ``value`` (gpflow's add_noise_cov fills the likelihood variance, a trainable parameter)."""
    value = _t(value)
    dims = tuple(int(d) for d in (_t(dims).tolist() if not isinstance(dims, (tuple, list)) else dims))
    return Tensor.wrap(value.reshape(()) * _torch.ones(dims, dtype=value.dtype))


def zeros(shape, dtype=float64):
    return Tensor.wrap(_torch.zeros(tuple(int(s) for s in shape), dtype=DType.torch(dtype)))


def ones(shape, dtype=float64):
    return Tensor.wrap(_torch.ones(tuple(int(s) for s in shape), dtype=DType.torch(dtype)))


# ---- shape manipulation ------------------------------------------------------------------------------------------------
def _shape_arg(shape):
    if isinstance(shape, _torch.Tensor):
        return tuple(int(s) for s in shape.tolist())
    return tuple(int(s) for s in shape)


def shape(x, out_type=None):
    """tf.shape: a rank-1 integer tensor (the reference indexes it, calls .numpy() on it and feeds it to tf.concat)."""
    return Tensor.wrap(_torch.tensor(list(_t(x).shape), dtype=_torch.int64))


def rank(x):
    """tf.rank: the reference only compares it with / does integer arithmetic on Python ints, so a Python int is returned."""
    return _t(x).dim()


def reshape(x, shape, name=None):
    return _t(x).reshape(_shape_arg(shape))


def transpose(x, perm=None, conjugate=False):
    x = _t(x)
    if perm is None:
        perm = tuple(reversed(range(x.dim())))
    return x.permute(tuple(int(p) for p in perm))


def expand_dims(x, axis):
    x = _t(x)
    return x.unsqueeze(int(axis))


def squeeze(x, axis=None):
    x = _t(x)
    if axis is None:
        return x.squeeze()
    if isinstance(axis, (list, tuple)):
        for a in sorted((a if a >= 0 else a + x.dim() for a in axis), reverse=True):
            x = x.squeeze(a)
        return x
    return x.squeeze(int(axis))


def broadcast_to(x, shape):
    return _t(x).broadcast_to(_shape_arg(shape)).clone()


def concat(values, axis, name=None):
    vals = [_t(v) for v in values]
    return Tensor.wrap(_torch.cat(vals, dim=int(axis)))


def stack(values, axis=0):
    return Tensor.wrap(_torch.stack([_t(v) for v in values], dim=int(axis)))


def gather(params, indices, axis=0):
    return _t(params).index_select(int(axis), _torch.as_tensor(list(indices) if not isinstance(indices, _torch.Tensor) else indices, dtype=_torch.int64))


# ---- elementwise / reductions ------------------------------------------------------------------------------------------
def sqrt(x):
    return _t(x).sqrt()


def exp(x):
    return _t(x).exp()


def abs(x):  # noqa: A001
    return _t(x).abs()


def square(x):
    x = _t(x)
    return x * x


def divide(x, y):
    return _t(x) / _t(y)


def add(x, y):
    return _t(x) + _t(y)


def _axis(axis):
    if axis is None:
        return None
    return tuple(int(a) for a in axis) if isinstance(axis, (list, tuple)) else int(axis)


def reduce_sum(x, axis=None, keepdims=False):
    x = _t(x)
    return x.sum() if axis is None else x.sum(dim=_axis(axis), keepdim=keepdims)


def reduce_prod(x, axis=None, keepdims=False):
    """Also called on a TensorShape (gsa/base.py:134,143 ``tf.reduce_prod(tensor.shape)``)."""
    if isinstance(x, (tuple, list, _torch.Size)):
        return Tensor.wrap(_torch.tensor(_math.prod(int(s) for s in x), dtype=_torch.int64))
    x = _t(x)
    if axis is None:
        return x.prod()
    return x.prod(dim=int(axis), keepdim=keepdims)


def matmul(a, b, transpose_a=False, transpose_b=False):
    a, b = _t(a), _t(b)
    if transpose_a:
        a = a.transpose(-1, -2)
    if transpose_b:
        b = b.transpose(-1, -2)
    return a @ b


def tensordot(a, b, axes):
    return Tensor.wrap(_torch.tensordot(_t(a), _t(b), dims=axes))


def einsum(equation: str, *operands):
    """tf.einsum accepts blanks inside the equation and an implicit '...'; torch.einsum accepts the same once blanks are removed."""
    return Tensor.wrap(_torch.einsum(equation.replace(' ', ''), *[_t(o) for o in operands]))


def assert_equal(x, y, message=None, **kwargs):
    a = x.tolist() if isinstance(x, _torch.Tensor) else x
    b = y.tolist() if isinstance(y, _torch.Tensor) else y
    if a != b:
        raise AssertionError(message or f'{a} != {b}')


def function(fn=None, **kwargs):
    """tf.function: eager execution is semantically identical for this path."""
    if fn is None:
        return lambda f: f
    return fn


class math_ns:
    exp = staticmethod(exp)
    sqrt = staticmethod(sqrt)
    log = staticmethod(lambda x: _t(x).log())
    square = staticmethod(square)
    softplus = staticmethod(lambda x: Tensor.wrap(_torch.nn.functional.softplus(_t(x), threshold=1e9)))


math = math_ns  # noqa: A001  (shadows the stdlib name inside this module only after its last use above)


class RaggedTensor:
    """Only ``from_row_lengths(values, row_lengths).to_tensor(default_value=0, shape=(L, L))`` (romcomma/gpf/base.py:45)."""

    def __init__(self, values, row_lengths):
        self.values, self.row_lengths = _t(values), tuple(int(r) for r in row_lengths)

    @classmethod
    def from_row_lengths(cls, values, row_lengths):
        return cls(values, row_lengths)

    def to_tensor(self, default_value=0, shape=None):
        rows = len(self.row_lengths)
        cols = max(self.row_lengths) if self.row_lengths else 0
        if shape is not None:
            rows, cols = int(shape[0]), int(shape[1])
        out = _torch.full((rows, cols), float(default_value), dtype=self.values.dtype)
        k = 0
        pieces = []
        for i, n in enumerate(self.row_lengths):      # scatter, differentiably
            for j in range(n):
                pieces.append((i, j, k))
                k += 1
        if pieces:
            idx_i = _torch.tensor([p[0] for p in pieces])
            idx_j = _torch.tensor([p[1] for p in pieces])
            out = out.index_put((idx_i, idx_j), self.values[_torch.tensor([p[2] for p in pieces])])
        return Tensor.wrap(out)


class GradientTape:
    """Eager torch autograd records everything already; ``jacobian`` is what romcomma/gpr/models.py:396 needs
    (``tape.jacobian(KXx, x)`` -> shape ``KXx.shape + x.shape``), one reverse pass per target element (small cases only)."""

    def __init__(self, persistent=False, watch_accessed_variables=True):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def watch(self, tensor):
        return None

    def jacobian(self, target, sources, **kwargs):
        target, sources = _t(target), _t(sources)
        flat = target.reshape(-1)
        rows = []
        for i in range(flat.numel()):
            g, = _torch.autograd.grad(flat[i], sources, retain_graph=True, allow_unused=True)
            rows.append(_torch.zeros_like(sources) if g is None else g)
        return Tensor.wrap(_torch.stack(rows).reshape(tuple(target.shape) + tuple(sources.shape)).detach())

    def gradient(self, target, sources, **kwargs):
        return Tensor.wrap(_torch.autograd.grad(_t(target), _t(sources), retain_graph=True)[0])


class _Data:
    class Dataset:
        @staticmethod
        def from_tensor_slices(x):
            return x


data = _Data
