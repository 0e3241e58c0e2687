"""Tensor carrier of the tensorflow stand-in: a torch.Tensor subclass whose ``.shape`` behaves like tf.TensorShape."""
from __future__ import annotations

import numpy as _np
import torch as _torch

newaxis = None


class DType:
    def __init__(self, name, tdtype):
        self.name, self._t = name, tdtype

    def __repr__(self):
        return f'tf.{self.name}'

    @staticmethod
    def torch(d):
        if d is None:
            return None
        if isinstance(d, DType):
            return d._t
        if isinstance(d, _torch.dtype):
            return d
        if d in (float, 'float64', _np.float64):
            return _torch.float64
        if d in (int, 'int32', _np.int32):
            return _torch.int32
        if d in ('int64', _np.int64):
            return _torch.int64
        if d in (bool, _np.bool_):
            return _torch.bool
        return _torch.from_numpy(_np.zeros(0, dtype=d)).dtype


float64, int32, int64, bool_ = DType('float64', _torch.float64), DType('int32', _torch.int32), DType('int64', _torch.int64), DType('bool', _torch.bool)


class TensorShape(tuple):
    """tf.TensorShape: tuple-like, ``as_list()``, slices stay TensorShapes, compares equal to tuples/lists."""

    def as_list(self):
        return list(self)

    def __getitem__(self, item):
        r = tuple.__getitem__(self, item)
        return TensorShape(r) if isinstance(item, slice) else r

    def __eq__(self, other):
        try:
            return tuple(self) == tuple(other)
        except TypeError:
            return False

    def __ne__(self, other):
        return not self.__eq__(other)

    __hash__ = tuple.__hash__

    def __add__(self, other):
        return TensorShape(tuple(self) + tuple(other))


class Tensor(_torch.Tensor):
    @staticmethod
    def wrap(t: _torch.Tensor) -> 'Tensor':
        return t if isinstance(t, Tensor) else t.as_subclass(Tensor)

    @property
    def shape(self):
        return TensorShape(_torch.Tensor.shape.__get__(self))

    @property
    def ndims(self):
        return self.dim()

    def numpy(self):
        return self.detach().as_subclass(_torch.Tensor).numpy()

    def __array__(self, dtype=None, copy=None):
        a = self.detach().as_subclass(_torch.Tensor).numpy()
        return a.astype(dtype) if dtype is not None else a

    def __iter__(self):                       # `min(tensor)` in romcomma/gpf/base.py:88
        return iter([self[i] for i in range(_torch.Tensor.shape.__get__(self)[0])])

    def __bool__(self):
        return bool(self.detach().as_subclass(_torch.Tensor).item())

    # tf.Tensor is an immutable value: ``a += b`` rebinds ``a`` to a new tensor (torch would mutate in place and alias), and a
    # deep copy may share storage.
    def __iadd__(self, other):
        return self + other

    def __isub__(self, other):
        return self - other

    def __imul__(self, other):
        return self * other

    def __itruediv__(self, other):
        return self / other

    def __deepcopy__(self, memo):
        return self


def as_t(x, dtype=None) -> Tensor:
    """Anything tensor-like (incl. gpflow-shim Parameters, numpy arrays, python scalars/sequences, TensorShapes) -> Tensor."""
    td = DType.torch(dtype)
    if hasattr(x, '__tf_tensor__'):
        x = x.__tf_tensor__()
    if isinstance(x, _torch.Tensor):
        t = x
    elif isinstance(x, _np.ndarray):
        t = _torch.from_numpy(_np.ascontiguousarray(x)) if x.dtype != object else _torch.tensor(x.tolist())
    elif isinstance(x, (bool, int)):
        t = _torch.tensor(x)
    elif isinstance(x, float):
        t = _torch.tensor(x, dtype=_torch.float64)
    else:
        if isinstance(x, (list, tuple)) and any(isinstance(e, _torch.Tensor) or hasattr(e, '__tf_tensor__') for e in x):
            t = _torch.stack([as_t(e) for e in x])
        else:
            t = _torch.as_tensor(_np.asarray(x))
    if td is not None and t.dtype != td:
        t = t.to(td)
    return Tensor.wrap(t)


class Variable(Tensor):
    """tf.Variable(value): a leaf tensor that requires grad (only used by predict_gradient, romcomma/gpr/models.py:387)."""

    def __new__(cls, value, dtype=None, **kwargs):
        t = as_t(value, dtype).detach().clone().as_subclass(_torch.Tensor)
        t.requires_grad_(t.dtype.is_floating_point)
        return t.as_subclass(cls)
