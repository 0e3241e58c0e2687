"""gpflow.base: Parameter (a constrained view of an unconstrained tf.Variable), Module, set_trainable."""
from __future__ import annotations

import numpy as _np
import torch as _torch

import tensorflow as tf
from tensorflow._core import Tensor as _Tensor, as_t as _t


class Module:
    """tf.Module-like container: parameters are discovered by walking attributes in sorted-name order (tf.Module semantics, which is
    what fixes the packing order of gpflow's Scipy optimizer)."""

    def __init__(self, name=None):
        self._name = name

    @property
    def name(self):
        return self._name

    def _walk(self, seen):
        if id(self) in seen:
            return
        seen.add(id(self))
        for key in sorted(vars(self)):
            yield from _walk_value(vars(self)[key], seen)

    @property
    def parameters(self):
        return tuple(self._walk(set()))

    @property
    def trainable_parameters(self):
        return tuple(p for p in self.parameters if p.trainable)

    @property
    def trainable_variables(self):
        return tuple(p.unconstrained_variable for p in self.trainable_parameters)


def _walk_value(v, seen):
    if isinstance(v, Parameter):
        if id(v) not in seen:
            seen.add(id(v))
            yield v
    elif isinstance(v, Module):
        yield from v._walk(seen)
    elif isinstance(v, (list, tuple)):
        for e in v:
            yield from _walk_value(e, seen)
    elif isinstance(v, dict):
        for k in sorted(v):
            yield from _walk_value(v[k], seen)


class Parameter:
    """value = transform.forward(unconstrained_variable).  Tensor-like through __tf_tensor__ / __torch_function__ / arithmetic dunders."""

    def __init__(self, value, *, transform=None, prior=None, trainable=True, dtype=None, name=None):
        if isinstance(value, Parameter):
            transform = transform or value.transform
            value = value.__tf_tensor__()
        v = _t(value, tf.float64).detach().clone().as_subclass(_torch.Tensor)
        self.transform = transform
        u = v if transform is None else transform.inverse(v)
        self.unconstrained_variable = u.clone().requires_grad_(True)
        self.trainable = trainable
        self.name = name

    # ---- tensor protocol ----
    def __tf_tensor__(self):
        u = self.unconstrained_variable
        return _Tensor.wrap(u if self.transform is None else self.transform.forward(u))

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        conv = lambda a: a.__tf_tensor__() if isinstance(a, Parameter) else a  # noqa: E731
        args = tuple(conv(a) if not isinstance(a, (list, tuple)) else type(a)(conv(e) for e in a) for a in args)
        kwargs = {k: conv(v) for k, v in (kwargs or {}).items()}
        return func(*args, **kwargs)

    def numpy(self):
        return self.__tf_tensor__().numpy()

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        return a.astype(dtype) if dtype is not None else a

    @property
    def shape(self):
        return self.__tf_tensor__().shape

    def __getitem__(self, item):
        return self.__tf_tensor__()[item]

    def __iter__(self):
        return iter(self.__tf_tensor__())

    def assign(self, value):
        v = _t(value, tf.float64).detach().as_subclass(_torch.Tensor)
        u = v if self.transform is None else self.transform.inverse(v)
        with _torch.no_grad():
            self.unconstrained_variable.copy_(u.reshape(self.unconstrained_variable.shape))

    def _binary(name):  # noqa: N805
        def op(self, other):
            return getattr(self.__tf_tensor__(), name)(_t(other))
        return op

    __add__ = _binary('__add__'); __radd__ = _binary('__radd__'); __sub__ = _binary('__sub__'); __rsub__ = _binary('__rsub__')
    __mul__ = _binary('__mul__'); __rmul__ = _binary('__rmul__'); __truediv__ = _binary('__truediv__'); __rtruediv__ = _binary('__rtruediv__')
    __pow__ = _binary('__pow__'); __lt__ = _binary('__lt__'); __le__ = _binary('__le__'); __gt__ = _binary('__gt__'); __ge__ = _binary('__ge__')

    def __neg__(self):
        return -self.__tf_tensor__()

    def __float__(self):
        return float(self.__tf_tensor__())


def set_trainable(model, flag: bool):
    """gpflow.utilities.set_trainable: a Parameter, a Module or an iterable of them."""
    if isinstance(model, Parameter):
        model.trainable = flag
    elif isinstance(model, Module):
        for p in model.parameters:
            p.trainable = flag
    else:
        for m in model:
            set_trainable(m, flag)
