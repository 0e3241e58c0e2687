""" Type and constant definitions (reference romcomma/base/definitions.py:36-93), with torch/numpy carriers instead of TensorFlow."""
from __future__ import annotations

from abc import abstractmethod  # noqa: F401  (re-exported: every reference module star-imports this file)
from pathlib import Path  # noqa: F401
from typing import *  # noqa: F401,F403

import numpy as np
import pandas as pd  # noqa: F401
import torch

from romcomma import gf_compat as gf  # noqa: F401
import romcomma.gpf as mf  # noqa: F401

EFFECTIVELY_ZERO = 1.0E-64  #: Tolerance when testing floats for equality.


def INT() -> Type:
    """ The ``dtype`` of ``int`` (gpflow default_int)."""
    return gf.config.default_int()


def FLOAT() -> Type:
    """ The ``dtype`` of ``float``: always float64 on this path."""
    return gf.config.default_float()


class classproperty:
    """Read-only class-level property. The reference stacks ``@classmethod @property`` (removed in Python 3.13)."""

    def __init__(self, fget):
        self.fget = fget
        self.__doc__ = fget.__doc__

    def __get__(self, instance, owner=None):
        return self.fget(owner if owner is not None else type(instance))


class NP:
    """ Extended numpy types."""
    Array = Tensor = Tensor1 = Tensor2 = Vector = Covector = Matrix = Tensor3 = Tensor4 = Tensor5 = Tensor6 = Tensor7 = Tensor8 = np.ndarray
    VectorLike = Union[int, float, Sequence[Union[int, float]], np.ndarray]
    MatrixLike = Union[VectorLike, Sequence[VectorLike]]
    CovectorLike = MatrixLike
    ArrayLike = TensorLike = Union[MatrixLike, Sequence[MatrixLike], Sequence[Sequence[MatrixLike]]]


class TF:
    """ Device tensor types: torch CUDA float64 tensors stand where the reference has tf.Tensor."""
    Array = Tensor = Tensor1 = Tensor2 = Vector = Covector = Matrix = Tensor3 = Tensor4 = Tensor5 = Tensor6 = Tensor7 = Tensor8 = torch.Tensor
    VectorLike = Union[int, float, Sequence[Union[int, float]], torch.Tensor]
    MatrixLike = Union[VectorLike, Sequence[VectorLike]]
    CovectorLike = MatrixLike
    ArrayLike = TensorLike = Union[MatrixLike, Sequence[MatrixLike], Sequence[Sequence[MatrixLike]]]
    Slice = PairOfInts = Sequence[int]      #: A slice [m0:m1], for indexing and marginalization.
    NaN = float('nan')
