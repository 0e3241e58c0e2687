""" Mean functions for gpf - prior predictions (reference romcomma/gpf/mean_functions.py:30-64). Only the zero mean is on the hot path."""
from __future__ import annotations

from typing import Sequence, Union

import torch

from romcomma._tensors import DeviceTensor, as_device


class MeanFunction:
    def __call__(self, X):
        raise NotImplementedError


class Zero(MeanFunction):
    def __init__(self, output_dim: int = 1):
        self.output_dim = output_dim

    def __call__(self, X):
        X = as_device(X)
        return torch.zeros((X.shape[0], self.output_dim), dtype=torch.float64, device=X.device)


class MOMeanFunction(MeanFunction):
    """ A tuple of L single-output mean functions; ``__call__`` returns the flattened (L*N,) prior mean, output-major."""

    def __init__(self, output_dim: int, mean_functions: Union['MOMeanFunction', MeanFunction, Sequence[MeanFunction]] = Zero()):
        if isinstance(mean_functions, MOMeanFunction):
            mean_functions = mean_functions.functions
        elif isinstance(mean_functions, MeanFunction):
            mean_functions = (mean_functions,) * output_dim
        self._functions = tuple(mean_functions)

    @property
    def output_dim(self):
        """ Also known as L."""
        return len(self._functions)

    @property
    def L(self):
        return self.output_dim

    @property
    def functions(self):
        return self._functions

    @property
    def is_zero(self) -> bool:
        return all(isinstance(f, Zero) for f in self._functions)

    def __call__(self, X):
        return DeviceTensor.wrap(torch.cat([f(X) for f in self._functions], dim=0).reshape(-1))
