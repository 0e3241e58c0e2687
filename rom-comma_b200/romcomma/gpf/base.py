"""``romcomma.gpf.base`` on the B200 path: the non-diagonal ``Variance`` matrix (reference romcomma/gpf/base.py:32-96).

An (L,L) symmetric positive-definite matrix held through its Cholesky factor: the diagonal is a positive parameter with lower
bound 1e-3 (softplus + shift), the strict lower triangle (row-major) a free parameter.  L x L work stays on the host;
``value_times_eye`` exists for API parity but the hot path never builds the dense (L,N,L,N) noise tensor - the gram kernel adds
E[l,l'] on the block diagonals directly (rc_gram).
"""
from __future__ import annotations

from typing import Tuple

import numpy as np

from romcomma import gf_compat as gf
from romcomma._tensors import DeviceTensor, HostTensor, as_device


class Variance(gf.Module):
    """ A non-diagonal Variance Matrix."""

    CHOLESKY_DIAGONAL_LOWER_BOUND = 1e-3

    def __init__(self, value, name: str = 'Variance', cholesky_diagonal_lower_bound: float = CHOLESKY_DIAGONAL_LOWER_BOUND):
        super().__init__(name=name)
        value = np.asarray(value.numpy() if hasattr(value, 'numpy') else value, dtype=np.float64)
        L = value.shape[-1]
        self._shape, self._broadcast_shape = (L, L), (L, 1, L, 1)
        if value.shape != self._shape:
            raise ValueError('Variance must have shape (L,L).')
        factor = np.linalg.cholesky(value)
        diagonal = np.diag(factor).copy()
        if diagonal.min() <= cholesky_diagonal_lower_bound:
            raise ValueError(f'The Cholesky diagonal of {name} must be strictly greater than {cholesky_diagonal_lower_bound}.')
        self._cholesky_diagonal = gf.Parameter(diagonal, transform=gf.positive(lower=cholesky_diagonal_lower_bound), name=name + '.cholesky_diagonal')
        self._tril = np.tril_indices(L, -1)   # row-major strict lower triangle == the gather mask of the reference (base.py:93)
        self._cholesky_lower_triangle = gf.Parameter(factor[self._tril], name=name + '.cholesky_lower_triangle')
        self._row_lengths = tuple(range(L))

    @property
    def shape(self) -> Tuple[int, int]:
        return self._shape

    def _cholesky_np(self) -> np.ndarray:
        C = np.zeros(self._shape)
        C[self._tril] = self._cholesky_lower_triangle.numpy()
        C[np.diag_indices(self._shape[0])] = self._cholesky_diagonal.numpy()
        return C

    def _value_np(self) -> np.ndarray:
        C = self._cholesky_np()
        return C @ C.T

    @property
    def cholesky(self) -> HostTensor:
        """ The (lower triangular) Cholesky decomposition of the covariance matrix."""
        return HostTensor(self._cholesky_np())

    @property
    def value(self) -> HostTensor:
        """ The covariance matrix, shape (L,L)."""
        return HostTensor(self._value_np())

    @property
    def value_to_broadcast(self) -> HostTensor:
        """ The covariance matrix, shape (L,1,L,1) ready to broadcast."""
        return HostTensor(self._value_np().reshape(self._broadcast_shape))

    def value_times_eye(self, N: int) -> DeviceTensor:
        """ variance[l,l'] * eye(N)[n,n'] as an (L,N,L,N) device tensor (API parity only; see module docstring)."""
        import torch
        L = self._shape[0]
        out = torch.zeros((L, N, L, N), dtype=torch.float64, device='cuda')
        idx = torch.arange(N, device='cuda')
        v = as_device(self._value_np())
        for l in range(L):
            for k in range(L):
                out[l, idx, k, idx] = v[l, k]
        return DeviceTensor.wrap(out)

    def _chain(self, dV: np.ndarray):
        """Gradient w.r.t. the two unconstrained parameter arrays, given dLoss/dV with the entries of V treated as independent:
        V = C C^T  =>  dLoss/dC = tril((dV + dV^T) C); the diagonal then picks up softplus' = sigmoid."""
        C = self._cholesky_np()
        dC = np.tril((dV + dV.T) @ C)
        p = self._cholesky_diagonal
        return np.diag(dC) * p.transform.dforward(p.unconstrained_variable), dC[self._tril]
