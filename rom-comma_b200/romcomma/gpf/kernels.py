"""``romcomma.gpf.kernels`` on the B200 path: the multi-output ARD stationary kernel (reference romcomma/gpf/kernels.py:37-154).

K_unit[l,n,l',n'] = exp(-1/2 sum_m (X[n,m]/ls[l,m] - X2[n',m]/ls[l',m])^2);  K = reshape(F[l,1,l',1] * K_unit, (LN, LN')).
One fused CUDA kernel (rc_gram) evaluates either form; the (L,N,L,N,M) scaled-difference tensor of the reference is never built.
"""
from __future__ import annotations

from abc import abstractmethod

import numpy as np
import torch

from romcomma import _capi
from romcomma import gf_compat as gf
from romcomma._tensors import DeviceTensor, HostTensor, as_device
from romcomma.gpf.base import Variance


class MOStationary(gf.Module):
    """ Base class for multi-output stationary kernels: kernels that depend on X, X2 only through the scaled difference."""

    def __init__(self, variance, lengthscales, name='Kernel', active_dims=None):
        """
        Args:
            variance: An (L,L) symmetric, positive definite matrix for the signal variance.
            lengthscales: An (L,M) matrix of positive lengthscales (scalars / (L,) vectors are broadcast, giving M=1).
            name: The name of this kernel.
            active_dims: Which of the input dimensions are used. The default None means all of them.
        """
        super().__init__(name=name)
        self._active_dims = active_dims
        self.variance = Variance(value=np.atleast_2d(variance), name=name + 'Variance')
        self._L = self.variance.shape[0]
        lengthscales = np.asarray(lengthscales.numpy() if hasattr(lengthscales, 'numpy') else lengthscales, dtype=np.float64)
        self._M = 1 if lengthscales.shape in ((), (1,), (1, 1), (self._L,)) else lengthscales.shape[-1]
        if lengthscales.shape == (self._L,):
            lengthscales = lengthscales.reshape(self._L, 1)
        lengthscales = np.broadcast_to(lengthscales, (self._L, self._M)).reshape(self._L, 1, self._M)
        self.lengthscales = gf.Parameter(lengthscales, transform=gf.positive(), trainable=False, name=name + 'Lengthscales')

    @property
    def L(self):
        return self._L

    @property
    def M(self):
        return self._M

    @property
    def lengthscales_neat(self) -> HostTensor:
        """ The kernel lengthscales as an (L,M) matrix."""
        return HostTensor(self.lengthscales.numpy().reshape(self._L, self._M))

    def _slice(self, X) -> torch.Tensor:
        X = as_device(X)
        if X.dim() != 2:
            raise IndexError(f'MOStationary only accepts inputs X of rank 2, which X.shape={tuple(X.shape)} does not obey.')
        if self._active_dims is not None:
            X = X[:, self._active_dims].contiguous()
        return X

    def _ls_device(self, M_in: int) -> torch.Tensor:
        """(L, M_in) lengthscales on the device (an isotropic kernel, M=1, is broadcast over the input columns)."""
        return _capi.dev(np.ascontiguousarray(np.broadcast_to(self.lengthscales_neat.numpy(), (self._L, M_in))))

    def _gram(self, X, X2, with_variance: bool) -> torch.Tensor:
        X = self._slice(X)
        X2 = None if X2 is None else self._slice(X2)
        F = _capi.dev(self.variance.value.numpy()[None]) if with_variance else None
        N, N2 = X.shape[0], (X.shape[0] if X2 is None else X2.shape[0])
        K = _capi.gram(X, X2, self._ls_device(X.shape[1]), F, None)
        return K[0, :self._L * N, :self._L * N2]

    def K_unit_variance(self, X, X2=None) -> DeviceTensor:
        """ The kernel with variance=ones(), shape (L,N,L,N2). This can be cached when only the variance is trainable."""
        X = self._slice(X)
        N, N2 = X.shape[0], (X.shape[0] if X2 is None else np.shape(X2)[0])
        return DeviceTensor.wrap(self._gram(X, X2, False).reshape(self._L, N, self._L, N2))

    @abstractmethod
    def K_d_unit_variance(self, d):
        """ The unit-variance kernel as a function of the (L,N,L,N,M) scaled difference (kept for API parity; small inputs only)."""
        raise NotImplementedError(f'You must implement K_d_unit_variance(self, d) in {type(self)}.')

    def K_d_apply_variance(self, K_d_unit_variance) -> DeviceTensor:
        """ Multiply the (L,N,L,N2) unit-variance kernel by the kernel variance and reshape to (LN, LN2)."""
        Ku = as_device(K_d_unit_variance)
        if Ku.dim() != 4:
            raise IndexError(f'MOStationary only accepts K_d_unit_variance of rank 4, which shape={tuple(Ku.shape)} does not obey.')
        L, N, _, N2 = Ku.shape
        if N != N2:
            raise IndexError('K_d_apply_variance on the device path expects a square (L,N,L,N) unit gram.')
        F = _capi.dev(self.variance.value.numpy())
        return DeviceTensor.wrap(_capi.apply_variance_noise(Ku.reshape(L * N, L * N2), F, None, L, N))

    def K_d(self, d) -> DeviceTensor:
        return self.K_d_apply_variance(self.K_d_unit_variance(d))

    def K(self, X, X2=None) -> DeviceTensor:
        return DeviceTensor.wrap(self._gram(X, X2, True))

    def K_diag(self, X) -> DeviceTensor:
        """ The reference's K_diag calls a method that does not exist and is unreachable; here: diag of K(X,X)."""
        N = as_device(X).shape[0]
        return DeviceTensor.wrap(as_device(np.repeat(np.diag(self.variance.value.numpy()), N)))

    def __call__(self, X, X2=None, *, full_cov=True, presliced=False) -> DeviceTensor:
        return self.K(X, X2)


class RBF(MOStationary):
    """ The radial basis function or squared exponential kernel,  k(d) = variance * exp(-1/2 |d|^2)."""

    def K_d_unit_variance(self, d) -> DeviceTensor:
        d = as_device(d)
        return DeviceTensor.wrap(torch.exp(-0.5 * torch.einsum('...M,...M->...', d, d)))
