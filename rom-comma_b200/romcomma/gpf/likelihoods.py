"""``romcomma.gpf.likelihoods`` on the B200 path: the non-diagonal Gaussian likelihood (reference romcomma/gpf/likelihoods.py:34-96)."""
from __future__ import annotations

from typing import Tuple

import numpy as np
import torch

from romcomma import _capi
from romcomma import gf_compat as gf
from romcomma._tensors import DeviceTensor, as_device
from romcomma.gpf.base import Variance


class MOGaussian(gf.Module):
    """ A multivariate Gaussian likelihood whose (L,L) noise covariance acts as E (x) I_N on the (LN,LN) gram."""

    def __init__(self, variance, **kwargs):
        super().__init__(name='MOGaussian')
        self.variance = Variance(variance, name='LikelihoodVariance')
        self.latent_dim = self.observation_dim = self.variance.shape[0]

    def N(self, data) -> int:
        """ The number of samples in data, assuming the last 2 dimensions have been concatenated to LN. """
        return int(data.shape[-1] / self.latent_dim)

    def split_axis_shape(self, data) -> Tuple[int, int]:
        return self.latent_dim, self.N(data)

    def add_to(self, Fvar) -> DeviceTensor:
        """ Fvar + E (x) I for a rank-2 (LN,LN) Fvar; the dense (L,N,L,N) noise tensor of the reference is never formed."""
        Fvar = as_device(Fvar)
        if Fvar.dim() != 2:
            raise IndexError(f'mogpflow.Likelihood only accepts Fvar of rank 2 at present, provided Fvar of rank {Fvar.dim()}.')
        L, N = self.latent_dim, self.N(Fvar)
        ones = _capi.dev(np.ones((L, L)))
        return DeviceTensor.wrap(_capi.apply_variance_noise(Fvar, ones, _capi.dev(self.variance.value.numpy()), L, N))

    def _conditional_mean(self, F):
        return F

    def _conditional_variance(self, F):
        return self.variance.value_times_eye(self.N(F))

    def _predict_mean_and_var(self, Fmu, Fvar):
        Fvar = as_device(Fvar)
        E = as_device(self.variance.value.numpy())
        L = self.latent_dim
        if Fvar.dim() == 4:
            lhvar = E.reshape(1, 1, L, L)
        elif Fvar.dim() == 3:
            lhvar = E.reshape(1, L, L)
        elif Fvar.dim() == 2:
            lhvar = torch.diagonal(E).reshape(1, L)
        else:
            raise IndexError(f'Fvar has {Fvar.dim()} dimensions, when it should have 2,3, or 4.')
        return Fmu, DeviceTensor.wrap(Fvar + lhvar)

    predict_mean_and_var = _predict_mean_and_var
