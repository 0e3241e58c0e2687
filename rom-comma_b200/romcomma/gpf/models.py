"""``romcomma.gpf.models`` on the B200 path: multi-output GP regression (reference romcomma/gpf/models.py:33-139).

LML = log N(vec(Y^T) | 0, F (x) K_unit + E (x) I): the gram, the FP64 Cholesky, the solve, the log-determinant and - for the
optimiser - the analytic gradient all run inside one C-ABI call (rc_lml_grad).  ``predict_f`` computes the mean and the marginal
variances through rc_gram / rc_potrf / rc_trsm_fwd / rc_predict_reduce; the full (L n*)^2 predictive covariance that the
reference builds and then discards (models.py:97-109) is not formed.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch

import romcomma.gpf as mf
from romcomma import _capi, lockstep
from romcomma import gf_compat as gf
from romcomma._tensors import DeviceTensor, HostTensor, as_device


class MOGPR(gf.Module):
    """ Gaussian process regression with L correlated outputs and a non-diagonal Gaussian likelihood."""

    def __init__(self, data, kernel: 'mf.kernels.MOStationary', mean_function: Optional['mf.mean_functions.MOMeanFunction'] = None,
                 noise_variance=1.0):
        """
        Args:
            data: (X, Y) with X of shape (N,M) and Y of shape (N,L).
            kernel: Must be well-formed, with an (L,L) variance and an (L,M) lengthscales matrix.
            mean_function: Defaults to Zero (the only mean on the device path).
            noise_variance: Broadcast to (L,L) and reduced to its diagonal, exactly as the reference constructor does.
        """
        super().__init__(name='MOGPR')
        X, Y = data
        self._X, self._Yd = as_device(X), as_device(Y)
        if self._X.dim() != 2:
            raise IndexError(f'X should be of rank 2 instead of {self._X.dim()}.')
        self._N, self._M = self._X.shape
        self._L = self._Yd.shape[-1]
        if tuple(self._Yd.shape) != (self._N, self._L):
            raise IndexError(f'Y.shape should be {(self._N, self._L)} instead of {tuple(self._Yd.shape)}.')
        self.data = (DeviceTensor.wrap(self._X), DeviceTensor.wrap(self._Yd))
        self._Y = DeviceTensor.wrap(self._Yd.T.reshape(-1, 1))     # (LN,1), output-major
        noise = np.asarray(noise_variance.numpy() if hasattr(noise_variance, 'numpy') else noise_variance, dtype=np.float64)
        noise = np.diag(np.diag(np.broadcast_to(noise, (self._L, self._L))))
        self.kernel = kernel
        self.likelihood = mf.likelihoods.MOGaussian(noise)
        self.mean_function = mf.mean_functions.MOMeanFunction(self._L) if mean_function is None else mean_function
        if not self.mean_function.is_zero:
            raise NotImplementedError('only the zero mean function is implemented on the device path (the reference never uses another).')
        self.num_latent_gps = 1
        self._K_unit_cache = None
        self._plans = {}

    @property
    def M(self):
        """ The input dimensionality."""
        return self._M

    @property
    def L(self):
        """ The output dimensionality."""
        return self._L

    @property
    def _K_unit_variance(self) -> DeviceTensor:
        """ The cached unit-variance gram (L,N,L,N); the reference builds it eagerly in its constructor (models.py:139)."""
        if self._K_unit_cache is None:
            self._K_unit_cache = self.kernel.K_unit_variance(self._X)
        return self._K_unit_cache

    @property
    def KXX(self) -> DeviceTensor:
        return self.kernel(self._X, self._X) if self.kernel.lengthscales.trainable else self.kernel.K_d_apply_variance(self._K_unit_variance)

    # -- LML (+ gradient) ------------------------------------------------------------------------------------------
    def _plan(self, flags: int) -> _capi.LmlGradPlan:
        if flags not in self._plans:
            self._plans.clear()
            self._plans[flags] = _capi.LmlGradPlan(self._X, self._Yd, self._L, 1, flags)
        return self._plans[flags]

    def _evaluate_device(self, flags: int) -> torch.Tensor:
        """Launch one LML(+grad) evaluation; returns the device result vector { lml, dF, dE, dls } without synchronising."""
        plan = self._plan(flags)
        ls = self.kernel._ls_device(self._M)
        F = _capi.dev(self.kernel.variance.value.numpy()[None])
        E = _capi.dev(self.likelihood.variance.value.numpy()[None])
        return plan(ls, F, E)

    def _evaluate(self, flags: int) -> dict:
        broker = lockstep.current()
        if broker is not None:     # a lock-step session (folds fitted side by side): join the batch; the selected inverse is a batch-of-one path
            return broker.evaluate(self, self._X, self._Yd, self._L, flags & ~_capi.RC_GRAD_F_DIAGONAL, np.broadcast_to(self.kernel.lengthscales_neat.numpy(), (self._L, self._M)),
                                   self.kernel.variance.value.numpy(), self.likelihood.variance.value.numpy())
        out = self._evaluate_device(flags).cpu().numpy()
        plan = self._plan(flags)
        if int(plan.info.cpu()[0]) != 0:
            raise _capi.RomcommaB200Error('Cholesky decomposition was not successful. The input might not be valid.')
        return plan.unpack(out)[0]

    def log_marginal_likelihood(self) -> HostTensor:
        return HostTensor(self._evaluate(_capi.RC_GRAD_NONE)['lml'])

    def maximum_log_likelihood_objective(self) -> HostTensor:
        return self.log_marginal_likelihood()

    def training_loss(self) -> HostTensor:
        return HostTensor(-self.log_marginal_likelihood())

    def _loss_and_grad(self, variables: Sequence[gf.Parameter]):
        """(-LML, gradients w.r.t. the unconstrained ``variables``): what tf.GradientTape gives gpflow's Scipy in the reference."""
        kv, lv = self.kernel.variance, self.likelihood.variance
        want_ls = any(v is self.kernel.lengthscales for v in variables)
        flags = _capi.RC_GRAD_VARIANCE | (_capi.RC_GRAD_LENGTHSCALES if want_ls else 0)
        F = kv.value.numpy()
        if not any(v is kv._cholesky_lower_triangle for v in variables) and not np.any(F - np.diag(np.diag(F))):
            flags |= _capi.RC_GRAD_F_DIAGONAL     # default trainables (gpr/kernels.py:54-57): only diag(F) is trained and F is diagonal
        res = self._evaluate(flags)
        dFd, dFl = kv._chain(res['dF'])
        dEd, dEl = lv._chain(res['dE'])
        grads = []
        for v in variables:
            if v is kv._cholesky_diagonal:
                g = dFd
            elif v is kv._cholesky_lower_triangle:
                g = dFl
            elif v is lv._cholesky_diagonal:
                g = dEd
            elif v is lv._cholesky_lower_triangle:
                g = dEl
            elif v is self.kernel.lengthscales:
                dls = res['dls'] if self.kernel.M == self._M else res['dls'].sum(axis=1, keepdims=True)
                g = dls.reshape(v.shape) * v.transform.dforward(v.unconstrained_variable)
            else:
                raise ValueError(f'{v.name} is not a parameter of this model.')
            grads.append(-np.asarray(g, dtype=np.float64))
        return -res['lml'], grads

    # -- prediction ------------------------------------------------------------------------------------------------
    def _predict(self, Xnew, full_cov, full_output_cov, y_instead_of_f):
        Xn = as_device(Xnew).reshape(-1, self._M)
        if full_cov or full_output_cov:
            if y_instead_of_f:       # inherited gpflow GPModel.predict_y (2.5.2) refuses these arguments, see GPflow issue 1461
                raise NotImplementedError('The predict_y method currently supports only the argument values full_cov=False and full_output_cov=False')
            mean, cov = gf.predict_full_core(self._X, self._Yd, self.kernel._ls_device(self._M), self.kernel.variance.value.numpy()[None],
                                             self.likelihood.variance.value.numpy()[None], Xn, self._L, 1)
            return DeviceTensor.wrap(mean[0]), DeviceTensor.wrap(gf.shape_full_covariance(cov[0], self._L, Xn.shape[0], bool(full_cov)))
        mean, var = gf.predict_core(self._X, self._Yd, self.kernel._ls_device(self._M), self.kernel.variance.value.numpy()[None],
                                    self.likelihood.variance.value.numpy()[None], Xn, self._L, 1, y_instead_of_f)
        return DeviceTensor.wrap(mean[0]), DeviceTensor.wrap(var[0])

    def predict_f(self, Xnew, full_cov: bool = False, full_output_cov: bool = False):
        """ Mean (n, L) of f at Xnew and its variance: marginal (n, L) by default; (n, L, L) with full_output_cov; (n, n, L, L) with full_cov, which
        implies full_output_cov - the shapes of the reference (gpf/models.py:94-109)."""
        return self._predict(Xnew, full_cov, full_output_cov, False)

    def predict_y(self, Xnew, full_cov: bool = False, full_output_cov: bool = False):
        """ As predict_f, with diag(E) added to the variance (MOGaussian._predict_mean_and_var, rank-2 branch)."""
        return self._predict(Xnew, full_cov, full_output_cov, True)
