"""The sliver of the gpflow API that rom-comma's hot path touches, re-hosted on the B200 C-ABI.

The reference builds on gpflow (pinned >=2.2.1,<=2.5.2, reference pyproject.toml:36) for parameters/transforms, the variant
model ``gf.models.GPR(gf.kernels.RBF, gf.likelihoods.Gaussian)`` (romcomma/gpr/models.py:340-342, gpr/kernels.py:176-177) and
``gf.optimizers.Scipy`` (gpr/models.py:359-361).  gpflow is not a dependency here: this module restates those pieces -
``Parameter`` with ``positive()`` = softplus (+ shift) transforms, ``set_trainable``, attribute-name-ordered
``trainable_variables`` (tf.Module semantics), the L-BFGS-B driver - and routes every FLOP-carrying call to
``librc_b200.so``.  Gradients are analytic, not taped.

Use as ``from romcomma import gf_compat as gf``.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Any, Callable, Iterable, Optional, Sequence, Tuple

import numpy as np
import scipy.optimize
import torch

from romcomma import _capi, lockstep
from romcomma._tensors import DeviceTensor, HostTensor, as_device


# ------------------------------------------------------------------------------------------------------------------
# config
# ------------------------------------------------------------------------------------------------------------------
class _Config:
    _float, _int = np.float64, np.int32

    @staticmethod
    def default_float():
        return _Config._float

    @staticmethod
    def default_int():
        return _Config._int

    @staticmethod
    def set_default_float(value):
        if np.dtype(value) != np.float64:
            raise ValueError('the B200 path computes in float64 only (romcomma/user/contexts.py:67 forces float64 as well).')

    @staticmethod
    def set_default_int(value):
        _Config._int = np.dtype(value).type


config = _Config()


# ------------------------------------------------------------------------------------------------------------------
# transforms and parameters
# ------------------------------------------------------------------------------------------------------------------
def _softplus(u):
    return np.logaddexp(0.0, u)


def _softplus_inverse(y):
    return y + np.log(-np.expm1(-y))


def _sigmoid(u):
    return 0.5 * (1.0 + np.tanh(0.5 * u))


class Transform:
    """y = softplus(u) + lower   (gpflow.utilities.positive); identity if ``positive`` is False."""

    def __init__(self, positive: bool = False, lower: float = 0.0):
        self.positive, self.lower = positive, float(lower)

    def forward(self, u):
        return _softplus(u) + self.lower if self.positive else np.asarray(u, dtype=np.float64)

    def inverse(self, y):
        return _softplus_inverse(np.asarray(y, dtype=np.float64) - self.lower) if self.positive else np.asarray(y, dtype=np.float64)

    def dforward(self, u):
        """dy/du, elementwise."""
        return _sigmoid(u) if self.positive else np.ones_like(u)


def positive(lower: Optional[float] = None) -> Transform:
    return Transform(True, 0.0 if lower is None else lower)


class Parameter:
    """A constrained value held through its unconstrained representation (gpflow.Parameter)."""

    def __init__(self, value, transform: Optional[Transform] = None, trainable: bool = True, name: str = 'Parameter'):
        self.transform = transform or Transform()
        value = np.array(value.numpy() if hasattr(value, 'numpy') else value, dtype=np.float64)
        if self.transform.positive and np.any(value <= self.transform.lower):
            raise ValueError(f'{name}: value must exceed the lower bound {self.transform.lower}.')
        self.unconstrained_variable = self.transform.inverse(value)
        self.trainable, self.name = trainable, name

    @property
    def shape(self):
        return self.unconstrained_variable.shape

    def numpy(self) -> np.ndarray:
        return self.transform.forward(self.unconstrained_variable)

    def value(self) -> HostTensor:
        return HostTensor(self.numpy())

    def assign(self, value):
        value = np.broadcast_to(np.asarray(value.numpy() if hasattr(value, 'numpy') else value, dtype=np.float64), self.shape)
        if self.transform.positive and np.any(value <= self.transform.lower):
            raise ValueError(f'{self.name}: value must exceed the lower bound {self.transform.lower}.')
        self.unconstrained_variable = self.transform.inverse(value)

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        return a.astype(dtype) if dtype is not None else a

    def __getitem__(self, item):
        return HostTensor(self.numpy()[item])

    def __repr__(self):
        return f'<Parameter {self.name} shape={self.shape} trainable={self.trainable} value={self.numpy()!r}>'


class Module:
    """Attribute-name-ordered parameter discovery, as tf.Module does (sorted ``vars(self)``, depth first)."""

    def __init__(self, name: Optional[str] = None):
        self.name = name or type(self).__name__

    def _flatten_parameters(self) -> Iterable[Parameter]:
        seen = set()

        def walk(obj):
            if isinstance(obj, Parameter):
                if id(obj) not in seen:
                    seen.add(id(obj))
                    yield obj
            elif isinstance(obj, Module):
                for key in sorted(vars(obj)):
                    yield from walk(vars(obj)[key])
            elif isinstance(obj, (list, tuple)):
                for item in obj:
                    yield from walk(item)

        yield from walk(self)

    @property
    def parameters(self) -> Tuple[Parameter, ...]:
        return tuple(self._flatten_parameters())

    @property
    def trainable_parameters(self) -> Tuple[Parameter, ...]:
        return tuple(p for p in self._flatten_parameters() if p.trainable)

    @property
    def trainable_variables(self) -> Tuple[Parameter, ...]:
        """The reference hands these to the optimizer; here a 'variable' is the Parameter holding the unconstrained array."""
        return self.trainable_parameters


def set_trainable(model, flag: bool):
    if isinstance(model, Parameter):
        model.trainable = bool(flag)
    else:
        for p in model.parameters:
            p.trainable = bool(flag)


# ------------------------------------------------------------------------------------------------------------------
# the variant model: gpflow GPR with an ARD squared-exponential kernel and a Gaussian likelihood
# ------------------------------------------------------------------------------------------------------------------
class _RBF(Module):
    """gpflow.kernels.RBF(variance, lengthscales): k = variance * exp(-1/2 |(x - x')/lengthscales|^2)."""

    def __init__(self, variance=1.0, lengthscales=1.0, name='RBF'):
        super().__init__(name)
        self.variance = Parameter(variance, transform=positive(), name='variance')
        self.lengthscales = Parameter(lengthscales, transform=positive(), name='lengthscales')

    def _ls_row(self, M: int) -> np.ndarray:
        return np.ascontiguousarray(np.broadcast_to(np.atleast_1d(self.lengthscales.numpy()), (M,)), dtype=np.float64).reshape(1, M)

    def __call__(self, X, X2=None, *, full_cov=True, presliced=False) -> DeviceTensor:
        X = as_device(X)
        M = X.shape[1]
        if not full_cov:
            return DeviceTensor.wrap(torch.full((X.shape[0],), float(self.variance.numpy()), dtype=torch.float64, device=X.device))
        X2 = None if X2 is None else as_device(X2)
        N, N2 = X.shape[0], (X.shape[0] if X2 is None else X2.shape[0])
        F = _capi.dev(np.reshape(self.variance.numpy(), (1, 1, 1)))
        K = _capi.gram(X, X2, _capi.dev(self._ls_row(M)), F, None)
        return DeviceTensor.wrap(K[0, :N, :N2])

    K = __call__


class _Gaussian(Module):
    """gpflow.likelihoods.Gaussian: variance with lower bound 1e-6."""
    DEFAULT_VARIANCE_LOWER_BOUND = 1e-6

    def __init__(self, variance=1.0, variance_lower_bound=DEFAULT_VARIANCE_LOWER_BOUND, name='Gaussian'):
        super().__init__(name)
        if variance <= variance_lower_bound:
            raise ValueError(f'The variance of the Gaussian likelihood must be strictly greater than {variance_lower_bound}')
        self.variance = Parameter(variance, transform=positive(lower=variance_lower_bound), name='variance')


class _GPR(Module):
    """gpflow.models.GPR on the B200 path (one output column)."""

    def __init__(self, data, kernel: _RBF, mean_function=None, noise_variance: float = 1.0, name='GPR'):
        super().__init__(name)
        X, Y = data
        self._Xd, self._Yd = as_device(X), as_device(Y)
        if self._Yd.dim() != 2 or self._Yd.shape[1] != 1:
            raise ValueError('gf.models.GPR on this path takes a single output column (the reference builds one GPR per output).')
        self.data = (DeviceTensor.wrap(self._Xd), DeviceTensor.wrap(self._Yd))
        self.kernel = kernel
        self.likelihood = _Gaussian(noise_variance)
        self.mean_function = mean_function
        self.num_latent_gps = 1
        self._plans = {}

    # -- LML (+ gradient) ------------------------------------------------------------------------------------------
    def _plan(self, flags: int) -> _capi.LmlGradPlan:
        if flags not in self._plans:
            self._plans.clear()   # one resident workspace at a time
            self._plans[flags] = _capi.LmlGradPlan(self._Xd, self._Yd, 1, 1, flags)
        return self._plans[flags]

    def _evaluate(self, flags: int) -> dict:
        M = self._Xd.shape[1]
        broker = lockstep.current()
        if broker is not None:     # a lock-step session: this evaluation joins the batch of the optimisers running beside this one
            return broker.evaluate(self, self._Xd, self._Yd, 1, flags, self.kernel._ls_row(M), np.reshape(self.kernel.variance.numpy(), (1, 1)),
                                   np.reshape(self.likelihood.variance.numpy(), (1, 1)))
        plan = self._plan(flags)
        out = plan(_capi.dev(self.kernel._ls_row(M)), _capi.dev(np.reshape(self.kernel.variance.numpy(), (1, 1, 1))),
                   _capi.dev(np.reshape(self.likelihood.variance.numpy(), (1, 1, 1))))
        host = out.cpu().numpy()
        if int(plan.info.cpu()[0]) != 0:
            raise _capi.RomcommaB200Error('Cholesky decomposition was not successful. The input might not be valid.')
        return plan.unpack(host)[0]

    def log_marginal_likelihood(self) -> HostTensor:
        return HostTensor(self._evaluate(_capi.RC_GRAD_NONE)['lml'])

    def maximum_log_likelihood_objective(self) -> HostTensor:
        return self.log_marginal_likelihood()

    def training_loss(self) -> HostTensor:
        return HostTensor(-self.log_marginal_likelihood())

    def _loss_and_grad(self, variables: Sequence[Parameter]):
        """(-LML, d(-LML)/d(unconstrained variables)) in the order of ``variables``."""
        want_ls = any(v is self.kernel.lengthscales for v in variables)
        res = self._evaluate(_capi.RC_GRAD_VARIANCE | (_capi.RC_GRAD_LENGTHSCALES if want_ls else 0))
        grads = []
        for v in variables:
            if v is self.kernel.lengthscales:
                g = res['dls'][0]
                g = g if v.shape == g.shape else np.reshape(np.sum(g), v.shape)   # isotropic: one shared lengthscale
            elif v is self.kernel.variance:
                g = np.reshape(res['dF'][0, 0], v.shape)
            elif v is self.likelihood.variance:
                g = np.reshape(res['dE'][0, 0], v.shape)
            else:
                raise ValueError(f'{v.name} is not a parameter of this model.')
            grads.append(-g * v.transform.dforward(v.unconstrained_variable))
        return -res['lml'], grads

    # -- prediction ------------------------------------------------------------------------------------------------
    def _predict(self, Xnew, full_cov, full_output_cov, y_instead_of_f):
        M = self._Xd.shape[1]
        if full_cov or full_output_cov:
            if y_instead_of_f:       # gpflow GPModel.predict_y (2.5.2) refuses these arguments, see GPflow issue 1461
                raise NotImplementedError('The predict_y method currently supports only the argument values full_cov=False and full_output_cov=False')
            Xn = as_device(Xnew)
            mean, cov = predict_full_core(self._Xd, self._Yd, _capi.dev(self.kernel._ls_row(M)), np.reshape(self.kernel.variance.numpy(), (1, 1, 1)),
                                          np.reshape(self.likelihood.variance.numpy(), (1, 1, 1)), Xn, 1, 1)
            # gpflow GPR.predict_f: full_cov -> [1, n, n]; full_output_cov alone -> [n, 1, 1]
            var = cov if full_cov else torch.diagonal(cov[0]).reshape(-1, 1, 1)
            return DeviceTensor.wrap(mean[0]), DeviceTensor.wrap(var)
        mean, var = predict_core(self._Xd, self._Yd, _capi.dev(self.kernel._ls_row(M)), np.reshape(self.kernel.variance.numpy(), (1, 1, 1)),
                                 np.reshape(self.likelihood.variance.numpy(), (1, 1, 1)), as_device(Xnew), 1, 1, y_instead_of_f)
        return DeviceTensor.wrap(mean[0]), DeviceTensor.wrap(var[0])

    def predict_f(self, Xnew, full_cov: bool = False, full_output_cov: bool = False):
        return self._predict(Xnew, full_cov, full_output_cov, False)

    def predict_y(self, Xnew, full_cov: bool = False, full_output_cov: bool = False):
        return self._predict(Xnew, full_cov, full_output_cov, True)


def predict_core(X, Y, ls, F, E, Xn, L: int, batch: int, y_instead_of_f: bool = False, fac=None):
    """Shared by the variant and covariant models. Returns mean and variance, each (batch, n*, L), on the device.

    mean = Kmn^T K^-1 y and var = diag(Knn) - colsumsq(L^-1 Kmn) (+ diag E): gpflow ``base_conditional`` (called at
    romcomma/gpf/models.py:97 and inside gpflow GPR.predict_f) without the full (L n*)^2 covariance.
    F, E are host arrays (batch, L, L).  ``fac``: an existing factorisation of the noisy gram at exactly these hyper-parameters
    (romcomma.gpr.models.MOGP keeps one per fitted GP); without it the gram is built and factorised here, as the reference does per call.
    """
    N, nstar = X.shape[0], Xn.shape[0]
    n, c = L * N, L * nstar
    n_pad, c_pad = _capi.padded(n), _capi.padded(c)
    F, E = np.asarray(F, dtype=np.float64).reshape(batch, L, L), np.asarray(E, dtype=np.float64).reshape(batch, L, L)
    dF, dE = _capi.dev(F), _capi.dev(E)
    if fac is None:
        Kmm = _capi.gram(X, None, ls, dF, dE, batch=batch, lower_only=True, pad_to=n, pad_identity=True)
        fac = _capi.Factorization(Kmm)
        fac.raise_if_failed()
    assert fac.batch == batch and fac.n_pad == n_pad
    y = torch.zeros((batch, n_pad), dtype=torch.float64, device=X.device)
    y[:, :n] = Y.reshape(N, batch, L).permute(1, 2, 0).reshape(batch, n)          # y = vec(Y^T) per problem (gpf/models.py:130)
    alpha = fac.trsv(y)
    Kmn = torch.empty((batch, n_pad, c_pad), dtype=torch.float64, device=X.device)
    _capi.check(_capi.lib().rc_gram(_capi.ptr(X), N, _capi.ptr(Xn), nstar, X.shape[1], _capi.ptr(ls), L, _capi.ptr(dF), None, _capi.ptr(Kmn), c_pad,
                                    n_pad * c_pad, n_pad, c_pad, 0, 0, batch, _capi.stream_ptr()), 'rc_gram')
    fac.trsm_fwd_(Kmn)
    kdiag = _capi.dev(np.ascontiguousarray(np.diagonal(F, axis1=1, axis2=2)))
    noise = _capi.dev(np.ascontiguousarray(np.diagonal(E, axis1=1, axis2=2))) if y_instead_of_f else None
    return _capi.predict_reduce(Kmn, alpha, L, nstar, kdiag, noise)


def predict_full_core(X, Y, ls, F, E, Xn, L: int, batch: int, fac=None):
    """Mean (batch, n*, L) and the FULL covariance Knn - A^T A of f at Xn, (batch, L n*, L n*) with index (l, i) -> l n* + i: gpflow
    ``base_conditional(full_cov=True)`` as MOGPR.predict_f calls it (romcomma/gpf/models.py:97).  A = L^-1 Kmn by rc_trsm_fwd, A^T A on the
    tensor-core tile kernel (rc_syrk_tn, accumulated straight onto the gram of the new points)."""
    N, nstar = X.shape[0], Xn.shape[0]
    n, c = L * N, L * nstar
    n_pad, c_pad = _capi.padded(n), _capi.padded(c)
    F, E = np.asarray(F, dtype=np.float64).reshape(batch, L, L), np.asarray(E, dtype=np.float64).reshape(batch, L, L)
    dF, dE = _capi.dev(F), _capi.dev(E)
    if fac is None:
        fac = _capi.Factorization(_capi.gram(X, None, ls, dF, dE, batch=batch, lower_only=True, pad_to=n, pad_identity=True))
        fac.raise_if_failed()
    y = torch.zeros((batch, n_pad), dtype=torch.float64, device=X.device)
    y[:, :n] = Y.reshape(N, batch, L).permute(1, 2, 0).reshape(batch, n)
    alpha = fac.trsv(y)
    Kmn = torch.empty((batch, n_pad, c_pad), dtype=torch.float64, device=X.device)
    _capi.check(_capi.lib().rc_gram(_capi.ptr(X), N, _capi.ptr(Xn), nstar, X.shape[1], _capi.ptr(ls), L, _capi.ptr(dF), None, _capi.ptr(Kmn), c_pad,
                                    n_pad * c_pad, n_pad, c_pad, 0, 0, batch, _capi.stream_ptr()), 'rc_gram')
    fac.trsm_fwd_(Kmn)
    kdiag = _capi.dev(np.ascontiguousarray(np.diagonal(F, axis1=1, axis2=2)))
    mean, _ = _capi.predict_reduce(Kmn, alpha, L, nstar, kdiag, None)
    cov = _capi.gram(Xn, None, ls, dF, None, batch=batch, pad_to=c)                  # Knn, zero padding
    _capi.syrk_tn(Kmn, alpha=-1.0, beta=1.0, C=cov)
    return mean, cov[:, :c, :c]


def shape_full_covariance(cov: torch.Tensor, L: int, nstar: int, full_cov: bool) -> torch.Tensor:
    """(L n*, L n*) -> what MOGPR.predict_f returns (romcomma/gpf/models.py:99-109): 'LNln -> LlNn', the diagonal over (N, n) unless full_cov,
    axes reversed: (n*, n*, L, L) for full_cov, (n*, L, L) for full_output_cov alone."""
    four = cov.reshape(L, nstar, L, nstar).permute(0, 2, 1, 3)            # [L, l, N, n]
    if not full_cov:
        return torch.diagonal(four, dim1=2, dim2=3).permute(2, 1, 0).contiguous()
    return four.permute(3, 2, 1, 0).contiguous()


# ------------------------------------------------------------------------------------------------------------------
# optimizer
# ------------------------------------------------------------------------------------------------------------------
class _Scipy:
    """gpflow.optimizers.Scipy: pack the unconstrained variables into one float64 vector and hand loss+gradient to
    scipy.optimize.minimize(jac=True).  The closure must be a bound ``training_loss`` of a model of this package (its
    analytic ``_loss_and_grad`` replaces the tf.GradientTape of the reference)."""

    @staticmethod
    def _model_of(closure: Callable):
        model = getattr(closure, '__self__', None)
        if model is None or not hasattr(model, '_loss_and_grad'):
            raise TypeError('Scipy needs closure=model.training_loss of a romcomma B200 model: gradients are analytic, not taped.')
        return model

    @staticmethod
    def initial_parameters(variables: Sequence[Parameter]) -> np.ndarray:
        """ gpflow.optimizers.Scipy.initial_parameters: the unconstrained variables packed into one float64 vector."""
        return np.concatenate([np.reshape(v.unconstrained_variable, -1) for v in variables]).astype(np.float64)

    @staticmethod
    def assign_tensors(variables: Sequence[Parameter], x: np.ndarray):
        offset = 0
        for v in variables:
            size = int(np.prod(v.shape))
            v.unconstrained_variable = np.array(x[offset:offset + size], dtype=np.float64).reshape(v.shape)
            offset += size

    @classmethod
    def eval_func(cls, closure: Callable, variables: Sequence[Parameter], compile: bool = True, allow_unused_variables: bool = False):
        """ gpflow.optimizers.Scipy.eval_func: the function x -> (loss, gradient) that ``minimize`` hands to scipy - ONE evaluation of the LML
        and its analytic gradient per call."""
        model, variables = cls._model_of(closure), tuple(variables)

        def fun(x):
            cls.assign_tensors(variables, x)
            loss, grads = model._loss_and_grad(variables)
            return float(loss), np.concatenate([np.reshape(g, -1) for g in grads]).astype(np.float64)
        return fun

    def minimize(self, closure: Callable, variables: Sequence[Parameter], method: Optional[str] = 'L-BFGS-B', step_callback=None, compile=True,
                 allow_unused_variables=False, **scipy_kwargs) -> scipy.optimize.OptimizeResult:
        variables = tuple(variables)
        if not variables:
            raise ValueError('no trainable variables')
        fun = self.eval_func(closure, variables)
        unpack = lambda x: self.assign_tensors(variables, x)
        x0 = self.initial_parameters(variables)
        callback = None
        if step_callback is not None:
            counter = [0]

            def callback(x):
                step_callback(counter[0], variables, [x])
                counter[0] += 1
        with lockstep.participating():      # inside lockstep.run_together: the evaluations of concurrently running fits are batched
            result = scipy.optimize.minimize(fun, x0, jac=True, method=method, callback=callback, **scipy_kwargs)
        unpack(result.x)
        return result


kernels = SimpleNamespace(RBF=_RBF, SquaredExponential=_RBF)
likelihoods = SimpleNamespace(Gaussian=_Gaussian)
models = SimpleNamespace(GPR=_GPR)
optimizers = SimpleNamespace(Scipy=_Scipy)
utilities = SimpleNamespace(positive=positive)
