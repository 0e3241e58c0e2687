""" The storage-backed Gaussian process ``MOGP`` (reference romcomma/gpr/models.py:35-463) on the B200 path.

``implementation`` is L independent single-output models (variant) or one ``mf.models.MOGPR`` (covariant); ``calibrate`` runs
L-BFGS-B on the unconstrained hyper-parameters with the analytic gradient from rc_lml_grad; ``predict``, ``K_cho`` and ``K_inv_Y``
run on the device and hand back device tensors.  Folder layout (kernel/, likelihood/, meta.json, test.csv, test_summary.csv) is the
reference's."""
from __future__ import annotations

from romcomma.base.definitions import *
from romcomma.data.storage import Fold, Frame
from romcomma.base.classes import Data, Model
from romcomma.gpr.kernels import Kernel
from romcomma import _capi, lockstep
from romcomma._tensors import DeviceTensor, as_device


class Likelihood(Model):

    class Data(Data):
        class NamedTuple(NamedTuple):
            """ variance: (L,L), (1,L) or (1,1) noise variance; (1,L) represents a diagonal (L,L) matrix.
                log_marginal: output only - records the log marginal likelihood."""
            variance: Any = np.atleast_2d(0.02)
            log_marginal: Any = np.atleast_2d(1.0)

    @classproperty
    def META(cls) -> Dict[str, Any]:
        return {'variance': True, 'covariance': True}

    @classproperty
    def VARIANCE_FLOOR(cls) -> float:
        return 1.0001E-6

    def __init__(self, parent: 'GPR', read_data: bool = False, **kwargs):
        super().__init__(parent.folder / 'likelihood', read_data, **kwargs)
        self._parent = parent

    @property
    def is_covariant(self) -> bool:
        return self._data.frames.variance.df.shape[0] > 1

    def calibrate(self, **kwargs) -> Dict[str, Any]:
        """ Merely sets the trainable parameters."""
        meta = self.META | kwargs
        if self.is_covariant:
            gf.set_trainable(self._parent._implementation[0].likelihood.variance._cholesky_diagonal, meta['variance'])
            gf.set_trainable(self._parent._implementation[0].likelihood.variance._cholesky_lower_triangle, meta['covariance'])
        else:
            for implementation in self._parent.implementation:
                gf.set_trainable(implementation.likelihood.variance, meta['variance'])
        return meta


# noinspection PyPep8Naming
class GPR(Model):
    """ Interface to a Gaussian Process."""

    class Data(Data):
        class NamedTuple(NamedTuple):
            """ kernel: a [[str]] identifying the type of Kernel (Kernel.TYPE_IDENTIFIER); never set externally."""
            kernel: Any = np.atleast_2d(None)

    @classproperty
    def KERNEL_FOLDER_NAME(cls) -> str:
        return 'kernel'

    def __init__(self, name: str, fold: Fold, is_read: bool | None, is_covariant: bool, is_isotropic: bool,
                 kernel_parameters: Kernel.Data | None = None, likelihood_variance: NP.Matrix | None = None):
        """
        Args:
            name: The name of this MOGP (a sub-folder of the fold).
            fold: The Fold housing this MOGP.
            is_read: If True the kernel and likelihood data are read from ``fold.folder/name``, otherwise defaults are used.
            is_covariant: Whether the outputs are treated as dependent.
            is_isotropic: Whether to restrict the kernel to be isotropic.
            kernel_parameters: A Kernel.Data to use instead of file/defaults.
            likelihood_variance: The likelihood variance to use instead of file/defaults.
        """
        self._fold = fold
        self._X, self._Y = self._fold.X.to_numpy(dtype=FLOAT(), copy=True), self._fold.Y.to_numpy(dtype=FLOAT(), copy=True)
        self._N, self._M, self._L = self._fold.N, self._fold.M, self._fold.L
        super().__init__(self._fold.folder / name, is_read)
        self._likelihood = Likelihood(self, is_read) if likelihood_variance is None else Likelihood(self, is_read, variance=likelihood_variance)
        if is_read and kernel_parameters is None:
            KernelType = Kernel.TypeFromIdentifier(self.data.frames.kernel.np[0, 0])
            self._kernel = KernelType(self._folder / self.KERNEL_FOLDER_NAME, is_read)
        else:
            if kernel_parameters is None:
                from romcomma.gpr import kernels
                kernel_parameters = kernels.RBF.Data(self._folder / self.KERNEL_FOLDER_NAME)
            KernelType = Kernel.TypeFromParameters(kernel_parameters)
            self._kernel = KernelType(self._folder / self.KERNEL_FOLDER_NAME, is_read, **kernel_parameters.asdict())
            self._data.replace(kernel=np.atleast_2d(KernelType.TYPE_IDENTIFIER))
        self.broadcast_parameters(is_covariant, is_isotropic)

    @property
    def fold(self) -> Fold:
        return self._fold

    @property
    def test_csv(self) -> Path:
        return self._folder / 'test.csv'

    @property
    def test_summary_csv(self) -> Path:
        return self._folder / 'test_summary.csv'

    @property
    def kernel(self) -> Kernel:
        return self._kernel

    @property
    def likelihood(self) -> Likelihood:
        return self._likelihood

    @property
    def L(self) -> int:
        return self._L

    @property
    def M(self) -> int:
        return self._M

    @property
    def N(self) -> int:
        return self._N

    @property
    @abstractmethod
    def implementation(self) -> Tuple[Any, ...]:
        """ An L-tuple of single-output models (variant) or a 1-tuple holding one multi-output model (covariant)."""

    @property
    @abstractmethod
    def X(self) -> Any:
        """ The training inputs."""

    @property
    @abstractmethod
    def Y(self) -> Any:
        """ The training outputs."""

    @property
    @abstractmethod
    def K_cho(self):
        """ The Cholesky factor of the noisy gram: (LN,LN) if covariant else (L,N,N)."""

    @property
    @abstractmethod
    def K_inv_Y(self):
        """ ChoSolve(K_cho, Y), shape (L,1,N)."""

    @abstractmethod
    def predict(self, x: NP.Matrix, y_instead_of_f: bool = True) -> Tuple[NP.Matrix, NP.Matrix]:
        """ The distribution of y (or f) at the (o,M) inputs x as (mean (o,L), std (o,L))."""

    def _predict_device(self, x: NP.Matrix, y_instead_of_f: bool = True):
        """ ``predict`` as device tensors (mean (o,L), std (o,L)); implementations that predict on the device override this to skip the round trip."""
        mean, std = self.predict(x, y_instead_of_f)
        return as_device(mean), as_device(std)

    def predict_df(self, x: NP.Matrix, y_instead_of_f: bool = True, is_normalized: bool = True) -> pd.DataFrame:
        """ Predictions as a DataFrame with M+L+L columns (X, Mean, SD)."""
        Y_heading = self._fold.meta['data']['Y_heading']
        mean, std = self.predict(x, y_instead_of_f)
        result = pd.DataFrame(np.concatenate([x, mean], axis=1), columns=self._fold.test_data.df.columns)
        predictive_std = result.loc[:, [Y_heading]].copy()
        predictive_std.iloc[:] = std
        if not is_normalized:
            result = self._fold.normalization.undo_from(result)
            predictive_std = self._fold.normalization.unscale_Y(predictive_std)
        result = result.rename(columns={Y_heading: 'Mean'}, level=0)
        return result.join([predictive_std.rename(columns={Y_heading: 'SD'}, level=0)])

    @abstractmethod
    def predict_gradient(self, x: NP.Matrix, y_instead_of_f: bool = True):
        """ The gradient GP dy/dx."""

    def test(self) -> Frame:
        """ Predict the fold's test data; writes test.csv (the test data followed by Mean, SD, Abs Error, Z Score and Outlier per output, plus
        Outlier for any / all outputs) and test_summary.csv (RMSE, mean SD and outlier fraction per output) - the reference's files
        (gpr/models.py:235-272), assembled in one pass."""
        data = self._fold.test_data.df
        Y_heading = self._fold.meta['data']['Y_heading']
        outputs = list(data[Y_heading].columns)
        L = len(outputs)
        # prediction, errors, z scores, outlier flags and the summary row stay on the device (rc_test_metrics); ONE copy brings the table back
        mean_d, std_d = self._predict_device(self._fold.test_x.values)
        reals_d, flags_d, summary_d = _capi.test_metrics(as_device(data[Y_heading].to_numpy(dtype=float)), mean_d, std_d)
        table = torch.cat([mean_d, std_d, reals_d, flags_d], dim=1).cpu().numpy()
        summary_row = summary_d.cpu().numpy()                                  # RMSE (L) | mean SD (L) | outlier fractions (L + 2)

        def columns(label, names=outputs):
            return pd.MultiIndex.from_tuples([(label, name) for name in names])

        reals = pd.DataFrame(table[:, :4 * L], index=data.index, columns=columns('Mean').append([columns('SD'), columns('Abs Error'), columns('Z Score')]))
        flags = pd.DataFrame(table[:, 4 * L:] != 0.0, index=data.index, columns=columns('Outlier').append(columns('Outlier', ['Any Output', 'All Outputs'])))
        result = Frame(self.test_csv, pd.concat([data, reals, flags], axis=1))
        summary = pd.DataFrame([summary_row], columns=columns('RMSE').append([columns('SD'), flags.columns]))
        Frame(self.test_summary_csv, summary)
        return result

    def broadcast_parameters(self, is_covariant: bool, is_isotropic: bool) -> 'GPR':
        """ Broadcast likelihood and kernel data to the requested shapes; the stored likelihood variance is re-diagonalised (quirk Q1)."""
        target_shape = (self._L, self._L) if is_covariant else (1, self._L)
        self._likelihood.data.frames.variance.broadcast_value(target_shape=target_shape, is_diagonal=True)
        self._kernel.broadcast_parameters(variance_shape=target_shape, M=1 if is_isotropic else self._M)
        self._implementation = None
        self._implementation = self.implementation
        return self


# noinspection PyPep8Naming
class MOGP(GPR):
    """ Implementation of a Gaussian Process on the B200 device path."""

    @classproperty
    def META(cls) -> Dict[str, Any]:
        return {'maxiter': 5000, 'gtol': 1E-16}

    @property
    def implementation(self) -> Tuple[Any, ...]:
        if self._implementation is None:
            variance = self._likelihood._data.frames.variance.np
            if self._likelihood.is_covariant:
                self._implementation = tuple(mf.models.MOGPR(data=(self._X, self._Y), kernel=kernel, mean_function=None, noise_variance=variance)
                                             for kernel in self._kernel.implementation)
            else:
                self._implementation = tuple(gf.models.GPR(data=(self._X, self._Y[:, [l]]), kernel=kernel, mean_function=None,
                                                           noise_variance=max(variance[0, l], self._likelihood.VARIANCE_FLOOR))
                                             for l, kernel in enumerate(self._kernel.implementation))
        return self._implementation

    def calibrate(self, method: str = 'L-BFGS-B', **kwargs) -> Dict[str, Any]:
        """ Optimize the hyper-parameters. ``kernel={...}`` / ``likelihood={...}`` override the trainability META of each."""
        self._fac_cache = self._kiy_cache = None
        meta = (self.read_meta() if self._meta_json.exists() else self.META)
        kernel_options = self._kernel.calibrate(**(meta.pop('kernel', {}) | kwargs.pop('kernel', {})))
        likelihood_options = self._likelihood.calibrate(**(meta.pop('likelihood', {}) | kwargs.pop('likelihood', {})))
        meta.update(kwargs)
        meta.pop('result', None)
        opt = gf.optimizers.Scipy()
        # The reference fits the L single-output GPs of a variant model one after another (gpr/models.py:359-361); they are independent, so here
        # they run side by side and every L-BFGS-B iteration evaluates all of them in one batched launch (romcomma.lockstep).  Same optimiser,
        # same trajectories, same result tuple.
        results = tuple(lockstep.run_together([(lambda gp=gp: opt.minimize(closure=gp.training_loss, variables=gp.trainable_variables, method=method,
                                                                            options=dict(meta))) for gp in self._implementation]))
        meta.update({'result': str(results), 'kernel': kernel_options, 'likelihood': likelihood_options})
        self.write_meta(meta)
        gps = self._implementation
        if self._likelihood.is_covariant:
            self._likelihood.data.replace(variance=gps[0].likelihood.variance.value.numpy(), log_marginal=gps[0].log_marginal_likelihood().numpy())
            self._kernel.data.replace(variance=gps[0].kernel.variance.value.numpy(), lengthscales=gps[0].kernel.lengthscales_neat.numpy())
        else:
            self._likelihood.data.replace(variance=tuple(float(gp.likelihood.variance.numpy()) for gp in gps),
                                          log_marginal=tuple(float(gp.log_marginal_likelihood()) for gp in gps))
            self._kernel.data.replace(variance=tuple(float(gp.kernel.variance.numpy()) for gp in gps),
                                      lengthscales=tuple(np.atleast_1d(gp.kernel.lengthscales.numpy()) for gp in gps))
        for gp in gps:            # the optimiser's workspaces (one or two n_pad^2 matrices per model) are not needed once the fit is recorded
            gp._plans.clear()
        return meta

    # -- device-side views of the hyper-parameters -----------------------------------------------------------------
    def _hyper(self):
        """ (ls (batch*L, M) device, F (batch,L,L) host, E (batch,L,L) host, L per problem, batch)."""
        gps = self._implementation
        if self._likelihood.is_covariant:
            gp = gps[0]
            self._ls_host = np.ascontiguousarray(np.broadcast_to(gp.kernel.lengthscales_neat.numpy(), (self._L, self._M)))
            return _capi.dev(self._ls_host), gp.kernel.variance.value.numpy()[None], gp.likelihood.variance.value.numpy()[None], self._L, 1
        ls = self._ls_host = np.concatenate([gp.kernel._ls_row(self._M) for gp in gps], axis=0)
        F = np.array([float(gp.kernel.variance.numpy()) for gp in gps]).reshape(-1, 1, 1)
        E = np.array([float(gp.likelihood.variance.numpy()) for gp in gps]).reshape(-1, 1, 1)
        return _capi.dev(ls), F, E, 1, self._L

    def _factorize(self):
        """ The Cholesky factorisation of the noisy gram at the CURRENT hyper-parameters, kept until they change.

        The reference factorises afresh in K_cho, again inside K_inv_Y (gpr/models.py:427-444: two Choleskys per Sobol calibrator, six per
        fold of user.run.gsa) and again in every predict; here predict, predict_gradient, K_cho, K_inv_Y and the Sobol calibrators of one
        fitted GP share one factor (same numbers: it is the same deterministic computation, done once).  The factor is read-only for
        its users; ``calibrate`` drops it before the optimiser allocates its own workspace."""
        ls, F, E, L, batch = self._hyper()
        key = (self._ls_host.tobytes(), np.ascontiguousarray(F).tobytes(), np.ascontiguousarray(E).tobytes())   # host copies: no device sync for a cache key
        cached = getattr(self, '_fac_cache', None)
        if cached is not None and cached[0] == key:
            return cached[1], L, batch
        self._fac_cache = None
        Xd = as_device(self._X)
        K = _capi.gram(Xd, None, ls, _capi.dev(F), _capi.dev(E), batch=batch, lower_only=True, pad_to=L * self._N, pad_identity=True)
        fac = _capi.Factorization(K)
        fac.raise_if_failed()
        self._fac_cache = (key, fac)
        return fac, L, batch

    def _predict_device(self, X: NP.Matrix, y_instead_of_f: bool = True):
        """ ``predict`` without the copy to the host: (mean (o,L), std (o,L)) as device tensors."""
        ls, F, E, L, batch = self._hyper()
        mean, var = gf.predict_core(as_device(self._X), as_device(self._Y), ls, F, E, as_device(np.asarray(X, dtype=FLOAT())), L, batch, y_instead_of_f,
                                    fac=self._factorize()[0])                   # (batch, o, L_problem)
        o = mean.shape[1]
        return mean.permute(1, 0, 2).reshape(o, -1).contiguous(), torch.sqrt(var).permute(1, 0, 2).reshape(o, -1).contiguous()

    def predict(self, X: NP.Matrix, y_instead_of_f: bool = True) -> Tuple[NP.Matrix, NP.Matrix]:
        mean, std = self._predict_device(X, y_instead_of_f)
        return np.atleast_2d(mean.cpu().numpy()), np.atleast_2d(std.cpu().numpy())

    def predict_gradient(self, x: NP.Matrix, y_instead_of_f: bool = True):
        """ The gradient GP dy/dx at the (o,M) inputs x (reference gpr/models.py:386-415): mean (o,L,M) and cov (O,o,L,M,m).

        Variant GPs only: the reference's covariant branch calls ``kernel(x)`` on an ``MOStationary`` whose ``__call__`` requires ``X2`` and
        raises a TypeError (:405); the same exception is raised here.  ``y_instead_of_f`` is accepted and ignored, as in the reference.
        All of it runs in the library: the Jacobian of k(X, x) and the mean (rc_predict_gradient_jacobian), the triangular solve with K_cho
        (rc_trsm_fwd), the (oM)^2 N contraction -W^T W on FP64 tensor-core tiles (rc_syrk_tn) and the assembly (rc_predict_gradient_finish).
        """
        if self._likelihood.is_covariant:
            raise TypeError("MOStationary.__call__() missing 1 required positional argument: 'X2'")
        xd = as_device(np.asarray(x, dtype=FLOAT()))
        ls, F, E, _, batch = self._hyper()                                                # ls (L,M) device; F, E (L,1,1) host
        KiY = self.K_inv_Y.as_subclass(torch.Tensor).reshape(self._L, self._N).contiguous()
        mean, var = _capi.predict_gradient(as_device(self._X), xd, ls, _capi.dev(F.reshape(-1)), KiY, self._factorize()[0])
        return DeviceTensor.wrap(mean), DeviceTensor.wrap(var)

    @property
    def X(self) -> DeviceTensor:
        """ The training inputs as an (N,M) design matrix."""
        return self._implementation[0].data[0]

    @property
    def Y(self) -> DeviceTensor:
        """ The training outputs as an (N,L) design matrix."""
        return DeviceTensor.wrap(as_device(self._Y))

    @property
    def K_cho(self) -> DeviceTensor:
        fac, L, batch = self._factorize()
        cho = fac.lower(L * self._N)
        return DeviceTensor.wrap(cho[0] if self._likelihood.is_covariant else cho)

    @property
    def K_inv_Y(self) -> DeviceTensor:
        fac, L, batch = self._factorize()
        cached = getattr(self, '_kiy_cache', None)
        if cached is not None and cached[0] is fac:          # same factor object = same hyper-parameters (see _factorize)
            return DeviceTensor.wrap(cached[1].clone())
        n = L * self._N
        y = torch.zeros((batch, fac.n_pad), dtype=torch.float64, device='cuda')
        y[:, :n] = as_device(self._Y).reshape(self._N, batch, L).permute(1, 2, 0).reshape(batch, n)
        x = fac.trsv(fac.trsv(y), transpose=True)
        x = x[:, :n].reshape(self._L, 1, self._N).contiguous()
        self._kiy_cache = (fac, x)
        return DeviceTensor.wrap(x.clone())

    def check_K_inv_Y(self, x: NP.Matrix) -> NP.Matrix:
        """ FOR TESTING PURPOSES ONLY. kernel(x, X) K_inv_Y - predicted mean: should be 0 to within numerical tolerance."""
        predicted = self.predict(x)[0]
        o = predicted.shape[0]
        KiY = self.K_inv_Y.numpy()
        if self._likelihood.is_covariant:
            kernel = self._implementation[0].kernel(x, self._X).numpy().reshape(self._L, o, self._L, self._N)
            result = np.einsum('loLN, LiN -> ol', kernel, KiY)
        else:
            kernel = np.stack([gp.kernel(x, self._X).numpy() for gp in self._implementation], axis=0)
            result = np.einsum('loN, liN -> ol', kernel, KiY)
        result -= predicted
        return np.sqrt(np.sum(result * result, axis=0) / o)
