""" Kernels whose hyper-parameters live on disk (``kernel/variance.csv``, ``kernel/lengthscales.csv``) - the counterpart of the reference's
romcomma/gpr/kernels.py:30-180, behind the same names, with own structure.

The shape of the stored variance decides the model family: one row, ``(1, L)``, means L independent single-output kernels (the *variant* model,
``gf.kernels.RBF`` per output); a square ``(L, L)`` matrix means one multi-output kernel (the *covariant* model, ``mf.kernels.RBF``).  The
lengthscales are ``(L, M)`` (anisotropic) or ``(L, 1)`` (isotropic).  ``implementation`` is the tuple of device-path kernel objects built from
those numbers; ``calibrate`` only marks which of their parameters the optimiser may move.
"""
from __future__ import annotations

from romcomma.base.definitions import *
from romcomma.base.classes import Data, Model

#: Smallest variance a variant kernel is built with (gpflow's positive() transform needs a strictly positive start).
_VARIANT_VARIANCE_FLOOR = 1.0005E-6


class Kernel(Model):
    """ Abstract interface to a Kernel: the contract with the MOGP interface."""

    class Data(Data):
        class NamedTuple(NamedTuple):
            """ variance: (L,L), or (1,L) standing for a diagonal (L,L) matrix, or (1,1) for one kernel shared by all outputs.
                lengthscales: (L,M) anisotropic or (L,1) isotropic."""
            variance: Any = np.atleast_2d(2.0)
            lengthscales: Any = np.atleast_2d(5.0)

    @classproperty
    def META(cls) -> Dict[str, Any]:
        """ What ``calibrate`` lets the optimiser move by default: the variances, not the covariances; the lengthscales of a variant model only."""
        return {'variance': True, 'covariance': False, 'lengthscales': {'variant': True, 'covariant': False}}

    @classproperty
    def TYPE_IDENTIFIER(cls) -> str:
        """ '<module tail>.<class name>', e.g. 'kernels.RBF': what GPR stores in kernel.csv."""
        return f'{cls.__module__.rsplit(".", 1)[-1]}.{cls.__name__}'

    @classmethod
    def _concrete_types(cls) -> Dict[str, Type['Kernel']]:
        return {kernel_type.TYPE_IDENTIFIER: kernel_type for kernel_type in cls.__subclasses__()}

    @classmethod
    def TypeFromIdentifier(cls, TypeIdentifier: str) -> Type['Kernel']:
        try:
            return cls._concrete_types()[TypeIdentifier]
        except KeyError:
            raise TypeError('Kernel.TypeIdentifier() of unrecognizable type.') from None

    @classmethod
    def TypeFromParameters(cls, parameters: Data) -> Type['Kernel']:
        matches = [kernel_type for kernel_type in cls._concrete_types().values() if isinstance(parameters, kernel_type.Data)]
        if not matches:
            raise TypeError('Kernel Data array of unrecognizable type.')
        return matches[0]

    def __init__(self, folder: Path | str, read_data: bool = False, **kwargs):
        super().__init__(folder, read_data, **kwargs)
        stored = self._data.frames
        self._L, self._M = stored.variance.df.shape[1], stored.lengthscales.df.shape[1]
        self.broadcast_parameters(stored.variance.df.shape, self._M)

    # -- shapes -----------------------------------------------------------------------------------------------------------------------
    @property
    def L(self) -> int:
        return self._L

    @property
    def M(self) -> int:
        return self._M

    @property
    def is_covariant(self) -> bool:
        return self._data.frames.variance.df.shape[0] > 1

    def broadcast_parameters(self, variance_shape: Tuple[int, int], M) -> 'Kernel':
        """ Bring the stored parameters to a (1,L) / (L,L) variance and M lengthscales per output (a diagonal variance stays diagonal when it is
        made square), then rebuild ``implementation`` from them."""
        frames = self._data.frames
        if tuple(frames.variance.df.shape) != tuple(variance_shape):
            frames.variance.broadcast_value(target_shape=variance_shape, is_diagonal=True)
            self._L = variance_shape[1]
        if tuple(frames.lengthscales.df.shape) != (self._L, M):
            frames.lengthscales.broadcast_value(target_shape=(self._L, M), is_diagonal=False)
            self._M = M
        self._implementation = None          # stale: built from the old shapes
        self._implementation = self.implementation
        return self

    # -- trainability -----------------------------------------------------------------------------------------------------------------
    def calibrate(self, **kwargs: Any) -> Dict[str, Any]:
        """ Merely sets which parameters are trainable; returns the options that were applied (META overridden by kwargs)."""
        options = self.META | kwargs
        if self.is_covariant:
            kernel = self._implementation[0]
            switches = ((kernel.variance._cholesky_diagonal, options['variance']), (kernel.variance._cholesky_lower_triangle, options['covariance']),
                        (kernel.lengthscales, options['lengthscales']['covariant']))
        else:
            switches = tuple(pair for kernel in self._implementation
                             for pair in ((kernel.variance, options['variance']), (kernel.lengthscales, options['lengthscales']['variant'])))
        for parameter, is_trainable in switches:
            gf.set_trainable(parameter, is_trainable)
        return options

    @property
    @abstractmethod
    def implementation(self) -> Tuple[Any, ...]:
        """ An L-tuple of single-output kernels, or a 1-tuple holding one multi-output kernel."""


class RBF(Kernel):
    """ The ARD squared-exponential kernel."""

    @property
    def implementation(self) -> Tuple[Any, ...]:
        if self._implementation is None:
            variance, lengthscales = self._data.frames.variance.np, self._data.frames.lengthscales.np
            if self.is_covariant:
                self._implementation = (mf.kernels.RBF(variance=variance, lengthscales=lengthscales),)
            else:
                self._implementation = tuple(gf.kernels.RBF(variance=max(float(v), _VARIANT_VARIANCE_FLOOR), lengthscales=ls)
                                             for v, ls in zip(variance[0], lengthscales))
        return self._implementation
