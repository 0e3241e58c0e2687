""" Storage-backed kernels for gpr (reference romcomma/gpr/kernels.py:30-180): parameters live in ``kernel/{variance,lengthscales}.csv``;
``implementation`` is a tuple of L independent ``gf.kernels.RBF`` (variance shape (1,L): the *variant* model) or one multi-output
``mf.kernels.RBF`` (variance shape (L,L): the *covariant* model)."""
from __future__ import annotations

from romcomma.base.definitions import *
from romcomma.base.classes import Data, Model


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
        return {'variance': True, 'covariance': False, 'lengthscales': {'variant': True, 'covariant': False}}

    @classproperty
    def TYPE_IDENTIFIER(cls) -> str:
        """ '<module tail>.<class name>', e.g. 'kernels.RBF': what GPR stores in kernel.csv."""
        return cls.__module__.split('.')[-1] + '.' + cls.__name__

    @classmethod
    def TypeFromIdentifier(cls, TypeIdentifier: str) -> Type['Kernel']:
        for KernelType in cls.__subclasses__():
            if KernelType.TYPE_IDENTIFIER == TypeIdentifier:
                return KernelType
        raise TypeError('Kernel.TypeIdentifier() of unrecognizable type.')

    @classmethod
    def TypeFromParameters(cls, parameters: Data) -> Type['Kernel']:
        for kernel_type in cls.__subclasses__():
            if isinstance(parameters, kernel_type.Data):
                return kernel_type
        raise TypeError('Kernel Data array of unrecognizable type.')

    def __init__(self, folder: Path | str, read_data: bool = False, **kwargs):
        super().__init__(folder, read_data, **kwargs)
        variance_shape = self._data.frames.variance.df.shape
        self._L, self._M = variance_shape[1], self._data.frames.lengthscales.df.shape[1]
        self.broadcast_parameters(variance_shape, self._M)

    def calibrate(self, **kwargs: Any) -> Dict[str, Any]:
        """ Merely sets which parameters are trainable."""
        meta = self.META | kwargs
        if self.is_covariant:
            gf.set_trainable(self._implementation[0].variance._cholesky_diagonal, meta['variance'])
            gf.set_trainable(self._implementation[0].variance._cholesky_lower_triangle, meta['covariance'])
            gf.set_trainable(self._implementation[0].lengthscales, meta['lengthscales']['covariant'])
        else:
            for implementation in self._implementation:
                gf.set_trainable(implementation.variance, meta['variance'])
                gf.set_trainable(implementation.lengthscales, meta['lengthscales']['variant'])
        return meta

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
        """ Broadcast to (1,L) / (L,L) variance and M lengthscales per output; a diagonal variance stays diagonal when made square."""
        if variance_shape != self._data.frames.variance.df.shape:
            self._data.frames.variance.broadcast_value(target_shape=variance_shape, is_diagonal=True)
            self._L = variance_shape[1]
        if (self._L, M) != self._data.frames.lengthscales.df.shape:
            self._data.frames.lengthscales.broadcast_value(target_shape=(self._L, M), is_diagonal=False)
            self._M = M
        self._implementation = None
        self._implementation = self.implementation
        return self

    @property
    @abstractmethod
    def implementation(self) -> Tuple[Any, ...]:
        """ An L-tuple of single-output kernels, or a 1-tuple holding one multi-output kernel."""


class RBF(Kernel):

    @property
    def implementation(self) -> Tuple[Any, ...]:
        if self._implementation is None:
            variance = self._data.frames.variance.np
            lengthscales = self._data.frames.lengthscales.np
            if variance.shape[0] == 1:
                self._implementation = tuple(gf.kernels.RBF(variance=max(variance[0, l], 1.0005E-6), lengthscales=lengthscales[l])
                                             for l in range(variance.shape[1]))
            else:
                self._implementation = (mf.kernels.RBF(variance=variance, lengthscales=lengthscales), )
        return self._implementation
