""" Test functions (reference romcomma/user/functions.py:30-173).  The reference wraps SALib's Ishigami, Sobol-G and Oakley-2004
evaluators; SALib is not a dependency here, so their published closed forms are restated below with the same call signatures."""
from __future__ import annotations

from romcomma.base.definitions import *


def ishigami_evaluate(X: NP.Matrix, A: float = 7.0, B: float = 0.1) -> NP.Vector:
    """ sin x1 + A sin^2 x2 + B x3^4 sin x1  (SALib.test_functions.Ishigami.evaluate)."""
    return np.sin(X[:, 0]) + A * np.sin(X[:, 1]) ** 2 + B * X[:, 2] ** 4 * np.sin(X[:, 0])


def sobol_g_evaluate(X: NP.Matrix, a=None, delta=None, alpha=None) -> NP.Vector:
    """ prod_i ((1 + alpha_i) |2 (x_i + delta_i - floor(x_i + delta_i)) - 1|^alpha_i + a_i) / (1 + a_i)  (SALib.test_functions.Sobol_G.evaluate)."""
    a = np.array([0, 1, 4.5, 9, 99, 99, 99, 99]) if a is None else np.asarray(a, dtype=float)
    delta = np.zeros_like(a, dtype=float) if delta is None else np.asarray(delta, dtype=float)
    alpha = np.ones_like(a, dtype=float) if alpha is None else np.asarray(alpha, dtype=float)
    shifted = X + delta
    mod_x = shifted - np.floor(shifted)
    return np.prod(((1 + alpha) * np.abs(2 * mod_x - 1) ** alpha + a) / (1 + a), axis=1)


def oakley2004_evaluate(X: NP.Matrix, A=None, M=None) -> NP.Vector:
    """ a1.x + a2.sin x + a3.cos x + x^T M x  (SALib.test_functions.oakley2004.evaluate)."""
    a1, a2, a3 = (np.asarray(a, dtype=float) for a in A)
    M = np.asarray(M, dtype=float)
    return X @ a1 + np.sin(X) @ a2 + np.cos(X) @ a3 + np.einsum('ni,ij,nj->n', X, M, X)


class Scalar:
    """ ``scalar(X, **kwargs)`` evaluates ``call(loc + scale * X[:, :m], **(self.kwargs | kwargs))`` as an (N,1) column."""

    def __init__(self, call: Callable[..., NP.Vector], loc: NP.VectorLike, scale: NP.VectorLike, m: int, **kwargs):
        self._call, self._loc, self._scale, self._m, self._kwargs = call, loc, scale, m, kwargs

    call = property(lambda self: self._call)
    loc = property(lambda self: self._loc)
    scale = property(lambda self: self._scale)
    m = property(lambda self: self._m)
    kwargs = property(lambda self: self._kwargs)

    def __call__(self, X: NP.Matrix, **kwargs) -> NP.Matrix:
        return np.reshape(self._call(self._loc + self._scale * X[:, :self._m], **(self._kwargs | kwargs)), (X.shape[0], 1))


class Vector(dict):
    """ A named dictionary of Scalar functions; calling it concatenates their columns into an (N,L) matrix."""

    def __init__(self, name: str, **kwargs: Scalar):
        super().__init__(**kwargs)
        self._name = name

    @classmethod
    def concat(cls, name: str, vectors: Sequence['Vector']) -> 'Vector':
        result = cls(name)
        for vector in vectors:
            result.update({f'{vector.name}.{key}': scalar for key, scalar in vector.items()})
        return result

    @property
    def name(self) -> str:
        return self._name

    @property
    def meta(self) -> Dict:
        return {'name': self.name, 'call': {l: function for l, function in enumerate(self.keys())}}

    def subVector(self, name: str, scalars: Sequence[str]) -> 'Vector':
        return Vector(name, **{scalar: self[scalar] for scalar in scalars})

    def __call__(self, X: NP.Matrix, **kwargs) -> NP.Matrix:
        return np.concatenate([scalar(X, **kwargs) for scalar in self.values()], axis=1)


def linspace(start: float, stop: float, shape: Sequence[int]) -> NP.Matrix:
    """ ``np.linspace`` spread through ``shape``."""
    return np.reshape(np.linspace(start, stop, int(np.prod(shape)), endpoint=True), shape)


_ISHIGAMI = {'call': ishigami_evaluate, 'loc': -np.pi, 'scale': 2 * np.pi}
_SOBOL_G = {'call': sobol_g_evaluate, 'loc': 0, 'scale': 1}
_OAKLEY2004 = {'call': oakley2004_evaluate, 'loc': -1, 'scale': 2}

ISHIGAMI = Vector(name='ishigami',
                  standard=Scalar(**_ISHIGAMI, m=3, A=7.0, B=0.1),
                  balanced=Scalar(**_ISHIGAMI, m=3, A=20.0, B=1.0),
                  sin=Scalar(**_ISHIGAMI, m=3, A=0.0, B=0.0))

SOBOL_G = Vector(name='sobol_g',
                 weak5_2=Scalar(**_SOBOL_G, m=5, a=np.array([3, 6, 9, 18, 27]), alpha=np.ones((5,)) * 2.0),
                 strong5_2=Scalar(**_SOBOL_G, m=5, a=np.array([1 / 2, 1, 2, 4, 8]), alpha=np.ones((5,)) * 2.0),
                 strong5_4=Scalar(**_SOBOL_G, m=5, a=np.array([1 / 2, 1, 2, 4, 8]), alpha=np.ones((5,)) * 4.0))


def _oakley(m: int) -> Vector:
    zeros = [np.zeros([m])] * 2
    return Vector(name='oakley2004',
                  lin7=Scalar(**_OAKLEY2004, m=m, A=[linspace(float(m), m / 2, [m])] + zeros, M=np.zeros([m, m])),
                  quad7=Scalar(**_OAKLEY2004, m=m, A=[linspace(float(m), m / 2, [m])] + zeros, M=linspace(float(m), 1.0, [m, m])),
                  balanced_quad7=Scalar(**_OAKLEY2004, m=m, A=[-linspace(float(m), m / 2, [m])] + zeros, M=linspace(1.0, float(m), [m, m])))


OAKLEY2004_5 = _oakley(5)
OAKLEY2004 = _oakley(7)
ALL = Vector.concat(name='all', vectors=(ISHIGAMI, SOBOL_G, OAKLEY2004))
