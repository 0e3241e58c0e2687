""" Deterministic synthetic workloads of the five benchmark configurations; shared by bench.py, smoke() and the tests.

Inputs: centred Latin hypercube (``scipy.stats.qmc.LatinHypercube(scramble=False, seed)``, as reference user/sample.py:54-67), test
functions of romcomma.user.functions, 4 % Gaussian noise, then the reference's normalisation (data/storage.py:469-485,532-558):
X -> probit of the uniform on [mean - sqrt3 std, mean + sqrt3 std], Y -> standardised.  Hyper-parameters of a timed evaluation:
lengthscales ~ U[0.5, 3], F = diag U[0.5, 2], E = 0.01 I (sigma_n^2 = 1e-2 keeps cond(K) moderate)."""
from __future__ import annotations

from typing import Dict, NamedTuple, Sequence

import numpy as np
import scipy.stats

from romcomma.user import functions

CONFIGS: Dict[str, dict] = {
    'cfg1': dict(N=256, M=3, L=1, seed=1, outputs=('ishigami.standard',)),
    'cfg2': dict(N=2048, M=10, L=1, seed=2, outputs=('sobol_g.weak5_2',), K=10),
    'cfg3': dict(N=4096, M=8, L=4, seed=3, outputs=('ishigami.standard', 'sobol_g.weak5_2', 'sobol_g.strong5_2', 'oakley5.lin7')),
    'cfg4': dict(N=16384, M=20, L=2, seed=4, outputs=('sobol_g.weak5_2', 'oakley5.lin7')),
    'cfg5': dict(N=8192, M=12, L=3, seed=5, outputs=('ishigami.standard', 'sobol_g.strong5_2', 'oakley5.lin7')),
}

_VECTORS = {'ishigami': functions.ISHIGAMI, 'sobol_g': functions.SOBOL_G, 'oakley5': functions.OAKLEY2004_5, 'oakley7': functions.OAKLEY2004}


class Workload(NamedTuple):
    X: np.ndarray        # (N,M) normalised inputs
    Y: np.ndarray        # (N,L) normalised outputs
    lengthscales: np.ndarray   # (L,M)
    F: np.ndarray        # (L,L)
    E: np.ndarray        # (L,L)
    name: str


def normalise(X: np.ndarray, Y: np.ndarray, margin: float = 1.0E-12):
    """ The reference's Normalization computed on all rows."""
    mean, std = X.mean(axis=0), X.std(axis=0, ddof=1)
    lo, rng = mean - np.sqrt(3) * std, 2 * np.sqrt(3) * std
    Xn = scipy.stats.norm.ppf(np.clip((X - lo) / rng, margin, 1 - margin))
    Yn = (Y - Y.mean(axis=0)) / Y.std(axis=0, ddof=1)
    return Xn, Yn


def make(N: int, M: int, L: int, seed: int, outputs: Sequence[str], noise: float = 0.04, name: str = '', full_F: bool = False, **_) -> Workload:
    U = scipy.stats.qmc.LatinHypercube(d=M, scramble=False, seed=seed).random(N)
    cols = []
    for spec in outputs:
        vec, key = spec.split('.')
        cols.append(_VECTORS[vec][key](U))
    Y = np.concatenate(cols, axis=1)
    assert Y.shape == (N, L)
    Y = Y + noise * Y.std(axis=0, keepdims=True) * np.random.default_rng(seed + 1).standard_normal((N, L))
    Xn, Yn = normalise(U, Y)
    rng = np.random.default_rng(seed + 2)
    ls = rng.uniform(0.5, 3.0, (L, M))
    F = np.diag(rng.uniform(0.5, 2.0, L))
    if full_F:
        A = rng.standard_normal((L, L))
        F = A @ A.T / L + np.eye(L)
    E = 0.01 * np.eye(L)
    return Workload(np.ascontiguousarray(Xn), np.ascontiguousarray(Yn), ls, F, E, name)


def config(name: str, N: int | None = None, **overrides) -> Workload:
    cfg = dict(CONFIGS[name]) | overrides
    if N is not None:
        cfg['N'] = N
    return make(name=name, **cfg)
