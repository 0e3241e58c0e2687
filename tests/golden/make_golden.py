"""Generates tests/golden/*.npz.  Run from the repo root:  python tests/golden/make_golden.py

The reference itself cannot be imported in this project's container (TensorFlow / GPflow / SALib absent, no network), so these
vectors do NOT come from running rom-comma.  They come from ``oracle/literal.py`` - the line-by-line torch-CPU float64
transliteration of the reference's TensorFlow code (same broadcast tensors, same einsum strings, autograd for the gradients) -
and are cross-checked here against the independent closed-form oracle (``oracle/gp.py``, ``oracle/sobol.py``) before being
written.  Parity therefore remains "unpinned" with respect to an actual run of the reference (see oracle/__init__.py).
The first case uses the only fixture inputs the reference holds: the data set of romcomma/gpf/tests.py:41-54.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import gp, literal, sobol  # noqa: E402

OUT = Path(__file__).resolve().parent


def case(name, X, Y, ls, F, E, Xs, slices, with_sobol=True):
    L, M = Y.shape[1], X.shape[1]
    uF, lowF = gp.variance_pack(F)
    uE, lowE = gp.variance_pack(E)
    uls = gp.softplus_inverse(ls)
    lml, grads = literal.lml_grad_unconstrained_mo(X, Y, uls, uF, lowF, uE, lowE)
    grads = [np.zeros(0) if gr is None else gr for gr in grads]          # L = 1: no strict lower triangle
    g = gp.lml_grad_mo(X, Y, ls, F, E)
    assert abs(lml - g['lml']) <= 1e-9 * abs(lml), (lml, g['lml'])
    dFd, dFl = gp.chain_variance(g['dF'], uF, lowF)
    dEd, dEl = gp.chain_variance(g['dE'], uE, lowE)
    for a, b in ((g['dls'] * gp.sigmoid(uls), grads[0]), (dFd, grads[1]), (dFl, grads[2]), (dEd, grads[3]), (dEl, grads[4])):
        assert np.allclose(a, b, rtol=1e-7, atol=1e-9), (name, a, b)
    mean, var = literal.predict_mo(X, Y, ls, F, E, Xs)
    mean2, var2 = gp.predict_mo(X, Y, ls, F, E, Xs)
    assert np.allclose(mean, mean2, rtol=1e-8, atol=1e-10) and np.allclose(var, var2, rtol=1e-8, atol=1e-10)
    Ed = np.diag(np.diag(E))                                    # quirk Q1: a GP re-read for GSA has diagonal noise
    KiY = gp.k_inv_y_mo(X, Y, ls, F, Ed)
    out = dict(X=X, Y=Y, ls=ls, F=F, E=E, Xs=Xs, slices=np.array(slices), lml=lml, g_uls=grads[0], g_uFd=grads[1], g_Flow=grads[2],
               g_uEd=grads[3], g_Elow=grads[4], dF=g['dF'], dE=g['dE'], dls=g['dls'], mean=mean, var=var, KiY=KiY)
    for tag, diag in (('diag', True), ('full', False)) if with_sobol else ():
        lit = literal.closed_sobol_literal(X, ls, F, KiY, diag)
        cal = sobol.ClosedSobol(X, ls, F, KiY, diag)
        V = np.stack([lit['V'](*s) for s in slices])
        V2 = np.stack([cal._V(*s) for s in slices])
        assert np.allclose(V, V2, rtol=1e-8, atol=1e-10), (name, tag)
        out[f'V_{tag}'] = V
        out[f'g0KY_{tag}'] = lit['g0KY']
    np.savez(OUT / f'{name}.npz', **out)
    print(name, 'lml', lml)


def main():
    # 1. the data set of the reference's manual smoke script, romcomma/gpf/tests.py:41-54
    data = np.linspace(start=1, stop=50, num=50).reshape(5, 10).transpose()
    X, Y = data[:, :3], data[:, 3:]
    ls = np.stack([0.01 * np.ones(3), 0.03 * np.ones(3)])
    case('gpf_tests', X, Y, ls, 0.5 * np.eye(2), 0.0001 * np.eye(2), X[:4] + 0.005, [(0, 3)], with_sobol=False)   # un-normalised X in 1..30: meaningless (and overflowing) for GSA
    # 2./3. random problems, diagonal and full F, non-diagonal E
    for name, (N, M, L, seed) in {'rand_a': (17, 5, 3, 0), 'rand_b': (33, 2, 2, 1), 'rand_c': (70, 4, 1, 2)}.items():
        rng = np.random.default_rng(seed)
        X = rng.normal(size=(N, M))
        Y = rng.normal(size=(N, L))
        ls = rng.uniform(0.5, 3.0, (L, M))
        A = rng.normal(size=(L, L))
        F = A @ A.T / L + np.eye(L)
        B = rng.normal(size=(L, L))
        E = 0.01 * (B @ B.T / L + np.eye(L))
        slices = [(0, M), (0, 1), (M - 1, M), (0, max(1, M // 2)), (1, M), (M, M)]
        case(name, X, Y, ls, F, E, rng.normal(size=(6, M)), slices)


if __name__ == '__main__':
    main()
