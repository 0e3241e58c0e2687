"""One scripted walk through rom-comma's public API for the dense-GP hot path.  TEST INFRASTRUCTURE ONLY.

The same function is executed twice:
  * by ``tests/golden/make_golden_from_reference.py`` against the REFERENCE'S OWN, UNMODIFIED source files
    (/root/reference/romcomma/{base,data,gpf,gpr,gsa}) running over the torch-backed tensorflow/gpflow stand-in of
    ``tests/golden/_refshim`` -> ``tests/golden/ref_*.npz``;
  * by ``tests/test_gpu_reference_golden.py`` against this repo's drop-in package (``rom-comma_b200/romcomma``, CUDA path through
    the C ABI) on the B200 box, whose results must equal the stored vectors (fold assignment bit-exact, floats rtol 1e-8/atol 1e-10).

``api`` is a namespace holding the modules/classes of whichever implementation is under test plus two adapters
(``loss_and_grads``, ``to_np``); everything else is the public surface SURVEY.md section 8(b) lists.
"""
from __future__ import annotations

import random
from pathlib import Path
from types import SimpleNamespace
from typing import Dict

import numpy as np
import pandas as pd

CASES = {
    # name: sizes, folds, and whether the covariant and/or the variant (independent outputs) model is exercised
    'ref_cov_a': dict(N=36, M=3, L=2, K=3, seed=11, shuffle=False, covariant=True, full_F=True),
    'ref_cov_b': dict(N=30, M=4, L=3, K=2, seed=12, shuffle=True, covariant=True, full_F=False),
    'ref_var_a': dict(N=40, M=3, L=2, K=4, seed=13, shuffle=False, covariant=False, full_F=False),
    'ref_var_b': dict(N=48, M=3, L=1, K=-5, seed=14, shuffle=True, covariant=False, full_F=False),    # cfg1-like: Ishigami, L=1 (quirk Q4)
}


def raw_data(case) -> pd.DataFrame:
    """Deterministic raw data in the two-level-header layout of ``data.csv`` (reference data/storage.py:44-46)."""
    N, M, L = case['N'], case['M'], case['L']
    rng = np.random.default_rng(case['seed'])
    U = rng.uniform(-np.pi, np.pi, size=(N, M))
    cols = [np.sin(U[:, 0]) + 7.0 * np.sin(U[:, 1]) ** 2 + 0.1 * U[:, 2] ** 4 * np.sin(U[:, 0]),     # Ishigami (user/functions.py:126)
            np.cos(U[:, 0]) * U[:, 1] + 0.5 * U[:, -1] ** 2,
            np.exp(-0.3 * U[:, 0] ** 2) + 0.2 * U[:, 1] * U[:, 2]]
    Y = np.stack(cols[:L], axis=1)
    Y = Y + 0.04 * Y.std(axis=0) * rng.normal(size=Y.shape)
    columns = pd.MultiIndex.from_tuples([('X', f'x.{m}') for m in range(M)] + [('Y', f'y.{l}') for l in range(L)])
    return pd.DataFrame(np.concatenate([U, Y], axis=1), columns=columns)


def hyper_parameters(case):
    L, M = case['L'], case['M']
    rng = np.random.default_rng(case['seed'] + 100)
    ls = rng.uniform(0.6, 2.5, (L, M))
    if case['covariant']:
        F = np.diag(rng.uniform(0.6, 1.8, L))
        if case['full_F']:
            A = rng.normal(size=(L, L))
            F = A @ A.T / L + np.eye(L)
        B = rng.normal(size=(L, L))
        E = 0.02 * (np.eye(L) + B @ B.T / L)       # a full noise covariance: the reference re-diagonalises it (quirk Q1)
    else:
        F = rng.uniform(0.6, 1.8, (1, L))
        E = rng.uniform(0.01, 0.05, (1, L))
    return ls, F, E


def _walk_named(model, api) -> Dict[str, object]:
    """Trainable variables keyed by attribute path, so the two implementations can be compared parameter by parameter."""
    named = {}
    if hasattr(model.kernel.variance, '_cholesky_diagonal'):
        cand = {'kernel.lengthscales': model.kernel.lengthscales,
                'kernel.variance.cholesky_diagonal': model.kernel.variance._cholesky_diagonal,
                'kernel.variance.cholesky_lower_triangle': model.kernel.variance._cholesky_lower_triangle,
                'likelihood.variance.cholesky_diagonal': model.likelihood.variance._cholesky_diagonal,
                'likelihood.variance.cholesky_lower_triangle': model.likelihood.variance._cholesky_lower_triangle}
    else:
        cand = {'kernel.lengthscales': model.kernel.lengthscales, 'kernel.variance': model.kernel.variance,
                'likelihood.variance': model.likelihood.variance}
    for key, p in cand.items():
        if p.trainable:
            named[key] = p
    return named


def _grads(model, api, out, tag):
    named = _walk_named(model, api)
    order = api.variable_order(model)                    # the order the optimizer packs them in (tf.Module attribute order)
    assert sorted(order) == sorted(named), (order, sorted(named))
    out[f'{tag}.order'] = np.array(order)
    loss, grads = api.loss_and_grads(model, [named[k] for k in order])
    out[f'{tag}.loss'] = np.float64(loss)
    for key, g in zip(order, grads):
        out[f'{tag}.d.{key}'] = np.asarray(g, dtype=np.float64).reshape(-1)


def run(api: SimpleNamespace, name: str, root: Path) -> Dict[str, np.ndarray]:
    case = CASES[name]
    to_np = api.to_np
    out: Dict[str, np.ndarray] = {}
    N, M, L, K = case['N'], case['M'], case['L'], case['K']
    df = raw_data(case)
    out['raw'] = df.values

    # ---- a14: repository, integer fold assignment, normalisation ------------------------------------------------------
    repo = api.Repository.from_df(Path(root) / name, df)
    random.seed(case['seed'])
    repo.into_K_folds(K, shuffle_before_folding=case['shuffle'])
    out['folds'] = np.array(list(repo.folds))
    for k in repo.folds:
        fold = api.Fold(repo, k)
        out[f'fold.{k}.train_index'] = np.asarray(fold.data.df.index.values, dtype=np.int64)
        out[f'fold.{k}.test_index'] = np.asarray(fold.test_data.df.index.values, dtype=np.int64)
    k_fit = abs(K) if K > 0 else 0                         # the improper fold (all data) when there is one
    fold = api.Fold(repo, k_fit)
    X, Y = fold.X.values, fold.Y.values
    out['X'], out['Y'] = np.array(X, dtype=np.float64), np.array(Y, dtype=np.float64)
    xs = np.array(fold.test_x.values[:7], dtype=np.float64) * 0.9 + 0.05
    out['xs'] = xs

    # ---- construct the GP with explicit hyper-parameters (defaults -> replace -> re-read, as user code does) ---------------
    ls, F, E = hyper_parameters(case)
    cov = case['covariant']
    gp_name = 'gp.c.a' if cov else 'gp.v.a'
    gp = api.MOGP(gp_name, fold, is_read=False, is_covariant=cov, is_isotropic=False)
    gp.kernel.data.replace(variance=F, lengthscales=ls)
    gp.likelihood.data.replace(variance=E)
    gp = api.MOGP(gp_name, fold, is_read=True, is_covariant=cov, is_isotropic=False)
    out['ls'] = np.array(gp.kernel.data.frames.lengthscales.np, dtype=np.float64)
    out['F'] = np.array(gp.kernel.data.frames.variance.np, dtype=np.float64)
    out['E'] = np.array(gp.likelihood.data.frames.variance.np, dtype=np.float64)     # after the constructor's re-diagonalisation
    out['E_given'] = E

    # ---- a5/a6/a7: LML and its gradient w.r.t. the unconstrained variables ---------------------------------------------------
    models = gp.implementation
    out['lml'] = np.array([float(to_np(m.log_marginal_likelihood())) for m in models])
    gp.kernel.calibrate()                                  # default trainables (gpr/kernels.py:54-57, gpr/models.py:57-60)
    gp.likelihood.calibrate()
    for i, m in enumerate(models):
        _grads(m, api, out, f'grad.default.{i}')
    gp.kernel.calibrate(variance=True, covariance=True, lengthscales={'variant': True, 'covariant': True})
    gp.likelihood.calibrate(variance=True, covariance=True)
    for i, m in enumerate(models):
        _grads(m, api, out, f'grad.all.{i}')
    if cov:
        out['KXX'] = to_np(models[0].KXX)                  # lengthscales trainable -> the recompute branch of MOGPR.KXX (a4)
    gp.kernel.calibrate()
    gp.likelihood.calibrate()

    # ---- a8: prediction (mean, STANDARD DEVIATION: quirk Q6) ------------------------------------------------------------------
    for flag, tag in ((True, 'y'), (False, 'f')):
        mean, std = gp.predict(xs, y_instead_of_f=flag)
        out[f'predict.{tag}.mean'], out[f'predict.{tag}.std'] = np.asarray(mean, dtype=np.float64), np.asarray(std, dtype=np.float64)

    # ---- 8(f)2: the gradient GP dy/dx at a few points (mean (o,L,M) and the reference's covariance tensor, gpr/models.py:386-415) -------
    #       Variant GPs only: for a covariant GP the reference itself fails (``kernel(x)`` at :405 calls MOStationary.__call__ without X2).
    if cov:
        try:
            gp.predict_gradient(xs[:3])
            out['predict_gradient.covariant_raises'] = np.array(False)
        except TypeError:
            out['predict_gradient.covariant_raises'] = np.array(True)
    else:
        mean, var = gp.predict_gradient(xs[:3])
        out['predict_gradient.mean'], out['predict_gradient.var'] = to_np(mean), to_np(var)

    # ---- a9: Cholesky factor and K^-1 Y as GSA consumes them -----------------------------------------------------------------
    out['K_cho'] = np.tril(to_np(gp.K_cho))
    out['K_inv_Y'] = to_np(gp.K_inv_Y)
    out['check_K_inv_Y'] = to_np(gp.check_K_inv_Y(xs))

    # ---- a10-a12: the ClosedSobol calibrator, diagonal and full F -----------------------------------------------------------------
    slices = [(0, M), (0, 1), (M - 1, M), (1, M), (1, 2), (M, M)]
    out['slices'] = np.array(slices)
    for diag in ((True, False) if cov else (True,)):
        tag = 'sobol.diag' if diag else 'sobol.full'
        cal = api.ClosedSobol(gp, is_F_diagonal=diag)
        for attr in ('g0', 'g0KY', 'G', 'Phi', 'S'):
            out[f'{tag}.{attr}'] = to_np(getattr(cal, attr))
        for j in range(3):
            out[f'{tag}.V{j}'] = to_np(cal.V[j])
            out[f'{tag}.Lambda2.{j}'] = to_np(cal.Lambda2[1][j])
        for s in slices:
            r = cal.marginalize(api.slice_arg(s))
            out[f'{tag}.marginalize.{s[0]}.{s[1]}.V'] = to_np(r['V'])
            out[f'{tag}.marginalize.{s[0]}.{s[1]}.S'] = to_np(r['S'])
    out['sobol.default_is_F_diagonal'] = np.array(bool(api.ClosedSobol(gp).is_F_diagonal))

    # ---- a13: the three kinds through gsa.models.Sobol, results as written to V.csv / S.csv before the 6-decimal formatting --------
    for kind in api.GSA.ALL_KINDS:
        for key, value in api.sobol_results(gp, kind, is_error_calculated=False).items():
            out[f'gsa.{kind.name.lower()}.{key}'] = value
    if getattr(api, 'with_error', True):
        for kind in api.GSA.ALL_KINDS:
            for key, value in api.sobol_results(gp, kind, is_error_calculated=True).items():
                out[f'gsa_err.{kind.name.lower()}.{key}'] = value
        # is_T_partial=False - what installation_test.py, csv_script.py and benchmark_script.py run (the MIXED rank equation, W[Mm], Q and the
        # full-model T; gsa/calibrators.py:169-170,342-346,358-372,388-402; T post-processed by gsa/models.py:211-213)
        for kind in api.GSA.ALL_KINDS:
            for key, value in api.sobol_results(gp, kind, is_error_calculated=True, is_T_partial=False).items():
                out[f'gsa_err_full.{kind.name.lower()}.{key}'] = value
        cal = api.ClosedSobolWithError(gp, is_T_partial=False)
        out['sobol_err_full.W_DIAGONAL'], out['sobol_err_full.W_MIXED'] = to_np(cal.W.DIAGONAL), to_np(cal.W.MIXED)
        out['sobol_err_full.Q'], out['sobol_err_full.T'] = to_np(cal.Q), to_np(cal.T)
        r = cal.marginalize(api.slice_arg((1, M)))
        out['sobol_err_full.marginalize.1.M.W'], out['sobol_err_full.marginalize.1.M.T'] = to_np(r['W']), to_np(r['T'])

    # ---- a6 through the optimizer: a short L-BFGS-B run from the given start (maxiter small; compared loosely) ------------------------
    meta = gp.calibrate(method='L-BFGS-B', maxiter=6)
    out['fit.ls'] = np.array(gp.kernel.data.frames.lengthscales.np, dtype=np.float64)
    out['fit.F'] = np.array(gp.kernel.data.frames.variance.np, dtype=np.float64)
    out['fit.E'] = np.array(gp.likelihood.data.frames.variance.np, dtype=np.float64)
    out['fit.lml'] = np.array(gp.likelihood.data.frames.log_marginal.np, dtype=np.float64)
    return out
