"""CPU: the oracle against the golden vectors, the literal transliteration, finite differences, quadrature and the reference's
own self-check identity (MOGP.check_K_inv_Y, romcomma/gpr/models.py:446-463)."""
import numpy as np
import pytest

from conftest import assert_close, random_problem
from oracle import gp, literal, sobol


def test_golden_lml_and_gradients(golden):
    for name, g in golden.items():
        X, Y, ls, F, E = g['X'], g['Y'], g['ls'], g['F'], g['E']
        assert_close(gp.lml_mo(X, Y, ls, F, E), g['lml'], what=f'{name} lml')
        r = gp.lml_grad_mo(X, Y, ls, F, E)
        uF, lowF = gp.variance_pack(F)
        uE, lowE = gp.variance_pack(E)
        dFd, dFl = gp.chain_variance(r['dF'], uF, lowF)
        dEd, dEl = gp.chain_variance(r['dE'], uE, lowE)
        # gradients can be tiny relative to the cancelling O(n) terms they are made of: atol scaled by the gradient norm
        for ours, key in ((r['dls'] * gp.sigmoid(gp.softplus_inverse(ls)), 'g_uls'), (dFd, 'g_uFd'), (dFl, 'g_Flow'), (dEd, 'g_uEd'), (dEl, 'g_Elow')):
            scale = max(1.0, float(np.abs(g[key]).max())) if g[key].size else 1.0
            assert_close(ours, g[key].reshape(np.shape(ours)), rtol=1e-7, atol=1e-9 * scale, what=f'{name} {key}')
        assert_close(gp.lml_grad_mo_lapack(X, Y, ls, F, E)['dF'], r['dF'], rtol=1e-7, atol=1e-8, what=f'{name} lapack dF')


def test_golden_predict(golden):
    for name, g in golden.items():
        mean, var = gp.predict_mo(g['X'], g['Y'], g['ls'], g['F'], g['E'], g['Xs'])
        assert_close(mean, g['mean'], what=f'{name} mean')
        assert_close(var, g['var'], what=f'{name} var')


def test_golden_sobol(golden):
    for name, g in golden.items():
        if 'V_diag' not in g:
            continue
        for tag, diag in (('diag', True), ('full', False)):
            cal = sobol.ClosedSobol(g['X'], g['ls'], g['F'], g['KiY'], diag)
            assert_close(cal.g0KY, g[f'g0KY_{tag}'], what=f'{name} g0KY {tag}')
            for s, V in zip(g['slices'], g[f'V_{tag}']):
                assert_close(cal._V(int(s[0]), int(s[1])), V, what=f'{name} V{tuple(s)} {tag}')


def test_known_answer_gpf_tests_dataset(golden):
    """Fixture inputs of romcomma/gpf/tests.py:41-54: lengthscales 0.01/0.03 on integer-spaced X make K exactly (F + E) (x) I
    in float64 (exp(-5000) underflows to 0), so the LML has a closed form."""
    g = golden['gpf_tests']
    Y = g['Y']
    s2 = 0.5 + 1e-4
    expected = -0.5 * np.sum(Y * Y) / s2 - 0.5 * Y.size * np.log(2 * np.pi * s2)
    assert_close(gp.lml_mo(g['X'], Y, g['ls'], g['F'], g['E']), expected, what='analytic LML')
    assert_close(g['lml'], expected, what='golden LML')


@pytest.mark.parametrize('full_F', [False, True])
def test_gradients_vs_central_differences(full_F):
    X, Y, ls, F, E = random_problem(24, 3, 2, seed=5, full_F=full_F)
    r = gp.lml_grad_mo(X, Y, ls, F, E)
    h = 1e-6

    def fd(fun, A, symmetric):
        G = np.zeros_like(A)
        for idx in np.ndindex(*A.shape):
            P, Q = A.copy(), A.copy()
            P[idx] += h
            Q[idx] -= h
            if symmetric and idx[0] != idx[1]:      # keep the matrix symmetric: the derivative is then dA[i,j] + dA[j,i]
                P[idx[::-1]] += h
                Q[idx[::-1]] -= h
            G[idx] = (fun(P) - fun(Q)) / (2 * h)
        return G

    def sym(G):
        return G + G.T - np.diag(np.diag(G))
    assert_close(sym(r['dF']), fd(lambda A: gp.lml_mo(X, Y, ls, A, E), F, True), rtol=1e-5, atol=1e-6, what='dF')
    assert_close(sym(r['dE']), fd(lambda A: gp.lml_mo(X, Y, ls, F, A), E, True), rtol=1e-5, atol=1e-5, what='dE')
    assert_close(r['dls'], fd(lambda A: gp.lml_mo(X, Y, A, F, E), ls, False), rtol=1e-5, atol=1e-6, what='dls')


def test_variant_path_matches_block_of_covariant():
    """L independent gpflow GPRs == the covariant model with diagonal F and E (the two code paths of gpr/models.py:332-343)."""
    X, Y, ls, F, E = random_problem(40, 3, 2, seed=7, full_E=False)
    total = sum(gp.lml_rbf(X, Y[:, l], ls[l], F[l, l], E[l, l]) for l in range(2))
    # diagonal F but different lengthscales per output still couples nothing: the cross blocks are multiplied by F[l,l'] = 0
    assert_close(total, gp.lml_mo(X, Y, ls, F, E), what='sum of variant LMLs')
    r = gp.lml_grad_rbf(X, Y[:, 0], ls[0], F[0, 0], E[0, 0])
    rc = gp.lml_grad_mo(X, Y, ls, F, E)
    assert_close(r['dvariance'], rc['dF'][0, 0], rtol=1e-7, what='dvariance')
    assert_close(r['dnoise'], rc['dE'][0, 0], rtol=1e-7, what='dnoise')
    assert_close(r['dls'], rc['dls'][0], rtol=1e-7, atol=1e-9, what='dls')


def test_check_K_inv_Y_identity():
    """kernel(x, X) K^-1 Y == predicted mean (the reference's MOGP.check_K_inv_Y)."""
    X, Y, ls, F, E = random_problem(30, 4, 3, seed=9, full_F=True)
    Xs = np.random.default_rng(1).normal(size=(5, 4))
    KiY = gp.k_inv_y_mo(X, Y, ls, F, E)
    kernel = gp.gram_mo(Xs, X, ls, F).reshape(3, 5, 3, 30)
    mean = gp.predict_mo(X, Y, ls, F, E, Xs)[0]
    assert_close(np.einsum('loLN,LiN->ol', kernel, KiY), mean, rtol=1e-7, atol=1e-9, what='check_K_inv_Y')
    cho = gp.k_cho_rbf(X, ls, np.diag(F), np.diag(E))
    KiYv = gp.k_inv_y_rbf(X, Y, ls, np.diag(F), np.diag(E))
    for l in range(3):
        K = cho[l] @ cho[l].T
        assert_close(K @ KiYv[l, 0], Y[:, l], rtol=1e-7, atol=1e-9, what='variant K K^-1 y')


def test_sobol_bilinear_form_is_a_gaussian_integral():
    """c^T H c (uncentred) == E_{x_m ~ N(0,1)}[ m_l(x_m) m_j(x_m) ] with m(x_m) the GP mean integrated over the other inputs:
    checked by Gauss-Hermite quadrature for a 1-D and a 2-D marginal."""
    X, Y, ls, F, E = random_problem(12, 3, 2, seed=11, full_E=False)
    KiY = gp.k_inv_y_mo(X, Y, ls, F, E)
    cal = sobol.ClosedSobol(X, ls, F, KiY, True, centre=False)
    lam2 = ls ** 2
    nodes, weights = np.polynomial.hermite_e.hermegauss(80)
    weights = weights / np.sqrt(2 * np.pi)

    def marginal_mean(l, dims, z):
        """m_l at the points z (Q, len(dims)): prod_{m in dims} exp(-(z-X)^2/(2 lam2)) * prod_{m not in dims} sqrt(lam2/(lam2+1)) exp(-X^2/(2(lam2+1)))"""
        rest = [m for m in range(3) if m not in dims]
        k = np.ones((z.shape[0], X.shape[0]))
        for i, m in enumerate(dims):
            k *= np.exp(-0.5 * (z[:, [i]] - X[None, :, m]) ** 2 / lam2[l, m])
        for m in rest:
            k *= np.sqrt(lam2[l, m] / (lam2[l, m] + 1)) * np.exp(-0.5 * X[None, :, m] ** 2 / (lam2[l, m] + 1))
        return F[l, l] * k @ KiY[l, 0]

    z1 = nodes[:, None]
    ref = np.array([[np.sum(weights * marginal_mean(l, [1], z1) * marginal_mean(j, [1], z1)) for j in range(2)] for l in range(2)])
    assert_close(cal._V(1, 2), ref, rtol=1e-9, atol=1e-12, what='1-D marginal')
    zz = np.stack(np.meshgrid(nodes, nodes, indexing='ij'), axis=-1).reshape(-1, 2)
    ww = np.outer(weights, weights).reshape(-1)
    ref2 = np.array([[np.sum(ww * marginal_mean(l, [0, 1], zz) * marginal_mean(j, [0, 1], zz)) for j in range(2)] for l in range(2)])
    assert_close(cal._V(0, 2), ref2, rtol=1e-9, atol=1e-12, what='2-D marginal')


def test_sobol_sweep_semantics_and_subsets():
    X, Y, ls, F, E = random_problem(20, 4, 2, seed=13, full_E=False)
    KiY = gp.k_inv_y_mo(X, Y, ls, F, E)
    out = sobol.sweep(X, ls, F, KiY)
    cal = sobol.ClosedSobol(X, ls, F, KiY)
    assert out[sobol.CLOSED]['S'].shape == (2, 2, 5)
    assert_close(out[sobol.CLOSED]['S'][..., -2], cal.S, what='closed[M-1] is the full model')
    assert_close(out[sobol.TOTAL]['S'][..., 0], cal.S - cal.marginalize((1, 4))['S'], what='total = full - closed(complement)')
    assert_close(np.diagonal(cal.S), np.ones(2), what='S_ll of the full model is 1')
    # a non-contiguous subset through the column permutation == direct masked evaluation
    direct = sobol.subset_V(X, ls, F, KiY, [0, 2])['V']
    perm = [0, 2, 1, 3]
    assert_close(direct, sobol.ClosedSobol(X[:, perm], ls[:, perm], F, KiY).marginalize((0, 2))['V'], what='subset {0,2}')
    # empty slice: H == 1, V = (sum c)(sum c)^T ~ 0 after centring
    assert np.abs(cal.marginalize((4, 4))['V']).max() < 1e-25


def test_literal_transliteration_matches_closed_form_full_F():
    X, Y, ls, F, E = random_problem(9, 3, 2, seed=15, full_F=True)
    KiY = gp.k_inv_y_mo(X, Y, ls, F, np.diag(np.diag(E)))
    lit = literal.closed_sobol_literal(X, ls, F, KiY, False)
    cal = sobol.ClosedSobol(X, ls, F, KiY, False)
    for s in [(0, 3), (1, 2), (0, 1), (2, 3)]:
        assert_close(cal._V(*s), lit['V'](*s), what=f'V{s}')


def test_oracle_test_metrics_match_a_literal_pandas_walk():
    """oracle/normalization.py::test_metrics (the checker of rc_test_metrics) against the reference's own sequence of pandas operations in GPR.test
    (gpr/models.py:241-270: copy the Y block, subtract the predictive mean, divide by the predictive sd, square > 4, any / all per row, then
    sum / count per column) on a frame with the two-level header of test.csv."""
    import pandas as pd
    from oracle import normalization
    rng = np.random.default_rng(8)
    n, L = 57, 3
    truth, mean, sd = rng.normal(size=(n, L)), rng.normal(size=(n, L)) * 0.4, rng.uniform(0.2, 1.2, (n, L))
    cols = pd.MultiIndex.from_tuples([('Y', f'y.{l}') for l in range(L)])
    Y = pd.DataFrame(truth, columns=cols)
    score = Y.copy()
    score.iloc[:] -= mean
    abs_err = abs(score.copy())
    score.iloc[:] /= sd
    outliers = pd.DataFrame(score.to_numpy() ** 2 > 4.0, columns=cols)          # (the reference assigns into a copy of the Y block; pandas 2 refuses the dtype change)
    out_np = outliers.to_numpy(dtype=float)
    any_all = np.column_stack((np.logical_or.reduce(out_np, axis=1), np.logical_and.reduce(out_np, axis=1)))
    rmse = ((abs_err ** 2).sum(axis=0) / abs_err.count(axis=0)) ** 0.5
    mean_sd = pd.DataFrame(sd).sum(axis=0) / n
    flags_ref = np.concatenate([out_np, any_all.astype(float)], axis=1)
    reals, flags, summary = normalization.test_metrics(truth, mean, sd)
    assert_close(reals[:, :L], abs_err.to_numpy(), rtol=1e-14, atol=0, what='Abs Error')
    assert_close(reals[:, L:], score.to_numpy(), rtol=1e-14, atol=0, what='Z Score')
    assert np.array_equal(flags, flags_ref)
    assert_close(summary[:L], rmse.to_numpy(), rtol=1e-14, atol=0, what='RMSE')
    assert_close(summary[L:2 * L], mean_sd.to_numpy(), rtol=1e-14, atol=0, what='mean SD')
    assert_close(summary[2 * L:], flags_ref.sum(axis=0) / n, rtol=1e-14, atol=0, what='outlier fractions')
