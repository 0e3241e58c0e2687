"""GPU: the drop-in Python API (romcomma.gpf / gpr / gsa / user.run) end to end against the oracle."""
import random

import numpy as np
import pandas as pd
import pytest
import torch

from conftest import assert_close, random_problem
from oracle import gp, sobol

pytestmark = pytest.mark.gpu


def test_gpf_mogpr_lml_predict_and_dlpack(golden):
    from romcomma.gpf import kernels, models
    g = golden['rand_a']
    X, Y, ls, F, E = g['X'], g['Y'], g['ls'], g['F'], g['E']
    Ed = np.diag(np.diag(E))                                                  # constructor keeps the diagonal only (quirk Q1)
    model = models.MOGPR((X, Y), kernels.RBF(F, ls), noise_variance=E)
    assert (model.L, model.M) == (3, 5)
    assert_close(model.likelihood.variance.value.numpy(), Ed, rtol=1e-12, what='noise re-diagonalised')
    assert_close(model.log_marginal_likelihood().numpy(), gp.lml_mo(X, Y, ls, F, Ed), what='lml')
    assert_close(model.training_loss().numpy(), -gp.lml_mo(X, Y, ls, F, Ed), what='training loss')
    assert_close(model.KXX.numpy(), gp.gram_mo(X, None, ls, F), what='KXX (cached unit gram branch)')
    K4 = model.kernel.K_unit_variance(X)
    assert tuple(K4.shape) == (3, 17, 3, 17)
    assert_close(model.likelihood.add_to(model.KXX).numpy(), gp.add_noise_mo(gp.gram_mo(X, None, ls, F), Ed), what='add_to')
    mean, var = model.predict_f(g['Xs'])
    rm, rv = gp.predict_mo(X, Y, ls, F, Ed, g['Xs'], y_instead_of_f=False)
    assert tuple(mean.shape) == (6, 3)
    assert_close(mean.numpy(), rm, what='f mean')
    assert_close(var.numpy(), rv, what='f var')
    ym, yv = model.predict_y(g['Xs'])
    assert_close(yv.numpy(), rv + np.diag(Ed)[None, :], what='y var')
    # results stay on the device and travel by DLPack
    assert mean.is_cuda
    back = torch.utils.dlpack.from_dlpack(mean.to_dlpack())
    assert back.data_ptr() == mean.data_ptr()
    # the K_d family on the materialised scaled difference (API parity, small inputs)
    d = (torch.as_tensor(X)[None, :, :] / torch.as_tensor(ls)[:, None, :])
    d = d[:, :, None, None, :] - d[None, None, :, :, :]
    assert_close(model.kernel.K_d(d).numpy(), gp.gram_mo(X, None, ls, F), what='K_d')


def test_gpf_training_reaches_the_oracle_optimum():
    """L-BFGS-B on the device LML+gradient: the optimum satisfies the oracle's first-order condition and beats the start."""
    from romcomma import gf_compat as gf
    from romcomma.gpf import kernels, models
    X, Y, ls, F, E = random_problem(60, 2, 2, seed=4, full_E=False)
    model = models.MOGPR((X, Y), kernels.RBF(F, ls), noise_variance=E)
    gf.set_trainable(model.kernel.lengthscales, True)
    start = float(model.log_marginal_likelihood())
    names = [v.name for v in model.trainable_variables]
    assert names == ['KernelLengthscales', 'KernelVariance.cholesky_diagonal', 'KernelVariance.cholesky_lower_triangle',
                     'LikelihoodVariance.cholesky_diagonal', 'LikelihoodVariance.cholesky_lower_triangle']
    # analytic gradient w.r.t. the unconstrained variables == oracle chain rule
    loss, grads = model._loss_and_grad(model.trainable_variables)
    r = gp.lml_grad_mo(X, Y, ls, F, E)
    uF, lowF = gp.variance_pack(F)
    uE, lowE = gp.variance_pack(E)
    dFd, dFl = gp.chain_variance(r['dF'], uF, lowF)
    dEd, dEl = gp.chain_variance(r['dE'], uE, lowE)
    ref = [-(r['dls'] * gp.sigmoid(gp.softplus_inverse(ls))).reshape(2, 1, 2), -dFd, -dFl, -dEd, -dEl]
    for a, b, nm in zip(grads, ref, names):
        assert_close(a, b, atol=1e-8, what=nm)
    res = gf.optimizers.Scipy().minimize(model.training_loss, model.trainable_variables, method='L-BFGS-B', options={'maxiter': 200, 'gtol': 1e-10})
    end = float(model.log_marginal_likelihood())
    assert end > start + 1.0 and res.nit > 3
    lsn, Fn, En = model.kernel.lengthscales_neat.numpy(), model.kernel.variance.value.numpy(), model.likelihood.variance.value.numpy()
    assert_close(end, gp.lml_mo(X, Y, lsn, Fn, En), what='LML at the optimum')
    g = gp.lml_grad_mo(X, Y, lsn, Fn, En)
    assert np.abs(g['dls']).max() < 1e-3 * max(1.0, abs(end))


@pytest.fixture()
def small_repo(tmp_path):
    from romcomma.user import functions, sample
    np.random.seed(1)
    random.seed(1)
    # the DOE draws from scipy's unseeded generator unless told otherwise: fix it, the fit below must be the same problem every run
    fn = sample.Function(tmp_path, lambda N, M: sample.DOE.latin_hypercube(N, M, seed=20261018), functions.ISHIGAMI.subVector('ish', ['standard', 'balanced']), N=120, M=3,
                         noise_variance=sample.GaussianNoise.Variance(2, 0.04, False, False), overwrite_existing=True)
    return fn.repo.into_K_folds(2)


def test_user_run_gpr_and_gsa_layout_and_values(small_repo):
    """user.run.gpr / gsa over folds: variant then covariant model, files in the reference layout, values equal to the oracle
    evaluated at the fitted hyper-parameters read back from the csv files."""
    from romcomma.data.storage import Fold
    from romcomma.gpr.models import MOGP
    from romcomma.gsa.models import GSA
    from romcomma.user import run
    repo = small_repo
    names = run.gpr('gpr', repo, is_read=None, is_covariant=None, is_isotropic=False, maxiter=40)
    assert names == ['gpr.v.a', 'gpr.c.a']
    gsa_names = run.gsa('gpr', repo, is_covariant=None, is_isotropic=False)
    assert [str(p) for p in gsa_names] == ['gpr.v.a/gsa/first_order', 'gpr.v.a/gsa/closed', 'gpr.v.a/gsa/total',
                                           'gpr.c.a/gsa/first_order', 'gpr.c.a/gsa/closed', 'gpr.c.a/gsa/total']
    for k in repo.folds:
        fold = Fold(repo, k)
        for name, cov in (('gpr.v.a', False), ('gpr.c.a', True)):
            folder = fold.folder / name
            for f in ('kernel.csv', 'meta.json', 'kernel/variance.csv', 'kernel/lengthscales.csv', 'likelihood/variance.csv',
                      'likelihood/log_marginal.csv', 'test.csv', 'test_summary.csv', 'gsa/closed/S.csv', 'gsa/total/V.csv', 'gsa/first_order/meta.json'):
                assert (folder / f).exists(), (k, name, f)
            var = pd.read_csv(folder / 'kernel/variance.csv', index_col=0).values
            assert var.shape == ((2, 2) if cov else (1, 2))
            assert pd.read_csv(folder / 'kernel/lengthscales.csv', index_col=0).values.shape == (2, 3)
    for f in ('gpr.v.a/test.csv', 'gpr.c.a/likelihood/variance.csv', 'gpr.v.a/gsa/closed/S.csv', 'gpr.c.a/gsa/total/meta.json'):
        assert (repo.folder / f).exists(), f
    # values: re-read the fitted variant GP of fold 0 and compare every device result with the oracle at the same hyper-parameters
    fold = Fold(repo, 0)
    gpv = MOGP('gpr.v.a', fold, is_read=True, is_covariant=False, is_isotropic=False)
    X, Y = fold.X.values, fold.Y.values
    ls = gpv.kernel.data.frames.lengthscales.np
    var = gpv.kernel.data.frames.variance.np[0]
    noise = gpv.likelihood.data.frames.variance.np[0]
    lml_csv = gpv.likelihood.data.frames.log_marginal.np[0]
    for l in range(2):
        assert_close(lml_csv[l], gp.lml_rbf(X, Y[:, l], ls[l], var[l], noise[l]), rtol=1e-8, what=f'log_marginal.csv[{l}]')
    xs = fold.test_x.values[:7]
    mean, std = gpv.predict(xs)
    for l in range(2):
        rm, rv = gp.predict_rbf(X, Y[:, l], ls[l], var[l], noise[l], xs)
        assert_close(mean[:, l], rm, what='variant mean')
        assert_close(std[:, l], np.sqrt(rv), what='variant std (quirk Q6: predict returns the standard deviation)')
    # GPR.test (device metrics, rc_test_metrics): test.csv and test_summary.csv against the oracle's restatement of gpr/models.py:235-272
    from oracle import normalization
    tx, ty = fold.test_x.values, fold.test_y.values
    tm = np.stack([gp.predict_rbf(X, Y[:, l], ls[l], var[l], noise[l], tx)[0] for l in range(2)], axis=1)
    ts = np.sqrt(np.stack([gp.predict_rbf(X, Y[:, l], ls[l], var[l], noise[l], tx)[1] for l in range(2)], axis=1))
    r_ref, f_ref, s_ref = normalization.test_metrics(ty, tm, ts)
    tcsv = pd.read_csv(fold.folder / 'gpr.v.a' / 'test.csv', header=[0, 1], index_col=0)
    assert_close(tcsv['Mean'].values, tm, what='test.csv Mean')
    assert_close(tcsv['Abs Error'].values, r_ref[:, :2], what='test.csv Abs Error')
    assert_close(tcsv['Z Score'].values, r_ref[:, 2:], what='test.csv Z Score')
    assert np.array_equal(tcsv['Outlier'].values.astype(float), f_ref), 'test.csv Outlier flags'
    scsv = pd.read_csv(fold.folder / 'gpr.v.a' / 'test_summary.csv', header=[0, 1], index_col=0)
    assert_close(scsv.values.reshape(-1), s_ref, what='test_summary.csv (RMSE, SD, outlier fractions)')
    # the fold's normalisation on the device reproduces the normalised data the fold holds (host pandas path, data/storage.py:469-485)
    raw_rows = repo.data.df.loc[fold.data.df.index].to_numpy(dtype=float)
    stats = fold.normalization.stats_of_tensor(torch.tensor(raw_rows).cuda())
    normed = fold.normalization.apply_to_tensor(torch.tensor(raw_rows).cuda(), stats).cpu().numpy()
    assert_close(normed, fold.data.df.to_numpy(dtype=float), what='apply_to_tensor vs the fold csv')
    assert_close(gpv.K_cho.numpy(), gp.k_cho_rbf(X, ls, var, noise), what='K_cho (L,N,N)')
    KiY = gpv.K_inv_Y.numpy()
    assert KiY.shape == (2, 1, X.shape[0])
    assert_close(KiY, gp.k_inv_y_rbf(X, Y, ls, var, noise), rtol=1e-7, atol=1e-8, what='K_inv_Y')
    assert np.all(gpv.check_K_inv_Y(xs) < 1e-9)
    ref = sobol.sweep(X, ls, var, gp.k_inv_y_rbf(X, Y, ls, var, noise), True)
    for kind in GSA.ALL_KINDS:
        S = pd.read_csv(fold.folder / 'gpr.v.a' / 'gsa' / kind.name.lower() / 'S.csv', index_col=[0, 1]).values.reshape(2, 2, 4)
        assert_close(S, ref[int(kind)]['S'], rtol=0, atol=6e-7, what=f'S.csv {kind.name} (csv holds 6 decimals)')
    # covariant GP of fold 1
    fold1 = Fold(repo, 1)
    gpc = MOGP('gpr.c.a', fold1, is_read=True, is_covariant=True, is_isotropic=False)
    X, Y = fold1.X.values, fold1.Y.values
    ls, F, E = gpc.kernel.data.frames.lengthscales.np, gpc.kernel.data.frames.variance.np, gpc.likelihood.data.frames.variance.np
    assert np.count_nonzero(E - np.diag(np.diag(E))) == 0, 'stored likelihood variance is re-diagonalised on read (quirk Q1)'
    mean, std = gpc.predict(xs)
    rm, rv = gp.predict_mo(X, Y, ls, F, E, xs)
    assert_close(mean, rm, what='covariant mean')
    assert_close(std, np.sqrt(rv), what='covariant std')
    assert_close(gpc.K_cho.numpy(), gp.k_cho_mo(X, ls, F, E), what='K_cho (LN,LN)')
    assert_close(gpc.K_inv_Y.numpy(), gp.k_inv_y_mo(X, Y, ls, F, E), rtol=1e-7, atol=1e-8, what='covariant K_inv_Y')
    assert np.all(gpc.check_K_inv_Y(xs) < 1e-9)


def test_closed_sobol_calibrator_attributes(small_repo):
    from romcomma.data.storage import Fold
    from romcomma.gpr.models import MOGP
    from romcomma.gsa.calibrators import ClosedSobol
    fold = Fold(small_repo, 2)
    gpm = MOGP('gp', fold, is_read=False, is_covariant=True, is_isotropic=False,
               likelihood_variance=np.array([[0.02, 0.0], [0.0, 0.03]]))
    gpm.kernel.data.replace(variance=np.array([[1.5, 0.3], [0.3, 0.8]]), lengthscales=np.array([[1.0, 2.0, 0.7], [0.5, 1.5, 2.5]]))
    gpm = MOGP('gp', fold, is_read=True, is_covariant=True, is_isotropic=False)
    X, Y = fold.X.values, fold.Y.values
    ls, F, E = gpm.kernel.data.frames.lengthscales.np, gpm.kernel.data.frames.variance.np, gpm.likelihood.data.frames.variance.np
    KiY = gp.k_inv_y_mo(X, Y, ls, F, E)
    for diag in (True, False):
        cal = ClosedSobol(gpm, is_F_diagonal=diag)
        ref = sobol.ClosedSobol(X, ls, F, KiY, diag)
        assert cal.is_F_diagonal == diag
        for j in range(3):
            assert_close(cal.Lambda2[1][j].numpy(), ref.Lambda2[1][j], what='Lambda2')
            assert_close(cal.Lambda2[-1][j].numpy(), ref.Lambda2[-1][j], what='Lambda2^-1')
        assert_close(cal.Phi.numpy(), ref.Phi, what='Phi')
        assert_close(cal.g0.numpy(), ref.g0, what='g0')
        assert_close(cal.g0KY.numpy(), ref.g0KY, rtol=1e-7, atol=1e-9, what='g0KY')
        assert_close(cal.G.numpy(), ref.G, what='G')
        for j in range(3):
            assert_close(cal.V[j].numpy(), ref.V[j], rtol=1e-7, atol=1e-9, what=f'V[{j}]')
        assert_close(cal.S.numpy(), ref.S, rtol=1e-7, atol=1e-9, what='S')
        out = cal.marginalize((1, 3))
        assert_close(out['V'].numpy(), ref.marginalize((1, 3))['V'], rtol=1e-7, atol=1e-9, what='marginalize V')
        assert_close(out['S'].numpy(), ref.marginalize((1, 3))['S'], rtol=1e-7, atol=1e-9, what='marginalize S')
        subs = cal.marginalize_subsets([[0, 2], [1]])
        assert_close(subs[0]['V'].numpy(), sobol.subset_V(X, ls, F, KiY, [0, 2], diag)['V'], rtol=1e-7, atol=1e-9, what='subset {0,2}')
        # total index of a subset = full - closed(complement) (gsa/models.py:207-210): for the prefix {0} it is what the TOTAL kind reports for m = 0
        tot = cal.marginalize_subsets([[0], [0, 2]], total=True)
        assert_close(tot[0]['S'].numpy(), ref.S - ref.marginalize((1, 3))['S'], rtol=1e-7, atol=1e-9, what='total index of {0}')
        assert_close(tot[1]['S'].numpy(), ref.S - sobol.subset_V(X, ls, F, KiY, [1], diag)['S'], rtol=1e-7, atol=1e-9, what='total index of {0,2}')
        assert_close(tot[1]['V'].numpy(), sobol.subset_V(X, ls, F, KiY, [1], diag)['V'], rtol=1e-7, atol=1e-9, what='V of the complement')
    # default is_F_diagonal: True unless the GP's meta says the kernel covariance was trained (quirk Q3)
    assert ClosedSobol(gpm).is_F_diagonal is True


def test_mogp_shares_one_factorisation_until_the_hyperparameters_change(small_repo):
    """MOGP keeps ONE Cholesky factor (and K^-1 y) per set of hyper-parameters for predict / K_cho / K_inv_Y / the Sobol calibrators, where
    the reference factorises in each of them; the cache must follow the parameters: same numbers as a fresh model before and after a
    change, and a different factor object after it."""
    from romcomma.data.storage import Fold
    from romcomma.gpr.models import MOGP
    fold = Fold(small_repo, 0)
    model = MOGP('cache.c.a', fold, is_read=False, is_covariant=True, is_isotropic=False)
    xs = fold.test_x.values[:5]
    fac0 = model._factorize()[0]
    m0, s0 = model.predict(xs)
    assert model._factorize()[0] is fac0, 'unchanged hyper-parameters: the factor is reused'
    kiy0 = model.K_inv_Y.numpy().copy()
    assert np.array_equal(model.K_inv_Y.numpy(), kiy0)
    X, Y = fold.X.values, fold.Y.values
    impl = model.implementation[0]
    ls, F, E = impl.kernel.lengthscales_neat.numpy(), impl.kernel.variance.value.numpy(), impl.likelihood.variance.value.numpy()
    rm, rv = gp.predict_mo(X, Y, ls, F, E, xs)
    assert_close(m0, rm, what='mean (cached factor)')
    assert_close(s0, np.sqrt(rv), what='std (cached factor)')
    # change a hyper-parameter in place, as the optimiser does
    v = impl.likelihood.variance._cholesky_diagonal
    v.unconstrained_variable = np.asarray(v.unconstrained_variable) + 0.3
    fac1 = model._factorize()[0]
    assert fac1 is not fac0, 'changed hyper-parameters: a new factorisation'
    E1 = impl.likelihood.variance.value.numpy()
    assert not np.allclose(E1, E)
    m1, s1 = model.predict(xs)
    rm1, rv1 = gp.predict_mo(X, Y, ls, F, E1, xs)
    assert_close(m1, rm1, what='mean after the change')
    assert_close(s1, np.sqrt(rv1), what='std after the change')
    assert_close(model.K_inv_Y.numpy(), gp.k_inv_y_mo(X, Y, ls, F, E1), rtol=1e-7, atol=1e-8, what='K_inv_Y after the change')


def test_installation_test_flow_not_partial(tmp_path):
    """The flow of the reference's installation_test.py (:33-93): Oakley 2004 (L = 3) on M = 7 inputs, N = 300, K = 2 folds (+ improper), variant
    GPs fitted isotropic then anisotropic, the three Sobol kinds WITH errors and is_T_partial=False (IS_GSA_ERROR_PARTIAL = False) - through
    user.run.gpr / user.run.gsa / user.results.Collect; T and W of one fold against the oracle at the fitted hyper-parameters."""
    from oracle import sobol_error
    from romcomma.data.storage import Fold
    from romcomma.gpr.models import MOGP
    from romcomma.gsa.models import GSA
    from romcomma.user import functions, results, run, sample
    np.random.seed(3)
    random.seed(3)
    noise = sample.GaussianNoise.Variance(len(functions.OAKLEY2004), 0.04, False, True)
    repo = sample.Function(tmp_path, lambda N, M: sample.DOE.latin_hypercube(N, M, seed=3), functions.OAKLEY2004, 300, 7, noise, None, True).repo
    repo = repo.into_K_folds(2).rotate_folds(None)
    models = run.gpr(name='gpr', repo=repo, is_read=False, is_covariant=False, is_isotropic=None, ignore_exceptions=False, maxiter=25)
    assert models == ['gpr.v.i', 'gpr.v.a']
    results.Collect({'test': {'header': [0, 1]}, 'test_summary': {'header': [0, 1], 'index_col': 0}},
                    {repo.folder / model: {'model': model} for model in models}, False).from_folders(repo.folder / 'gpr', True)
    run.gsa('gpr', repo, is_covariant=False, is_isotropic=False, kinds=GSA.ALL_KINDS, is_error_calculated=True, ignore_exceptions=False,
            is_T_partial=False)
    kind_names = [kind.name.lower() for kind in GSA.ALL_KINDS]
    results.Collect({'S': {}, 'V': {}, 'T': {}, 'W': {}}, {f'{repo.folder / model}/gsa/{k}': {'model': model, 'kind': k} for k in kind_names for model in models},
                    True).from_folders(repo.folder / 'gsa', True)
    for f in ('gpr/test_summary.csv', 'gsa/T.csv', 'gsa/W.csv', 'gpr.v.a/gsa/total/T.csv'):
        assert (repo.folder / f).exists(), f
    fold = Fold(repo, 0)
    gpv = MOGP('gpr.v.a', fold, is_read=True, is_covariant=False, is_isotropic=False)
    X, Y = fold.X.values, fold.Y.values
    ls, var, noise_v = gpv.kernel.data.frames.lengthscales.np, gpv.kernel.data.frames.variance.np[0], gpv.likelihood.data.frames.variance.np[0]
    ref = sobol_error.ClosedSobolWithError(X, ls, var, gp.k_inv_y_rbf(X, Y, ls, var, noise_v), gp.k_cho_rbf(X, ls, var, noise_v), is_T_partial=False)
    from romcomma.gsa.models import Sobol
    for kind in GSA.ALL_KINDS:
        want = sobol_error.sobol_kind_with_error(ref, int(kind))
        folder = fold.folder / 'gpr.v.a' / 'gsa' / kind.name.lower()
        for key in ('S', 'V', 'W'):
            got = pd.read_csv(folder / f'{key}.csv', index_col=[0, 1]).values.reshape(3, 3, -1)
            assert_close(got, want[key], rtol=0, atol=6e-7, what=f'{kind.name} {key}.csv (6 decimals)')
        T_csv = pd.read_csv(folder / 'T.csv', index_col=[0, 1]).values.reshape(3, 3, -1)
        assert T_csv.shape == want['T'].shape == want['S'].shape, 'T carries the full-model column when not partial'
        # full precision, not through the csv.  T = sqrt(|Q_m|)/V2 with Q_m = W[mm] - 2 V W[Mm]/V1 + V^2 Q: three terms of either sign that largely
        # cancel (the full model's diagonal cancels to rounding noise), so |Q_m| = T^2 V4 is compared with an atol of 1e-7 of the terms' magnitude.
        gsa = Sobol(gpv, kind, m=-1, is_error_calculated=True, is_T_partial=False)
        gsa.calibrate()
        got = {k: np.asarray(v, dtype=float) for k, v in gsa.results.items()}
        for key in ('S', 'V'):
            assert_close(got[key], want[key], rtol=1e-7, atol=1e-9, what=f'{kind.name} {key}')
        # W = mu_phi_mu - mu_psi_mu and Q_m = W[mm] - 2 V W[Mm]/V1 + V^2 Q are differences of nearly equal terms (here |W| ~ 1e-4 of them), each
        # quadratic in K^-1 y, whose relative rounding error is eps * cond(K) ~ 1e-16 * 1e6 for this fitted model (noise 3e-3, variance 3): any
        # two float64 evaluations - the oracle and the reference's own code included - differ by ~1e-10 of the TERMS, and each term is itself a sum
        # c^T Q c over sample pairs with coefficients of either sign (sum |c||Q||c| ~ 1e2..1e4 times the term).  Measured on the B200: 1.2e-8 of the
        # terms; the reference's own code against the oracle on the same kind of model: 3e-6 of |Q_m|.  Tolerance: 1e-7 of the terms.
        per_slice = [ref.marginalize(s) for s in sobol.m_slices(int(kind), 7)]
        W_scale, Q_scale = (np.stack([r[key] for r in per_slice], axis=-1) for key in ('W_scale', 'Q_scale'))
        err = np.abs(got['W'] - want['W'])
        assert np.all(err <= 1e-7 * W_scale + 1e-10), f'{kind.name} W: worst err/terms {np.max(err / W_scale):.2e}'      # + 1e-10: the empty slice [M:M] of TOTAL is (sum c)^2 - ... = cancellation noise of mean-centred coefficients
        V4 = ref.V[4][..., None]
        T_slices = (got['T'][..., :-1] - got['T'][..., -1:]) if kind == GSA.Kind.TOTAL else got['T'][..., :-1]      # models.py:212 adds the full-model T
        T_want = (want['T'][..., :-1] - want['T'][..., -1:]) if kind == GSA.Kind.TOTAL else want['T'][..., :-1]
        err = np.abs(T_slices ** 2 - T_want ** 2) * V4
        assert np.all(err <= 1e-7 * Q_scale + 1e-10), f'{kind.name} |Q_m|: worst err/terms {np.max(err / Q_scale):.2e}'
        assert_close(T_csv, got['T'], rtol=0, atol=6e-7, what=f'{kind.name} T.csv holds the computed T to 6 decimals')


def test_lockstep_fits_equal_sequential_fits(small_repo, monkeypatch):
    """Folds (and the outputs of a variant GP) fitted side by side with batched evaluations (romcomma.lockstep) follow the same L-BFGS-B
    trajectories as the reference's one-after-another loop: identical csv files for the variant model, and the batching really happens."""
    from romcomma import lockstep
    from romcomma.user import run
    repo = small_repo
    sizes = []
    original = lockstep.EvaluationBroker._launch

    def spy(self, key, group, stream_index):
        sizes.append(len(group))
        return original(self, key, group, stream_index)
    monkeypatch.setattr(lockstep.EvaluationBroker, '_launch', spy)
    names = run.gpr('together', repo, is_read=False, is_covariant=None, is_isotropic=False, maxiter=30)
    assert names == ['together.v.a', 'together.c.a']
    assert max(sizes) >= 4, f'3 folds x 2 outputs should share launches, batch sizes seen: {sorted(set(sizes))}'
    monkeypatch.setenv('ROMCOMMA_B200_LOCKSTEP', '0')
    run.gpr('sequential', repo, is_read=False, is_covariant=None, is_isotropic=False, maxiter=30)
    for k in repo.folds:
        for model in ('v.a', 'c.a'):
            for csv in ('kernel/lengthscales.csv', 'kernel/variance.csv', 'likelihood/variance.csv', 'likelihood/log_marginal.csv', 'test.csv'):
                a = pd.read_csv(repo.fold_folder(k) / f'together.{model}' / csv, index_col=0, header=[0, 1] if csv == 'test.csv' else 0).values
                b = pd.read_csv(repo.fold_folder(k) / f'sequential.{model}' / csv, index_col=0, header=[0, 1] if csv == 'test.csv' else 0).values
                if model == 'v.a':
                    assert np.array_equal(a, b), f'fold {k} {model} {csv}: lock-step and sequential fits differ'
                else:   # the covariant fit on its own uses the selected inverse (default trainables); in a batch the full inverse: same to rounding
                    assert_close(a, b, rtol=1e-6, atol=1e-8, what=f'fold {k} {model} {csv}')


def test_gsa_base_gaussian_reproduces_the_closed_sobol_chain():
    """romcomma.gsa.base.Gaussian / diag_det keep the reference's broadcasting contract (gsa/base.py:52-126): the Gaussian-ratio chain of
    ClosedSobol._calibrate / _V (gsa/calibrators.py:60-92) written with them, as a user of the reference would, gives the oracle's g0 and V - and
    the numbers the fused kernel (rc_sobol_contract) produces."""
    from romcomma.gsa.base import Gaussian, diag_det, sym_check
    X, Y, ls, F, E = random_problem(24, 3, 2, seed=9, full_E=False)
    KiY = gp.k_inv_y_mo(X, Y, ls, F, E)
    ref = sobol.ClosedSobol(X, ls, F, KiY, True)
    Xd, Lam = torch.as_tensor(X, device='cuda'), torch.as_tensor(ls, device='cuda')
    L2p1 = (Lam * Lam)[:, None, :] + 1.0                                    # Lambda2[1][1], (L,1,M)
    Phi = 1.0 / L2p1
    pre = torch.sqrt(torch.prod((Lam * Lam)[:, None, :] * Phi, dim=-1)) * torch.as_tensor(np.diag(F).reshape(2, 1).copy(), device='cuda')
    g0 = torch.exp(Gaussian(mean=Xd[None, None, ...], variance=L2p1, is_variance_diagonal=True, LBunch=2).exponent) * pre[..., None]
    assert_close(g0.cpu().numpy(), ref.g0, what='g0 through gsa.base.Gaussian')
    g0KY = torch.as_tensor(ref.g0KY, device='cuda')
    G = torch.einsum('lLM,NM->lLNM', Phi, Xd)

    def V(m0, m1):
        g, phi = G[..., m0:m1], Phi[..., m0:m1]
        Gamma = 1 - phi
        Psi = Gamma[:, :, None, None, :] + Gamma[None, None, ...] - torch.einsum('lLM,jJM->lLjJM', Gamma, Gamma)
        PsiPhi = torch.einsum('lLjJM,lLM->lLjJM', Psi, phi)
        PhiG = torch.einsum('lLM,jJnM->lLjJnM', phi, g).unsqueeze(2)
        PhiGauss = Gaussian(mean=g, variance=phi, is_variance_diagonal=True, LBunch=2)
        H = Gaussian(mean=PhiG, variance=PsiPhi, ordinate=g[..., None, None, None, :], is_variance_diagonal=True, LBunch=2)
        H = H / PhiGauss.expand_dims([-1, -2, -3])
        assert np.array_equal(H.det.cpu().numpy(), diag_det(H.cho_diag).cpu().numpy()), 'det == diag_det(cho_diag)'
        return torch.einsum('lLN,lLNjJn,jJn->lj', g0KY, H.pdf, g0KY)
    for s in ((0, 3), (1, 2), (0, 2)):
        Vs = V(*s)
        assert_close(Vs.cpu().numpy(), ref._V(*s), what=f'V{s} through gsa.base.Gaussian')
        assert float(sym_check(Vs, [1, 0]).cpu()) < 1e-24
    # the full (non-diagonal) variance branch: a 2 x 2 covariance against the explicit quadratic form
    S = torch.tensor([[2.0, 0.3], [0.3, 1.0]], dtype=torch.float64, device='cuda')
    z = torch.tensor([[0.5, -1.0]], dtype=torch.float64, device='cuda')
    gfull = Gaussian(mean=z, variance=S, is_variance_diagonal=False)
    want = -0.5 * (z @ torch.linalg.inv(S) @ z.T).item()
    assert_close(gfull.exponent.cpu().numpy().reshape(-1)[0], want, what='full-covariance exponent')
    assert_close(gfull.det.cpu().numpy().reshape(-1)[0], np.sqrt(np.linalg.det(S.cpu().numpy())), what='sqrt det')


def test_predict_gradient_and_full_covariance_against_oracle(small_repo):
    """MOGP.predict_gradient (variant GP: Jacobian kernel, TRSM, -W^T W on tensor-core tiles, assembly - all in the library) and predict_f with
    full_cov / full_output_cov (Knn - A^T A through rc_syrk_tn) against the oracle; shapes as the reference returns them."""
    from romcomma.data.storage import Fold
    from romcomma.gpf import kernels, models
    from romcomma.gpr.models import MOGP
    fold = Fold(small_repo, 0)
    gpv = MOGP('pg.v.a', fold, is_read=False, is_covariant=False, is_isotropic=False)
    gpv.kernel.data.replace(variance=np.array([[1.3, 0.7]]), lengthscales=np.array([[0.8, 1.7, 1.1], [2.0, 0.6, 1.4]]))
    gpv.likelihood.data.replace(variance=np.array([[0.02, 0.05]]))
    gpv = MOGP('pg.v.a', fold, is_read=True, is_covariant=False, is_isotropic=False)
    X, Y = fold.X.values, fold.Y.values
    ls, var, noise = gpv.kernel.data.frames.lengthscales.np, gpv.kernel.data.frames.variance.np[0], gpv.likelihood.data.frames.variance.np[0]
    for o in (1, 5, 50):                                            # 50 x 3 = 150 columns: more than one 128-tile of right-hand sides
        xs = fold.test_x.values[:o] * 0.9 + 0.02
        mean, cov = gpv.predict_gradient(xs)
        rm, rv = gp.predict_gradient_rbf(X, Y, ls, var, noise, xs)
        assert tuple(mean.shape) == (o, 2, 3) and tuple(cov.shape) == (o, o, 2, 3, 3)
        assert_close(mean.numpy(), rm, rtol=1e-7, atol=1e-9, what=f'predict_gradient mean (o={o})')
        assert_close(cov.numpy(), rv, rtol=1e-7, atol=1e-9, what=f'predict_gradient covariance (o={o})')
    # full predictive covariance of the covariant model
    Xr, Yr, lsr, F, E = random_problem(70, 3, 2, seed=12, full_F=True, full_E=False)
    model = models.MOGPR((Xr, Yr), kernels.RBF(F, lsr), noise_variance=E)
    xs = Xr[:9] + 0.1
    Kmm = gp.add_noise_mo(gp.gram_mo(Xr, None, lsr, F), E)
    Kmn, Knn = gp.gram_mo(Xr, xs, lsr, F), gp.gram_mo(xs, xs, lsr, F)
    A = np.linalg.solve(np.linalg.cholesky(Kmm), Kmn)
    full = (Knn - A.T @ A).reshape(2, 9, 2, 9)                       # [L, N, l, n]
    mean, v4 = model.predict_f(xs, full_cov=True)
    rm, rv = gp.predict_mo(Xr, Yr, lsr, F, E, xs, y_instead_of_f=False)
    assert tuple(v4.shape) == (9, 9, 2, 2)
    assert_close(mean.numpy(), rm, what='mean (full_cov)')
    assert_close(v4.numpy(), full.transpose(3, 1, 2, 0), what='full_cov: (n, N, l, L) = reversed "LNln -> LlNn"')
    _, v3 = model.predict_f(xs, full_output_cov=True)
    assert tuple(v3.shape) == (9, 2, 2)
    assert_close(v3.numpy(), np.stack([full[:, i, :, i].T for i in range(9)]), what='full_output_cov: (n, l, L)')
    assert_close(np.stack([np.diag(v3.numpy()[i]) for i in range(9)]), rv, what='its diagonal is the marginal variance')
    with pytest.raises(NotImplementedError):
        model.predict_y(xs, full_cov=True)
    # the variant GP (gpflow GPR shapes): [1, n, n] and [n, 1, 1]
    gpr0 = gpv.implementation[0]
    _, vc = gpr0.predict_f(xs0 := fold.test_x.values[:6], full_cov=True)
    Kv = gp.gram_rbf(X, None, ls[0], var[0]) + noise[0] * np.eye(X.shape[0])
    Av = np.linalg.solve(np.linalg.cholesky(Kv), gp.gram_rbf(X, xs0, ls[0], var[0]))
    assert tuple(vc.shape) == (1, 6, 6)
    assert_close(vc.numpy()[0], gp.gram_rbf(xs0, xs0, ls[0], var[0]) - Av.T @ Av, what='variant full covariance')


# ---- fold data on the device: normalisation and GPR.test's numbers ---------------------------------------------------------
@pytest.mark.parametrize('name', ['ref_cov_a', 'ref_var_a', 'ref_cov_b', 'ref_var_b'])
def test_device_normalisation_reproduces_the_reference_run(name):
    """rc_column_stats + rc_normalize on the raw rows of the fold the reference fitted: the X and Y the reference's own Normalization produced
    (tests/golden/ref_*.npz), the oracle's statistics, and undo_from back to the raw rows where the clip did not bite."""
    from conftest import GOLDEN
    from oracle import normalization
    from romcomma import _capi as C
    g = np.load(GOLDEN / f'{name}.npz')
    raw, M = g['raw'], g['X'].shape[1]
    import importlib.util
    spec = importlib.util.spec_from_file_location('tests_golden_scenario_for_normalisation', GOLDEN / 'scenario.py')
    scenario = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(scenario)
    K = scenario.CASES[name]['K']
    rows = raw[g[f'fold.{K if K > 0 else 0}.train_index']]                     # the fold the scenario fits: the improper one when there is one
    assert rows.shape[0] == g['X'].shape[0]
    d = C.dev(rows)
    st = C.column_stats(d)
    assert_close(st.cpu().numpy(), normalization.stats(rows), rtol=1e-12, atol=1e-14, what='mean, std, rng, min, max')
    out = C.normalize(d, M, st)
    assert_close(out.cpu().numpy()[:, :M], g['X'], what='normalised X vs the reference run')
    assert_close(out.cpu().numpy()[:, M:], g['Y'], what='normalised Y vs the reference run')
    back = C.normalize(out, M, st, undo=True).cpu().numpy()
    ref_back = normalization.undo_from(normalization.apply_to(rows, M, normalization.stats(rows)), M, normalization.stats(rows))
    assert_close(back, ref_back, what='undo_from')


def test_device_normalisation_clips_and_handles_large_folds():
    """20000 x 13 samples with outliers beyond [min, max]: the clip at UNIFORM_MARGIN (probit = -+7.03), against the oracle."""
    from oracle import normalization
    from romcomma import _capi as C
    rng = np.random.default_rng(11)
    data = rng.normal(size=(20000, 13)) * rng.uniform(0.1, 50.0, 13) + rng.uniform(-5, 5, 13)
    data[::97, :5] *= 40.0
    st = normalization.stats(data)
    out = C.normalize(C.dev(data), 10, C.column_stats(C.dev(data))).cpu().numpy()
    ref = normalization.apply_to(data, 10, st)
    assert np.abs(ref[:, :10]).max() > 7.0, 'the clip is exercised'
    assert_close(out, ref, what='apply_to')


@pytest.mark.parametrize('n,L', [(1, 1), (205, 1), (300, 3), (2048, 4)])
def test_device_test_metrics(n, L):
    from oracle import normalization
    from romcomma import _capi as C
    rng = np.random.default_rng(n + L)
    truth, mean = rng.normal(size=(n, L)), rng.normal(size=(n, L)) * 0.3
    sd = rng.uniform(0.2, 1.5, (n, L))
    reals, flags, summary = (t.cpu().numpy() for t in C.test_metrics(C.dev(truth), C.dev(mean), C.dev(sd)))
    r_ref, f_ref, s_ref = normalization.test_metrics(truth, mean, sd)
    assert_close(reals, r_ref, rtol=1e-13, atol=0, what='abs error, z score')
    assert np.array_equal(flags, f_ref), 'outlier flags'
    assert_close(summary, s_ref, rtol=1e-12, atol=1e-15, what='RMSE, mean SD, outlier fractions')
