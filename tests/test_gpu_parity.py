"""GPU: the CUDA path, called through the C ABI (romcomma._capi -> librc_b200.so), against the CPU oracle on identical seeded
inputs; golden vectors; and size-independent properties at the full benchmark size.  Tolerance: rtol 1e-8 / atol 1e-10
(BASELINE.json north_star) unless a test states why it differs."""
import numpy as np
import pytest
import torch

from conftest import assert_close, random_problem
from oracle import gp, sobol

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def C():
    assert torch.cuda.is_available(), 'these tests need a GPU'
    from romcomma import _capi
    _capi.lib()
    return _capi


def _dense_K(X, ls, F, E):
    return gp.add_noise_mo(gp.gram_mo(X, None, ls, F), E)


# ---- gram -------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('N,M,L,full_F', [(50, 3, 2, True), (64, 1, 1, False), (130, 5, 3, False), (7, 2, 4, True), (257, 20, 2, True)])
def test_gram_matches_oracle(C, N, M, L, full_F):
    X, Y, ls, F, E = random_problem(N, M, L, seed=N, full_F=full_F)
    dX, dls = C.dev(X), C.dev(ls)
    n = L * N
    K = C.gram(dX, None, dls, C.dev(F[None]), C.dev(E[None]))[0, :n, :n].cpu().numpy()
    assert_close(K, _dense_K(X, ls, F, E), what='K')
    Ku = C.gram(dX, None, dls, None, None)[0, :n, :n].cpu().numpy()
    assert_close(Ku.reshape(L, N, L, N), gp.gram_mo_unit(X, None, ls), what='K_unit')
    X2 = np.random.default_rng(1).normal(size=(11, M))
    Kx = C.gram(dX, C.dev(X2), dls, C.dev(F[None]), None)[0, :n, :L * 11].cpu().numpy()
    assert_close(Kx, gp.gram_mo(X, X2, ls, F), what='K(X, X2)')
    # cached-unit-gram branch of MOGPR.KXX
    Kc = C.apply_variance_noise(C.dev(Ku), C.dev(F), C.dev(E), L, N).cpu().numpy()
    assert_close(Kc, _dense_K(X, ls, F, E), what='apply_variance_noise')


def test_gram_variant_batch_matches_gpflow_form(C):
    """Variant path: per-output gpflow kernels.RBF.  The device uses the difference form of r^2, gpflow the expanded
    -2XX^T+|x|^2+|x'|^2 form; they agree to rounding of r^2 (|dK| <~ 1e-15 here)."""
    X, Y, ls, F, E = random_problem(90, 4, 3, seed=3, full_E=False)
    var, noise = np.diag(F).copy(), np.diag(E).copy()
    K = C.gram(C.dev(X), None, C.dev(ls), C.dev(var.reshape(3, 1, 1)), C.dev(noise.reshape(3, 1, 1)), batch=3)[:, :90, :90].cpu().numpy()
    for l in range(3):
        ref = gp.gram_rbf(X, None, ls[l], var[l])
        ref[np.diag_indices(90)] += noise[l]
        assert_close(K[l], ref, what=f'variant K[{l}]')


# ---- factorisation ----------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('n', [5, 128, 200, 384, 700, 1100])
def test_potrf_solves_and_inverse(C, n):
    rng = np.random.default_rng(n)
    A = rng.normal(size=(n, n))
    K = A @ A.T / n + np.eye(n)
    Kp = C.pad_identity(C.dev(K))
    fac = C.Factorization(Kp)
    fac.raise_if_failed()
    Lc = np.linalg.cholesky(K)
    assert_close(fac.lower(n)[0].cpu().numpy(), Lc, rtol=1e-8, atol=1e-10, what='L')
    assert_close(fac.logdet_half().cpu().numpy()[0], np.log(np.diag(Lc)).sum(), what='sum log L_ii')
    y = rng.normal(size=n)
    yp = torch.zeros((1, fac.n_pad), dtype=torch.float64, device='cuda')
    yp[0, :n] = C.dev(y)
    a = fac.trsv(yp)
    assert_close(a[0, :n].cpu().numpy(), np.linalg.solve(Lc, y), rtol=1e-8, atol=1e-10, what='L^-1 y')
    assert_close(fac.trsv(a, transpose=True)[0, :n].cpu().numpy(), np.linalg.solve(K, y), rtol=1e-8, atol=1e-10, what='K^-1 y')
    B = rng.normal(size=(n, 37))
    Bp = torch.zeros((1, fac.n_pad, C.padded(37)), dtype=torch.float64, device='cuda')
    Bp[0, :n, :37] = C.dev(B)
    fac.trsm_fwd_(Bp)
    assert_close(Bp[0, :n, :37].cpu().numpy(), np.linalg.solve(Lc, B), rtol=1e-8, atol=1e-10, what='L^-1 B')
    Kinv = C.extract_lower(fac.inverse_(), n, symmetrize=True)[0].cpu().numpy()
    assert_close(Kinv, np.linalg.inv(K), rtol=1e-8, atol=1e-10, what='K^-1')


@pytest.mark.parametrize('group', [1, 2, 3, 4, 8])
def test_potrf_group_widths(C, group, monkeypatch):
    """RC_POTRF_GROUP block columns are factored between two trailing updates (default 4: rank-512 updates); every width, with a ragged
    last group (n_pad = 1408 = 11 blocks), must give the same factor to rounding."""
    monkeypatch.setenv('RC_POTRF_GROUP', str(group))
    n = 1400
    rng = np.random.default_rng(group)
    A = rng.normal(size=(n, n))
    K = A @ A.T / n + np.eye(n)
    fac = C.Factorization(C.pad_identity(C.dev(K)))
    fac.raise_if_failed()
    assert_close(fac.lower(n)[0].cpu().numpy(), np.linalg.cholesky(K), rtol=1e-8, atol=1e-10, what=f'L (group {group})')


@pytest.mark.parametrize('n,nrhs,sb', [(1300, 1152, None), (1300, 1152, 1), (1300, 1152, 3), (900, 2000, None), (300, 1100, None), (1300, 300, 8)])
def test_trsm_wide_right_hand_sides(C, n, nrhs, sb, monkeypatch):
    """B <- L^-1 B.  More than 1024 columns take the two-level form (K = 1024 updates below an 8-block super-block, RC_TRSM_SB overrides);
    all super-block sizes, ragged last super-blocks and narrow/wide cases against LAPACK."""
    if sb is not None:
        monkeypatch.setenv('RC_TRSM_SB', str(sb))
    rng = np.random.default_rng(n + nrhs)
    A = rng.normal(size=(n, n))
    K = A @ A.T / n + np.eye(n)
    fac = C.Factorization(C.pad_identity(C.dev(K)))
    fac.raise_if_failed()
    B = rng.normal(size=(n, nrhs))
    Bp = torch.zeros((1, fac.n_pad, C.padded(nrhs)), dtype=torch.float64, device='cuda')
    Bp[0, :n, :nrhs] = C.dev(B)
    fac.trsm_fwd_(Bp)
    import scipy.linalg
    ref = scipy.linalg.solve_triangular(np.linalg.cholesky(K), B, lower=True)
    assert_close(Bp[0, :n, :nrhs].cpu().numpy(), ref, rtol=1e-8, atol=1e-10, what='L^-1 B')
    assert float(Bp[0, n:, :].abs().sum()) == 0.0 and float(Bp[0, :, nrhs:].abs().sum()) == 0.0, 'padding must stay zero'


@pytest.mark.parametrize('n,nrhs', [(4096, 128), (4480, 384), (5120, 1024)])
def test_trsm_through_inverted_super_blocks(C, n, nrhs, monkeypatch):
    """rc_trsm_fwd_sbinv (one large factor, few right-hand sides: diagonal super-blocks of 1024 rows inverted once, then one triangular product
    and one rank-1024 update per super-block; n = 4480 leaves a ragged rest of three blocks for the block substitution) against the block
    substitution of rc_trsm_fwd and against L X = B."""
    rng = np.random.default_rng(n)
    G = rng.normal(size=(n, 64))
    A = G @ G.T / 64 + np.diag(rng.uniform(0.5, 2.0, n))
    fac = C.Factorization(C.dev(A)[None].clone())
    fac.raise_if_failed()
    B0 = C.dev(rng.normal(size=(1, n, nrhs)))
    X = fac.trsm_fwd_(B0.clone())
    assert getattr(fac, '_sbwork', None) is not None, 'the super-block path ran'
    monkeypatch.setenv('RC_TRSM_SBINV', '0')
    Xref = fac.trsm_fwd_(B0.clone())
    assert_close(X.cpu().numpy(), Xref.cpu().numpy(), rtol=1e-9, atol=1e-11, what='super-block inverses vs block substitution')
    Lm = C.extract_lower(fac.A, n)[0]
    assert_close((Lm @ X[0]).cpu().numpy(), B0[0].cpu().numpy(), rtol=1e-9, atol=1e-10, what='L X = B')


def test_potrf_batched(C):
    rng = np.random.default_rng(0)
    Ks = []
    for z in range(3):
        A = rng.normal(size=(150, 150))
        Ks.append(A @ A.T / 150 + (z + 1) * np.eye(150))
    fac = C.Factorization(C.pad_identity(C.dev(np.stack(Ks))))
    fac.raise_if_failed()
    Ls = fac.lower(150).cpu().numpy()
    for z in range(3):
        assert_close(Ls[z], np.linalg.cholesky(Ks[z]), what=f'L[{z}]')
    Kinv = C.extract_lower(fac.inverse_(), 150, symmetrize=True).cpu().numpy()
    for z in range(3):
        assert_close(Kinv[z], np.linalg.inv(Ks[z]), what=f'Kinv[{z}]')


def test_potrf_reports_first_bad_pivot(C):
    """A matrix that is not positive definite: the reference raises tf InvalidArgumentError; here info = 1-based pivot index."""
    K = np.eye(300)
    K[200, 200] = -1.0
    fac = C.Factorization(C.pad_identity(C.dev(K)))
    assert int(fac.info.cpu()[0]) == 201
    with pytest.raises(C.RomcommaB200Error):
        fac.raise_if_failed()


# ---- LML + gradient ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('N,M,L,full_F', [(50, 3, 2, True), (200, 5, 3, False), (300, 4, 1, False), (33, 2, 4, True), (129, 8, 2, False)])
def test_lml_grad_covariant(C, N, M, L, full_F):
    X, Y, ls, F, E = random_problem(N, M, L, seed=N + 1, full_F=full_F)
    plan = C.LmlGradPlan(C.dev(X), C.dev(Y), L, 1, C.RC_GRAD_VARIANCE | C.RC_GRAD_LENGTHSCALES)
    res = plan.unpack(plan(C.dev(ls), C.dev(F[None]), C.dev(E[None])).cpu().numpy())[0]
    assert plan.info.cpu().tolist() == [0]
    ref = gp.lml_grad_mo(X, Y, ls, F, E)
    assert_close(res['lml'], ref['lml'], what='lml')
    # gradients are sums of O(n^2) terms of either sign: atol is scaled to the size of the terms (1e-10 * n)
    for k in ('dF', 'dE', 'dls'):
        assert_close(res[k], ref[k], atol=1e-10 * L * N, what=k)
    # value-only call, and the cached-unit-gram branch
    plan0 = C.LmlGradPlan(C.dev(X), C.dev(Y), L, 1, C.RC_GRAD_NONE)
    Ku = C.gram(C.dev(X), None, C.dev(ls), None, None)[0, :L * N, :L * N].contiguous()
    v0 = plan0(C.dev(ls), C.dev(F[None]), C.dev(E[None])).cpu().numpy()[0, 0]
    v1 = plan0(C.dev(ls), C.dev(F[None]), C.dev(E[None]), Ku).cpu().numpy()[0, 0]
    assert_close(v0, ref['lml'], what='lml (value only)')
    assert_close(v1, ref['lml'], what='lml (cached K_unit)')


@pytest.mark.parametrize('N,M,L', [(200, 5, 3), (128, 3, 4), (131, 4, 2), (50, 2, 4), (300, 6, 1)])
def test_lml_grad_selected_inverse(C, N, M, L):
    """RC_GRAD_F_DIAGONAL (default trainables, diagonal F): K^-1 is only formed on its diagonal (l,l) blocks; diag(dF) and the whole
    of dE must still match the oracle, the off-diagonal entries of dF are returned as zero."""
    X, Y, ls, F, E = random_problem(N, M, L, seed=N + 7, full_F=False)
    plan = C.LmlGradPlan(C.dev(X), C.dev(Y), L, 1, C.RC_GRAD_VARIANCE | C.RC_GRAD_F_DIAGONAL)
    res = plan.unpack(plan(C.dev(ls), C.dev(F[None]), C.dev(E[None])).cpu().numpy())[0]
    ref = gp.lml_grad_mo(X, Y, ls, F, E)
    assert_close(res['lml'], ref['lml'], what='lml')
    assert_close(np.diag(res['dF']), np.diag(ref['dF']), atol=1e-10 * L * N, what='diag dF')
    assert_close(res['dE'], ref['dE'], atol=1e-10 * L * N, what='dE')
    assert np.all(res['dF'][~np.eye(L, dtype=bool)] == 0.0)
    assert np.all(res['dls'] == 0.0)


@pytest.mark.parametrize('N,M,L,panels', [(700, 4, 3, 8), (1100, 3, 2, 8), (1024, 5, 2, 4), (600, 3, 4, 2), (900, 2, 3, 16)])
def test_lml_grad_overlapped_inverse(C, N, M, L, panels, monkeypatch):
    """The default gradient path of one matrix runs Z_ii = trtri(L_ii) and T_i = L[below, P_i] Z_ii of every column panel on a side stream
    inside the factorisation (chol.cu: potrf_trtri_lower) - ragged last panels included.  It must match the oracle, the one-stream
    sequence (RC_NO_OVERLAP) to rounding, and itself bit for bit from call to call."""
    monkeypatch.setenv('RC_OVERLAP_PANELS', str(panels))
    X, Y, ls, F, E = random_problem(N, M, L, seed=N + 3, full_F=False)
    args = (C.dev(ls), C.dev(F[None]), C.dev(E[None]))
    flags = C.RC_GRAD_VARIANCE | C.RC_GRAD_LENGTHSCALES
    plan = C.LmlGradPlan(C.dev(X), C.dev(Y), L, 1, flags)
    out1 = plan(*args).cpu().numpy().copy()
    plan.work.fill_(0xFF)   # poisoned scratch (NaN bit patterns) must not matter
    out2 = plan(*args).cpu().numpy().copy()
    assert plan.info.cpu().tolist() == [0]
    assert np.array_equal(out1, out2), 'overlapped evaluation is not bitwise reproducible'
    serial = C.LmlGradPlan(C.dev(X), C.dev(Y), L, 1, flags | C.RC_NO_OVERLAP)
    out_s = serial(*args).cpu().numpy()
    res, res_s = plan.unpack(out1)[0], serial.unpack(out_s)[0]
    ref = gp.lml_grad_mo(X, Y, ls, F, E)
    assert_close(res['lml'], ref['lml'], what='lml')
    for k in ('dF', 'dE', 'dls'):
        assert_close(res[k], ref[k], atol=1e-10 * L * N, what=k)
        assert_close(res[k], res_s[k], atol=1e-10 * L * N, what=k + ' (overlapped vs one stream)')
    # the selected-inverse variant on the same path
    sel = C.LmlGradPlan(C.dev(X), C.dev(Y), L, 1, C.RC_GRAD_VARIANCE | C.RC_GRAD_F_DIAGONAL)
    r2 = sel.unpack(sel(*args).cpu().numpy())[0]
    assert_close(np.diag(r2['dF']), np.diag(ref['dF']), atol=1e-10 * L * N, what='diag dF (selected)')
    assert_close(r2['dE'], ref['dE'], atol=1e-10 * L * N, what='dE (selected)')


def test_lml_grad_variant_batch(C):
    X, Y, ls, F, E = random_problem(180, 6, 3, seed=21, full_E=False)
    var, noise = np.diag(F).copy(), np.diag(E).copy()
    plan = C.LmlGradPlan(C.dev(X), C.dev(Y), 1, 3, C.RC_GRAD_VARIANCE | C.RC_GRAD_LENGTHSCALES)
    res = plan.unpack(plan(C.dev(ls), C.dev(var.reshape(3, 1, 1)), C.dev(noise.reshape(3, 1, 1))).cpu().numpy())
    for l in range(3):
        ref = gp.lml_grad_rbf(X, Y[:, l], ls[l], var[l], noise[l])
        assert_close(res[l]['lml'], ref['lml'], what=f'lml[{l}]')
        assert_close(res[l]['dF'][0, 0], ref['dvariance'], atol=1e-8, what=f'dvariance[{l}]')
        assert_close(res[l]['dE'][0, 0], ref['dnoise'], atol=1e-8, what=f'dnoise[{l}]')
        assert_close(res[l]['dls'][0], ref['dls'], atol=1e-8, what=f'dls[{l}]')


def test_lml_grad_golden(C, golden):
    for name, g in golden.items():
        X, Y, ls, F, E = g['X'], g['Y'], g['ls'], g['F'], g['E']
        L = Y.shape[1]
        plan = C.LmlGradPlan(C.dev(X), C.dev(Y), L, 1, C.RC_GRAD_VARIANCE | C.RC_GRAD_LENGTHSCALES)
        res = plan.unpack(plan(C.dev(ls), C.dev(F[None]), C.dev(E[None])).cpu().numpy())[0]
        assert_close(res['lml'], g['lml'], what=f'{name} lml')
        scale = 1e-10 * max(1.0, np.abs(g['dF']).max(), np.abs(g['dE']).max())
        assert_close(res['dF'], np.tril(g['dF']) + np.tril(g['dF'], -1).T, rtol=1e-7, atol=max(1e-9, scale), what=f'{name} dF')
        assert_close(res['dE'], np.tril(g['dE']) + np.tril(g['dE'], -1).T, rtol=1e-7, atol=max(1e-9, scale), what=f'{name} dE')
        assert_close(res['dls'], g['dls'], rtol=1e-7, atol=max(1e-9, 1e-10 * np.abs(g['dls']).max()), what=f'{name} dls')


def test_lml_not_positive_definite_sets_info(C):
    X, Y, ls, F, E = random_problem(60, 2, 2, seed=2)
    plan = C.LmlGradPlan(C.dev(X), C.dev(Y), 2, 1, C.RC_GRAD_NONE)
    plan(C.dev(ls), C.dev(-F[None]), C.dev(E[None]))
    assert int(plan.info.cpu()[0]) > 0


# ---- Sobol ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('N,M,L,full_F', [(50, 3, 2, True), (200, 5, 3, False), (64, 4, 1, False), (65, 2, 2, True), (10, 6, 2, False)])
def test_sobol_matches_oracle(C, N, M, L, full_F):
    X, Y, ls, F, E = random_problem(N, M, L, seed=N + 5, full_F=full_F, full_E=False)
    KiY = gp.k_inv_y_mo(X, Y, ls, F, E)
    dX, dls = C.dev(X), C.dev(ls)
    for diag in (True, False):
        cal = sobol.ClosedSobol(X, ls, F, KiY, diag)
        Fin = np.diag(F).copy() if diag else F
        Phi, g0, g0KY = C.sobol_prepare(dX, dls, C.dev(Fin), C.dev(KiY.reshape(L, N)), diag)
        assert_close(Phi.cpu().numpy().reshape(cal.Phi.shape), cal.Phi, what='Phi')
        assert_close(g0.cpu().numpy().reshape(cal.g0.shape), cal.g0, what='g0')
        assert_close(g0KY.cpu().numpy().reshape(cal.g0KY.shape), cal.g0KY, what='g0KY')
        slices = [(0, M), (0, 1), (M - 1, M), (0, 2), (1, M), (M, M)] + [(m, m + 1) for m in range(M)]
        V = C.sobol_contract(dX, Phi, g0KY, L, diag, [C.slice_mask(*s) for s in slices]).cpu().numpy()
        for i, s in enumerate(slices):
            assert_close(V[i], cal._V(*s), what=f'V{s} diag={diag}')
        assert_close(V[i], V[i].T, atol=1e-14, what='V symmetric')


def test_sobol_noncontiguous_subsets_and_long_lists(C):
    """Arbitrary subsets (bit masks) and more than 64 subsets per call (chunked launches)."""
    X, Y, ls, F, E = random_problem(70, 7, 2, seed=8, full_E=False)
    KiY = gp.k_inv_y_mo(X, Y, ls, F, E)
    Phi, g0, g0KY = C.sobol_prepare(C.dev(X), C.dev(ls), C.dev(np.diag(F).copy()), C.dev(KiY.reshape(2, 70)), True)
    masks = list(range(1, 128))                     # all 127 non-empty subsets of 7 inputs
    V = C.sobol_contract(C.dev(X), Phi, g0KY, 2, True, masks).cpu().numpy()
    for mask in (0b0000101, 0b1010010, 0b1111111, 0b0100000):
        subset = [m for m in range(7) if (mask >> m) & 1]
        assert_close(V[mask - 1], sobol.subset_V(X, ls, F, KiY, subset)['V'], what=f'subset {subset}')


def test_sobol_golden(C, golden):
    for name, g in golden.items():
        if 'V_diag' not in g:
            continue
        X, ls, F, KiY = g['X'], g['ls'], g['F'], g['KiY']
        L, N = KiY.shape[0], X.shape[0]
        for tag, diag in (('diag', True), ('full', False)):
            Fin = np.diag(F).copy() if diag else F
            Phi, g0, g0KY = C.sobol_prepare(C.dev(X), C.dev(ls), C.dev(Fin), C.dev(KiY.reshape(L, N)), diag)
            V = C.sobol_contract(C.dev(X), Phi, g0KY, L, diag, [C.slice_mask(int(a), int(b)) for a, b in g['slices']]).cpu().numpy()
            assert_close(V, g[f'V_{tag}'], what=f'{name} V {tag}')


# ---- properties at benchmark scale (oracle too slow / too large there) ---------------------------------------------------
@pytest.mark.parametrize('cfg', ['cfg3', 'cfg5', 'cfg4'])
def test_full_size_properties(C, cfg):
    """Full benchmark sizes (cfg3: n = 16384, cfg5: n = 24576, cfg4: n = 32768 = 256 blocks, the only size at which the factorisation
    takes its 8-block-column groups): K K^-1 v = v, L L^T v = K v, the LML and gradient identities from the
    explicit inverse, bitwise run-to-run reproducibility."""
    from romcomma import synthetic
    w = synthetic.config(cfg)
    (N, M), L = w.X.shape, w.Y.shape[1]
    n = N * L
    dX, dY, dls, dF, dE = C.dev(w.X), C.dev(w.Y), C.dev(w.lengthscales), C.dev(w.F[None]), C.dev(w.E[None])
    K = C.gram(dX, None, dls, dF, dE)[0]
    assert (K - K.T).abs().max().item() == 0.0, 'gram is not exactly symmetric'
    fac = C.Factorization(C.gram(dX, None, dls, dF, dE, pad_to=n, pad_identity=True, lower_only=True))
    fac.raise_if_failed()
    v = torch.randn(1, n, dtype=torch.float64, device='cuda', generator=torch.Generator('cuda').manual_seed(0))
    x = fac.trsv(fac.trsv(v), transpose=True)                         # K^-1 v
    resid = (K @ x[0] - v[0]).abs().max().item()
    assert resid < 1e-9, resid
    Lc = fac.lower(n)[0]
    assert ((Lc @ (Lc.T @ v[0])) - K @ v[0]).abs().max().item() < 1e-9
    Kinv = C.extract_lower(fac.inverse_(), n, symmetrize=True)[0]
    assert ((Kinv @ v[0]) - x[0]).abs().max().item() < 1e-8 * x.abs().max().item()
    plan = C.LmlGradPlan(dX, dY, L, 1, C.RC_GRAD_VARIANCE)
    a = plan(dls, dF, dE).clone()
    b = plan(dls, dF, dE).clone()
    assert torch.equal(a, b), 'LML+grad is not bitwise reproducible'
    logdet = Lc.diagonal().log().sum().item()
    alpha = fac  # noqa: F841  (factor has been overwritten by its inverse; recompute the quadratic form from K^-1 v identity instead)
    y = dY.T.reshape(-1)
    quad = (y @ (Kinv @ y)).item()
    assert_close(a[0, 0].item(), -0.5 * quad - 0.5 * n * np.log(2 * np.pi) - logdet, rtol=1e-10, what='LML identity')
    # gradient identity: dLML/dE[l,l] = 1/2 (|a_l|^2 - tr Kinv_ll)
    ky = Kinv @ y
    for l in range(L):
        ref = 0.5 * ((ky[l * N:(l + 1) * N] ** 2).sum() - Kinv[l * N:(l + 1) * N, l * N:(l + 1) * N].diagonal().sum()).item()
        assert_close(a[0, 1 + L * L + l * L + l].item(), ref, rtol=1e-8, atol=1e-6, what=f'dE[{l},{l}]')


@pytest.mark.parametrize('cfg', ['cfg3', 'cfg5', 'cfg4'])
def test_full_size_sobol_properties(C, cfg):
    """Size-independent properties of the Sobol contractions at full benchmark size: V symmetric in the outputs; the sweep form and the
    general-subset kernel agree (a non-structured subset equals a prefix after permuting the inputs); the row-tile parts that the ranks
    of a multi-GPU sweep own add up to the whole; run-to-run bitwise reproducibility."""
    from romcomma import synthetic
    w = synthetic.config(cfg)
    (N, M), L = w.X.shape, w.Y.shape[1]
    dX, Lam, Fd = C.dev(w.X), w.lengthscales, np.diag(w.F).copy()
    KiY = torch.randn(L, N, dtype=torch.float64, device='cuda', generator=torch.Generator('cuda').manual_seed(11)) * 0.1
    Phi, g0, g0KY = C.sobol_prepare(dX, C.dev(Lam), C.dev(Fd), KiY, True)
    slices = [(m, m + 1) for m in range(M)] + [(0, m + 1) for m in range(M)] + [(m + 1, M) for m in range(M)] + [(0, M)]
    masks = [C.slice_mask(*sl) for sl in slices]
    V = C.sobol_contract(dX, Phi, g0KY, L, True, masks)
    V2 = C.sobol_contract(dX, Phi, g0KY, L, True, masks)
    assert torch.equal(V, V2), 'sweep is not bitwise reproducible'
    Vh = V.cpu().numpy()
    scale = np.abs(Vh).max()
    assert np.abs(Vh - Vh.transpose(0, 2, 1)).max() <= 1e-12 * scale, 'V is not symmetric in the outputs'
    assert_close(Vh[2 * M - 1], Vh[3 * M], rtol=1e-12, what='closed [0:M] equals the full slice')
    assert_close(Vh[M], Vh[0], rtol=1e-12, what='closed [0:1] equals first-order [0:1]')
    assert np.abs(Vh[3 * M - 1]).max() <= 1e-9 * scale, 'the empty slice [M:M] must vanish (up to cancellation)'
    # general-subset kernel vs sweep form: {1, 2} is not a prefix/suffix/single; with inputs reordered (1, 2, 0, 3, ...) it is the prefix [0:2]
    perm = [1, 2, 0] + list(range(3, M))
    Vg = C.sobol_contract(dX, Phi, g0KY, L, True, [0b110])
    dXp = C.dev(np.ascontiguousarray(w.X[:, perm]))
    Phip, g0p, g0KYp = C.sobol_prepare(dXp, C.dev(np.ascontiguousarray(Lam[:, perm])), C.dev(Fd), KiY, True)
    Vp = C.sobol_contract(dXp, Phip, g0KYp, L, True, [C.slice_mask(0, 2)])
    assert_close(Vg.cpu().numpy(), Vp.cpu().numpy(), rtol=1e-9, atol=1e-12 * scale, what='general-subset kernel vs sweep form')
    # pair-space parts (what each rank of a 3-GPU sweep would compute) add up
    parts = sum(C.sobol_contract(dX, Phi, g0KY, L, True, masks, part=r, nparts=3) for r in range(3))
    assert_close(parts.cpu().numpy(), Vh, rtol=1e-11, atol=1e-13 * scale, what='row-tile parts add up')


# ---- Sobol errors (ClosedSobolWithError) -----------------------------------------------------------------------------------
@pytest.mark.parametrize('N,M,L,covariant', [(50, 3, 2, True), (130, 5, 3, True), (70, 2, 1, False), (200, 20, 2, False), (257, 12, 3, True), (100, 17, 2, True)])
def test_sobol_error_matches_oracle(C, N, M, L, covariant):
    """rc_sobol_error (V and W for a list of marginal subsets, incl. non-contiguous and empty ones) against oracle/sobol_error.py."""
    from oracle import sobol_error
    X, Y, ls, F, E = random_problem(N, M, L, seed=N + M, full_E=False)
    if covariant:
        KiY, cho = gp.k_inv_y_mo(X, Y, ls, F, E), gp.k_cho_mo(X, ls, F, E)
        K = C.gram(C.dev(X), None, C.dev(ls), C.dev(F[None]), C.dev(E[None]), lower_only=True, pad_to=L * N, pad_identity=True)
    else:
        var, noise = np.diag(F).copy(), np.diag(E).copy()
        KiY, cho = gp.k_inv_y_rbf(X, Y, ls, var, noise), gp.k_cho_rbf(X, ls, var, noise)
        K = C.gram(C.dev(X), None, C.dev(ls), C.dev(var.reshape(L, 1, 1)), C.dev(noise.reshape(L, 1, 1)), batch=L, lower_only=True, pad_to=N,
                   pad_identity=True)
    fac = C.Factorization(K)
    fac.raise_if_failed()
    ref = sobol_error.ClosedSobolWithError(X, ls, np.diag(F), KiY, cho)
    dX, dLam, dF = C.dev(X), C.dev(ls), C.dev(np.diag(F).copy())
    Phi, g0, g0KY = C.sobol_prepare(dX, dLam, dF, C.dev(KiY.reshape(L, N)), True)
    subsets = [list(range(M)), [0], [M - 1], list(range(1, M)), [], [0, M - 1], list(range(0, M, 2))]
    masks = [sum(1 << i for i in set(s)) for s in subsets]
    V, W = C.sobol_error(dX, dLam, dF, Phi, g0, g0KY, fac, masks)
    V, W = V.cpu().numpy(), W.cpu().numpy()
    for k, sub in enumerate(subsets):
        perm = sub + [i for i in range(M) if i not in sub]
        rp = sobol_error.ClosedSobolWithError(X[:, perm], ls[:, perm], np.diag(F), KiY, cho)       # subset as a prefix slice
        out = rp.marginalize((0, len(sub)))
        assert_close(V[k], out['V'], rtol=1e-7, atol=1e-10, what=f'V {sub}')
        assert_close(W[k], out['W'], rtol=1e-7, atol=1e-10, what=f'W {sub}')
    assert_close(W[0], ref.W, rtol=1e-7, atol=1e-10, what='full-model W')


@pytest.mark.parametrize('N,M,L,covariant', [(50, 3, 2, True), (130, 5, 3, True), (70, 2, 1, False), (150, 14, 2, False), (257, 12, 3, True)])
def test_sobol_error_mixed_matches_oracle(C, N, M, L, covariant):
    """rc_sobol_error_mixed (is_T_partial=False: W[mm] and the MIXED covariances W[Mm] with the full model) against oracle/sobol_error.py, for
    structured slices (sweep form), general subsets and M > 12 (no sweep form: the full model's psi factor comes from the general kernel)."""
    from oracle import sobol_error
    X, Y, ls, F, E = random_problem(N, M, L, seed=N + 2 * M, full_E=False)
    if covariant:
        KiY, cho = gp.k_inv_y_mo(X, Y, ls, F, E), gp.k_cho_mo(X, ls, F, E)
        K = C.gram(C.dev(X), None, C.dev(ls), C.dev(F[None]), C.dev(E[None]), lower_only=True, pad_to=L * N, pad_identity=True)
    else:
        var, noise = np.diag(F).copy(), np.diag(E).copy()
        KiY, cho = gp.k_inv_y_rbf(X, Y, ls, var, noise), gp.k_cho_rbf(X, ls, var, noise)
        K = C.gram(C.dev(X), None, C.dev(ls), C.dev(var.reshape(L, 1, 1)), C.dev(noise.reshape(L, 1, 1)), batch=L, lower_only=True, pad_to=N,
                   pad_identity=True)
    fac = C.Factorization(K)
    fac.raise_if_failed()
    dX, dLam, dF = C.dev(X), C.dev(ls), C.dev(np.diag(F).copy())
    Phi, g0, g0KY = C.sobol_prepare(dX, dLam, dF, C.dev(KiY.reshape(L, N)), True)
    subsets = [list(range(M)), [0], [M - 1], list(range(1, M)), [0, M - 1], list(range(0, M, 2))]
    masks = [sum(1 << i for i in set(s)) for s in subsets]
    V, W, WMm = (t.cpu().numpy() for t in C.sobol_error(dX, dLam, dF, Phi, g0, g0KY, fac, masks, mixed=True))
    V1, W1 = (t.cpu().numpy() for t in C.sobol_error(dX, dLam, dF, Phi, g0, g0KY, fac, masks))
    assert np.array_equal(V, V1) and np.array_equal(W, W1), 'the MIXED jobs must not change V and W[mm]'
    ref = sobol_error.ClosedSobolWithError(X, ls, np.diag(F), KiY, cho, is_T_partial=False)
    assert_close(W[0], ref.W_DIAGONAL, rtol=1e-7, atol=1e-10, what='full-model W.DIAGONAL')
    assert_close(WMm[0], ref.W_MIXED, rtol=1e-7, atol=1e-10, what='full-model W.MIXED')
    for k, sub in enumerate(subsets):
        perm = sub + [i for i in range(M) if i not in sub]
        rp = sobol_error.ClosedSobolWithError(X[:, perm], ls[:, perm], np.diag(F), KiY, cho, is_T_partial=False)       # subset as a prefix slice
        out = rp.marginalize((0, len(sub)))
        assert_close(V[k], out['V'], rtol=1e-7, atol=1e-10, what=f'V {sub}')
        assert_close(W[k], out['W'], rtol=1e-7, atol=1e-10, what=f'W {sub}')
        assert_close(WMm[k], out['WMm'], rtol=1e-7, atol=1e-10, what=f'WMm {sub}')


def test_sobol_pair_space_parts_add_up(C):
    """rc_sobol_contract_part: the row-tile parts a multi-GPU sweep all-reduces sum to the single-GPU result, for structured slices
    (sweep form) and general subsets (one exp per subset) alike."""
    N, M, L = 333, 6, 3
    X, Y, ls, F, E = random_problem(N, M, L, seed=21, full_E=False)
    KiY = gp.k_inv_y_mo(X, Y, ls, F, E)
    dX = C.dev(X)
    Phi, g0, g0KY = C.sobol_prepare(dX, C.dev(ls), C.dev(np.diag(F).copy()), C.dev(KiY.reshape(L, N)), True)
    masks = [C.slice_mask(0, M), C.slice_mask(2, 3), C.slice_mask(0, 4), C.slice_mask(3, M), 0, 0b101001, 0b010110, C.slice_mask(1, 4)]
    full = C.sobol_contract(dX, Phi, g0KY, L, True, masks).cpu().numpy()
    cal = sobol.ClosedSobol(X, ls, F, KiY, True)
    assert_close(full[0], cal.V[0], rtol=1e-8, atol=1e-12, what='full model')
    assert_close(full[7], cal._V(1, 4), rtol=1e-8, atol=1e-12, what='interior contiguous slice (general path)')
    assert_close(full[5], sobol.subset_V(X, ls, F, KiY, [0, 3, 5])['V'], rtol=1e-8, atol=1e-12, what='subset {0,3,5}')
    for nparts in (2, 5):
        total = sum(C.sobol_contract(dX, Phi, g0KY, L, True, masks, None, part, nparts).cpu().numpy() for part in range(nparts))
        assert_close(total, full, rtol=1e-10, atol=1e-12, what=f'sum of {nparts} parts (summation order differs)')


@pytest.mark.parametrize('N,M,L', [(100, 15, 2), (70, 20, 1), (130, 11, 3), (64, 1, 2)])
def test_sobol_sweep_form_all_widths(C, N, M, L):
    """Every register-tiling variant of the sweep-form kernel (M <= 4, 8, 12, 20) on the slice families GSA uses, against the oracle."""
    X, Y, ls, F, E = random_problem(N, M, L, seed=7 * N + M, full_E=False)
    KiY = gp.k_inv_y_mo(X, Y, ls, F, E)
    dX = C.dev(X)
    Phi, g0, g0KY = C.sobol_prepare(dX, C.dev(ls), C.dev(np.diag(F).copy()), C.dev(KiY.reshape(L, N)), True)
    slices = [(m, m + 1) for m in range(M)] + [(0, m + 1) for m in range(M)] + [(m + 1, M) for m in range(M)] + [(0, M)]
    V = C.sobol_contract(dX, Phi, g0KY, L, True, [C.slice_mask(*s) for s in slices]).cpu().numpy()
    cal = sobol.ClosedSobol(X, ls, F, KiY, True)
    for k, s in enumerate(slices):
        assert_close(V[k], cal._V(*s), what=f'slice {s}')


@pytest.mark.parametrize('M', list(range(1, 21)))
def test_sobol_sweep_register_form_every_M(C, M):
    """The register form of the sweep kernel is instantiated per M (1..20; two rows per thread up to M = 8, one beyond; three CTAs per SM up to
    M = 12, two beyond): every instantiation,
    with ragged tiles (N = 150: 2.3 tiles) and two outputs, on all structured slices incl. the empty one, against the oracle."""
    N, L = 150, 2
    X, Y, ls, F, E = random_problem(N, M, L, seed=100 + M, full_E=False)
    KiY = gp.k_inv_y_mo(X, Y, ls, F, E)
    dX = C.dev(X)
    Phi, g0, g0KY = C.sobol_prepare(dX, C.dev(ls), C.dev(np.diag(F).copy()), C.dev(KiY.reshape(L, N)), True)
    slices = [(m, m + 1) for m in range(M)] + [(0, m + 1) for m in range(M)] + [(m + 1, M) for m in range(M)] + [(0, M)]
    V = C.sobol_contract(dX, Phi, g0KY, L, True, [C.slice_mask(*s) for s in slices]).cpu().numpy()
    cal = sobol.ClosedSobol(X, ls, F, KiY, True)
    # the empty slice [M:M] (the last of the suffix family) is (sum of the mean-centred coefficients)^2: exactly the rounding of N^2 products
    csum = np.abs(cal.g0KY[:, 0]).sum(axis=1)
    for k, s in enumerate(slices):
        atol = 1e-10 + (16 * np.finfo(float).eps * float(csum.max()) ** 2 if s[0] == s[1] else 0.0)
        assert_close(V[k], cal._V(*s), atol=atol, what=f'M={M} slice {s}')


@pytest.mark.parametrize('form,bound', [(0, 1.0e-15), (1, 6.0e-16)])
def test_device_exp_forms_against_long_double(C, form, bound):
    """The two device exps of the pairwise kernels (polynomial: exp_pairwise; table: exp_tab) against an 80-bit exp: maximum relative error on
    the range the kernels use, the single-FMA reduction of the table form on large arguments (error grows as 3.3e-17 |x|), the clamp below
    -708, NaN propagation, and exactness at 0."""
    rng = np.random.default_rng(5)
    x = np.concatenate([rng.uniform(-40.0, 2.0, 400000), rng.uniform(-1.0, 1.0, 100000), np.linspace(-0.7, 0.7, 100001)])
    y = C.debug_exp(C.dev(x), form).cpu().numpy()
    ref = np.exp(x.astype(np.longdouble))
    rel = np.abs((y.astype(np.longdouble) - ref) / ref).astype(np.float64)
    assert rel.max() <= bound + (3.4e-17 * 40 if form == 1 else 0.0), (form, rel.max())
    inner = np.abs(x) <= 1.0
    assert rel[inner].max() <= bound, (form, rel[inner].max())
    big = rng.uniform(-700.0, 700.0, 100000)
    yb = C.debug_exp(C.dev(big), form).cpu().numpy()
    refb = np.exp(big.astype(np.longdouble))
    relb = np.abs((yb.astype(np.longdouble) - refb) / refb).astype(np.float64)
    assert relb.max() <= bound + (3.4e-17 * 700 if form == 1 else 0.0), (form, relb.max())
    special = np.array([0.0, -708.0, -709.0, -1e4, -1e300, -np.inf, np.nan])
    ys = C.debug_exp(C.dev(special), form).cpu().numpy()
    assert ys[0] == 1.0
    assert np.all(ys[1:6] >= 0.0) and np.all(ys[1:6] < 1e-300)
    assert np.isnan(ys[6])


@pytest.mark.parametrize('N', [300, 700])
def test_lml_grad_plan_cuda_graph_replay_is_bit_identical(C, N):
    """Opt-in CUDA-graph replay of the evaluation (LmlGradPlan(use_graph=True)): same bits as the eager launches, for changing hyper-parameters.
    N=700 (n_pad = 1408, 11 blocks) takes the overlapped potrf+trtri, whose two internal streams are forked and joined inside the capture."""
    X, Y, ls, F, E = random_problem(N, 4, 2, seed=31, full_E=False)
    dX, dY = C.dev(X), C.dev(Y)
    flags = C.RC_GRAD_VARIANCE | C.RC_GRAD_LENGTHSCALES
    eager, graph = C.LmlGradPlan(dX, dY, 2, 1, flags, use_graph=False), C.LmlGradPlan(dX, dY, 2, 1, flags, use_graph=True)
    before = C.launch_count()
    for scale in (1.0, 1.1, 0.9, 1.25):
        args = (C.dev(ls * scale), C.dev(F[None] * scale), C.dev(E[None]))
        a = eager(*args).cpu().numpy().copy()
        b = graph(*args).cpu().numpy().copy()
        assert np.array_equal(a, b), scale
    assert graph._graph is not None
    assert C.launch_count() - before >= 8 * 20, 'replayed launches are counted'


def test_potrf_lookahead_same_bits_as_one_stream(C):
    """The look-ahead factorisation (chol.cu: potrf_lookahead - serial chain on a high-priority stream, trailing updates as yielding launches on
    a low-priority one; taken from 32 blocks up) computes the same tiles with the same K ranges in the same order as the one-stream sequence:
    identical bits, run after run, also against a poisoned workspace; and the right numbers."""
    N, M, L = 1500, 4, 3                                   # n_pad = 4608 = 36 blocks: groups of 2 block columns, 10 split steps + the tail
    X, Y, ls, F, E = random_problem(N, M, L, seed=77, full_E=False)
    dX, dY, args = C.dev(X), C.dev(Y), (C.dev(ls), C.dev(F[None]), C.dev(E[None]))
    for flags in (C.RC_GRAD_NONE, C.RC_GRAD_VARIANCE | C.RC_GRAD_F_DIAGONAL, C.RC_GRAD_VARIANCE | C.RC_GRAD_LENGTHSCALES):
        la, one = C.LmlGradPlan(dX, dY, L, 1, flags), C.LmlGradPlan(dX, dY, L, 1, flags | C.RC_NO_OVERLAP)
        a = la(*args).cpu().numpy().copy()
        b = one(*args).cpu().numpy().copy()
        la.work.fill_(0xFF)
        a2 = la(*args).cpu().numpy().copy()
        assert la.info.cpu().tolist() == [0]
        assert np.array_equal(a, b), f'look-ahead differs from the one-stream sequence (flags {flags})'
        assert np.array_equal(a, a2), 'look-ahead is not reproducible'
    assert_close(a[0, 0], gp.lml_mo(X, Y, ls, F, E), what='LML')
    # the factor itself through rc_potrf (look-ahead by default) against LAPACK
    rng = np.random.default_rng(5)
    n = 4200
    B = rng.normal(size=(n, n))
    K = B @ B.T / n + np.eye(n)
    fac = C.Factorization(C.pad_identity(C.dev(K)))
    fac.raise_if_failed()
    assert_close(fac.lower(n)[0].cpu().numpy(), np.linalg.cholesky(K), rtol=1e-8, atol=1e-10, what='L (look-ahead)')


@pytest.mark.parametrize('N,M,L,diag', [(70, 7, 2, True), (130, 8, 2, True), (65, 5, 3, True), (40, 3, 2, True), (50, 9, 1, True), (45, 7, 2, False)])
def test_sobol_lattice_form_all_subsets(C, N, M, L, diag):
    """The all-subsets sweep: lists that hold at least half of a block of the subset lattice (the 2^min(6,M) subsets sharing their high inputs)
    take the lattice kernel (one exp for the high inputs, one per low input, the low patterns by a depth-first product walk).  Every subset
    against the one-exp-per-subset kernel (reached through sparse lists), sampled ones against the oracle; duplicates, a missing empty set and
    a half-requested block included; the row-tile parts of a multi-GPU sweep add up."""
    X, Y, ls, F, E = random_problem(N, M, L, seed=3 * N + M, full_F=not diag, full_E=False)
    KiY = gp.k_inv_y_mo(X, Y, ls, F, E)
    dX = C.dev(X)
    Fin = np.diag(F).copy() if diag else F
    Phi, g0, g0KY = C.sobol_prepare(dX, C.dev(ls), C.dev(Fin), C.dev(KiY.reshape(L, N)), diag)
    every = list(range(1, 2 ** M))                                      # all non-empty subsets: block 0 lacks the empty one
    V = C.sobol_contract(dX, Phi, g0KY, L, diag, every).cpu().numpy()
    stride = 2 ** M // 4 if M > 3 else 2
    sparse = np.empty_like(V)
    for k in range(stride):                                             # lists with < half of every block: sweep form / general kernel
        idx = [i for i, m in enumerate(every) if m % stride == k]
        sparse[idx] = C.sobol_contract(dX, Phi, g0KY, L, diag, [every[i] for i in idx]).cpu().numpy()
    scale = np.abs(sparse).max()
    assert_close(V, sparse, rtol=1e-9, atol=1e-12 * scale, what='lattice form vs one exp per subset')
    rng = np.random.default_rng(M)
    for mask in [int(m) for m in rng.choice(every, 4, replace=False)] + [2 ** M - 1]:
        subset = [m for m in range(M) if (mask >> m) & 1]
        assert_close(V[mask - 1], sobol.subset_V(X, ls, F, KiY, subset, diag)['V'], what=f'subset {subset}')
    # a duplicate, the empty set, and a block of which exactly half is requested
    odd = [0, 5, 5, 2 ** M - 1] + [m for m in every if m % 2 == 0][: 2 ** (min(6, M) - 1)]
    Vo = C.sobol_contract(dX, Phi, g0KY, L, diag, odd).cpu().numpy()
    for k, mask in enumerate(odd):
        want = sparse[mask - 1] if mask else C.sobol_contract(dX, Phi, g0KY, L, diag, [0]).cpu().numpy()[0]
        assert_close(Vo[k], want, rtol=1e-9, atol=1e-12 * scale, what=f'list entry {k} (mask {mask})')
    parts = sum(C.sobol_contract(dX, Phi, g0KY, L, diag, every, None, r, 3).cpu().numpy() for r in range(3))
    assert_close(parts, V, rtol=1e-10, atol=1e-12 * scale, what='row-tile parts add up (lattice form)')


def test_lml_grad_multi_tiny_problems_many_outputs(C):
    """A batch whose smaller problems pack several outputs into one 64-row tile (found by tools/fuzz_parity.py: the gradient reduction sized its
    per-tile output slots for the largest problem of the batch)."""
    L, M = 4, 4
    probs = [random_problem(N, M, L, seed=90 + N, full_F=True, full_E=False) for N in (64, 9, 23, 3)]
    plan = C.LmlGradMultiPlan([C.dev(p[0]) for p in probs], [C.dev(p[1]) for p in probs], L, C.RC_GRAD_VARIANCE | C.RC_GRAD_LENGTHSCALES)
    res = plan.unpack(plan(C.dev(np.concatenate([p[2] for p in probs])), C.dev(np.stack([p[3] for p in probs])), C.dev(np.stack([p[4] for p in probs]))).cpu().numpy())
    assert plan.info.cpu().tolist() == [0] * 4
    for z, p in enumerate(probs):
        ref = gp.lml_grad_mo(*p)
        assert_close(res[z]['lml'], ref['lml'], what=f'lml[{z}]')
        for k in ('dF', 'dE', 'dls'):
            assert_close(res[z][k], ref[k], atol=1e-10 * L * p[0].shape[0] + 1e-9, what=f'{k}[{z}]')


@pytest.mark.parametrize('L', [1, 2])
def test_lml_grad_multi_independent_problems(C, L):
    """rc_lml_grad_multi: problems with their OWN inputs, outputs and sample counts (folds) in one batched call.  Problems of equal padded size
    are computed exactly as on their own (identical bits to rc_lml_grad on that problem); a smaller problem padded up to the largest one agrees
    to rounding; all against the oracle."""
    M = 4
    Ns = [250, 256, 231, 120]                         # padded sizes (L = 1): 256, 256, 256, 128 -> the last one is padded up to 256 in the batch
    probs = [random_problem(N, M, L, seed=40 + N, full_E=False) for N in Ns]
    flags = C.RC_GRAD_VARIANCE | C.RC_GRAD_LENGTHSCALES
    plan = C.LmlGradMultiPlan([C.dev(p[0]) for p in probs], [C.dev(p[1]) for p in probs], L, flags)
    ls = np.concatenate([p[2] for p in probs], axis=0)
    F, E = np.stack([p[3] for p in probs]), np.stack([p[4] for p in probs])
    out = plan(C.dev(ls), C.dev(F), C.dev(E)).cpu().numpy().copy()
    assert plan.info.cpu().tolist() == [0] * len(Ns)
    res = plan.unpack(out)
    for z, (X, Y, lsz, Fz, Ez) in enumerate(probs):
        ref = gp.lml_grad_mo(X, Y, lsz, Fz, Ez)
        assert_close(res[z]['lml'], ref['lml'], what=f'lml[{z}]')
        for k in ('dF', 'dE', 'dls'):
            assert_close(res[z][k], ref[k], atol=1e-10 * L * Ns[z], what=f'{k}[{z}]')
        single = C.LmlGradPlan(C.dev(X), C.dev(Y), L, 1, flags)
        alone = single(C.dev(lsz), C.dev(Fz[None]), C.dev(Ez[None])).cpu().numpy()[0]
        if C.padded(L * Ns[z]) == C.padded(L * max(Ns)):
            assert np.array_equal(out[z], alone), f'problem {z}: batched and single evaluation differ in bits'
        else:
            assert_close(out[z], alone, rtol=1e-9, atol=1e-9, what=f'problem {z} padded up')
    # value only
    plan0 = C.LmlGradMultiPlan([C.dev(p[0]) for p in probs], [C.dev(p[1]) for p in probs], L, C.RC_GRAD_NONE)
    v = plan0(C.dev(ls), C.dev(F), C.dev(E)).cpu().numpy()[:, 0]
    assert_close(v, [r['lml'] for r in res], rtol=1e-12, what='value-only call')


def test_lml_grad_many_inputs(C):
    """M = 64 inputs (round-1 advice: the gradient reduction was capped at 48 KB of shared memory, i.e. M <= 36-40, while gram / predict take
    M <= 80 and Sobol M <= 64): value and all gradients against the oracle."""
    X, Y, ls, F, E = random_problem(150, 64, 2, seed=64, full_E=False)
    ls = ls * 6.0                                         # 64 inputs: keep the kernel from vanishing
    plan = C.LmlGradPlan(C.dev(X), C.dev(Y), 2, 1, C.RC_GRAD_VARIANCE | C.RC_GRAD_LENGTHSCALES)
    res = plan.unpack(plan(C.dev(ls), C.dev(F[None]), C.dev(E[None])).cpu().numpy())[0]
    ref = gp.lml_grad_mo(X, Y, ls, F, E)
    assert_close(res['lml'], ref['lml'], what='lml (M = 64)')
    for k in ('dF', 'dE', 'dls'):
        assert_close(res[k], ref[k], atol=1e-10 * 300, what=f'{k} (M = 64)')


def test_capi_rejects_bad_arguments(C):
    """Null pointers and unpadded sizes come back as status -2 with a message, not as a launch failure (round-1 advice)."""
    import ctypes
    lib = C.lib()
    A = torch.zeros((1, 256, 256), dtype=torch.float64, device='cuda')
    work = C.workspace(lib.rc_potrf_bufsize(256, 1))
    out = torch.zeros(1, dtype=torch.float64, device='cuda')
    assert lib.rc_logdet(None, 256, 1, C.ptr(out), C.stream_ptr()) == -2
    assert lib.rc_logdet(C.raw_ptr(work), 200, 1, C.ptr(out), C.stream_ptr()) == -2 and b'multiple of 128' in lib.rc_last_error()
    assert lib.rc_trsv(C.ptr(A), 256, 256, 256 * 256, 1, C.raw_ptr(work), None, C.ptr(out), 256, 0, C.stream_ptr()) == -2
    assert lib.rc_trsm_fwd(C.ptr(A), 256, 256, 256 * 256, 0, C.raw_ptr(work), C.ptr(A), 256, 256, 256 * 256, C.stream_ptr()) == -2
    assert lib.rc_potri(C.ptr(A), 250, 256, 256 * 256, 1, C.raw_ptr(work), C.ptr(A), 256, 256 * 256, C.stream_ptr()) == -2
    assert lib.rc_pad_identity(None, 10, 100, C.ptr(A), 256, 256, 256 * 256, 1, C.stream_ptr()) == -2
    assert lib.rc_extract_lower(C.ptr(A), 256, 256 * 256, None, 10, 100, 1, 0, C.stream_ptr()) == -2
    # round 2: the data kernels, the super-block solve and the exp hook
    assert lib.rc_column_stats(C.ptr(A), 1, 4, C.ptr(A), C.stream_ptr()) == -2 and b'fewer than two rows' in lib.rc_last_error()
    assert lib.rc_normalize(C.ptr(A), 10, 5, 4, C.ptr(A), 1e-12, 1, C.ptr(A), C.stream_ptr()) == -2                 # M > columns
    assert lib.rc_normalize(C.ptr(A), 10, 2, 4, C.ptr(A), 1e-12, 0, C.ptr(A), C.stream_ptr()) == -2 and b'direction' in lib.rc_last_error()
    assert lib.rc_normalize(C.ptr(A), 10, 2, 4, C.ptr(A), 0.7, 1, C.ptr(A), C.stream_ptr()) == -2 and b'margin' in lib.rc_last_error()
    assert lib.rc_test_metrics(C.ptr(A), C.ptr(A), None, 10, 2, C.ptr(A), C.ptr(A), C.ptr(A), C.stream_ptr()) == -2
    assert lib.rc_trsm_sbinv_prepare(C.ptr(A), 200, 256, C.raw_ptr(work), C.raw_ptr(work), C.stream_ptr()) == -2
    assert lib.rc_trsm_fwd_sbinv(C.ptr(A), 256, 256, C.raw_ptr(work), C.raw_ptr(work), 128, C.ptr(A), 256, 256, C.stream_ptr()) == -2   # nrhs > nrhs_max
    assert lib.rc_debug_exp(C.ptr(A), C.ptr(A), 16, 2, C.stream_ptr()) == -2
