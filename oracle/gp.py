"""Oracle (float64 numpy/scipy) for gram, LML, gradients and prediction.  TEST INFRASTRUCTURE ONLY.

Two model families, exactly as the reference dispatches them (romcomma/gpr/models.py:332-343):

* covariant ("mo"): one (LN,LN) gram, romcomma/gpf/kernels.py:74-113,153-154 + gpf/likelihoods.py:64-67 +
  gpf/models.py:73-111.  Index (l,n) -> l*N+n (gpf/kernels.py:103-104), y = vec(Y^T) (gpf/models.py:130).
* variant ("rbf"): L independent gpflow ``GPR(kernels.RBF, likelihoods.Gaussian)`` models
  (gpr/models.py:340-342, gpr/kernels.py:176-177); gpflow's ``square_distance`` form of r^2.

Hyper-parameter transforms follow gpflow ``positive()`` (softplus, optionally shifted), see
romcomma/gpf/base.py:88-94 and gpflow.likelihoods.Gaussian (lower bound 1e-6).
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla

LOG2PI = np.log(2.0 * np.pi)


# ----------------------------------------------------------------------------------------------------------------
# transforms (gpflow.utilities.positive == tfp Softplus [+ Shift])
# ----------------------------------------------------------------------------------------------------------------
def softplus(u):
    u = np.asarray(u, dtype=np.float64)
    return np.logaddexp(0.0, u)


def softplus_inverse(y):
    y = np.asarray(y, dtype=np.float64)
    return y + np.log(-np.expm1(-y))


def sigmoid(u):
    u = np.asarray(u, dtype=np.float64)
    return 0.5 * (1.0 + np.tanh(0.5 * u))


def variance_unpack(u_diag, lower, floor=1e-3):
    """gpf/base.py:42-55 - Variance.cholesky / .value from (unconstrained diagonal, strict lower triangle row-major)."""
    L = len(u_diag)
    C = np.zeros((L, L))
    C[np.tril_indices(L, -1)] = np.asarray(lower, dtype=np.float64)  # row-major strict lower == mask of base.py:93
    C[np.diag_indices(L)] = softplus(u_diag) + floor
    return C, C @ C.T


def variance_pack(value, floor=1e-3):
    """gpf/base.py:71-96 - the parameters a Variance(value) is created with."""
    value = np.atleast_2d(np.asarray(value, dtype=np.float64))
    C = np.linalg.cholesky(value)
    d = np.diag(C).copy()
    if d.min() <= floor:
        raise ValueError('Cholesky diagonal must exceed its lower bound')
    return softplus_inverse(d - floor), C[np.tril_indices(len(d), -1)].copy()


# ----------------------------------------------------------------------------------------------------------------
# covariant gram
# ----------------------------------------------------------------------------------------------------------------
def gram_mo_unit(X, X2, ls):
    """K_unit[l,n,l',n'] = exp(-1/2 sum_m (X[n,m]/ls[l,m] - X2[n',m]/ls[l',m])^2), gpf/kernels.py:82,154. Returns (L,N,L,N2)."""
    X, X2, ls = np.asarray(X, float), np.asarray(X if X2 is None else X2, float), np.atleast_2d(np.asarray(ls, float))
    L = ls.shape[0]
    A = X[None, :, :] / ls[:, None, :]       # (L,N,M)
    B = X2[None, :, :] / ls[:, None, :]      # (L,N2,M)
    out = np.empty((L, X.shape[0], L, X2.shape[0]))
    for l in range(L):
        for j in range(L):
            d = A[l][:, None, :] - B[j][None, :, :]
            out[l, :, j, :] = np.exp(-0.5 * np.einsum('abm,abm->ab', d, d))
    return out


def gram_mo(X, X2, ls, F):
    """gpf/kernels.py:94-116: (LN, LN2) = reshape(F[l,1,l',1] * K_unit)."""
    K = gram_mo_unit(X, X2, ls)
    F = np.asarray(F, float)
    L, N, _, N2 = K.shape
    return (F[:, None, :, None] * K).reshape(L * N, L * N2)


def add_noise_mo(K, E):
    """gpf/likelihoods.py:64-67 + gpf/base.py:62-69: K + E[l,l'] * delta(n,n')."""
    E = np.asarray(E, float)
    L = E.shape[0]
    N = K.shape[0] // L
    return K + np.kron(E, np.eye(N))


def diag_noise(noise_variance, L):
    """gpf/models.py:131-133 (quirk Q1): noise is always broadcast to (L,L) and stripped to its diagonal."""
    return np.diag(np.diag(np.broadcast_to(np.asarray(noise_variance, float), (L, L))))


def lml_from_K(K, y):
    """gpflow.logdensities.multivariate_normal with mu = 0, as used at gpf/models.py:81-82."""
    Lc = np.linalg.cholesky(K)
    alpha = sla.solve_triangular(Lc, y, lower=True)
    return float(-0.5 * alpha @ alpha - 0.5 * len(y) * LOG2PI - np.log(np.diag(Lc)).sum()), Lc, alpha


def lml_mo(X, Y, ls, F, E):
    X, Y = np.asarray(X, float), np.asarray(Y, float)
    y = Y.T.reshape(-1)
    K = add_noise_mo(gram_mo(X, None, ls, F), E)
    return lml_from_K(K, y)[0]


def lml_grad_mo(X, Y, ls, F, E):
    """LML and its gradient w.r.t. the entries of F, E (treated as independent) and the lengthscales.

    Returns dict(lml, dF (L,L), dE (L,L), dls (L,M)).  W = a a^T - K^-1, dLML/dtheta = 1/2 tr(W dK/dtheta).
    """
    X, Y, ls, F, E = (np.asarray(a, float) for a in (X, Y, np.atleast_2d(ls), F, E))
    N, M = X.shape
    L = Y.shape[1]
    y = Y.T.reshape(-1)
    U = gram_mo_unit(X, None, ls)
    K = add_noise_mo((F[:, None, :, None] * U).reshape(L * N, L * N), E)
    lml, Lc, _ = lml_from_K(K, y)
    Kinv = sla.cho_solve((Lc, True), np.eye(L * N))
    a = Kinv @ y
    W = (np.outer(a, a) - Kinv).reshape(L, N, L, N)
    dF = 0.5 * np.einsum('anbm,anbm->ab', W, U)
    dE = 0.5 * np.einsum('anbn->ab', W)
    WFU = W * F[:, None, :, None] * U
    Xs = X[None, :, :] / ls[:, None, :]                       # (L,N,M)
    dls = np.zeros((L, M))
    for l in range(L):
        for j in range(L):
            d = Xs[l][:, None, :] - Xs[j][None, :, :]          # (N,N,M)
            dls[l] += np.einsum('ab,abm,am->m', WFU[l, :, j, :], d, X / ls[l] ** 2)
    return {'lml': lml, 'dF': dF, 'dE': dE, 'dls': dls}


def chain_variance(dV, u_diag, lower, floor=1e-3):
    """Push d/dV (entries independent) through V = C C^T and the softplus(+floor) diagonal (gpf/base.py:42-55,88-94)."""
    C, _ = variance_unpack(u_diag, lower, floor)
    dC = np.tril((dV + dV.T) @ C)
    return np.diag(dC) * sigmoid(u_diag), dC[np.tril_indices(len(u_diag), -1)]


def predict_mo(X, Y, ls, F, E, Xs, y_instead_of_f=True):
    """gpf/models.py:84-111 + gpflow base_conditional + gpf/likelihoods.py:80-89. Returns (mean (n,L), var (n,L))."""
    X, Y, ls, F, E, Xs = (np.asarray(a, float) for a in (X, Y, np.atleast_2d(ls), F, E, Xs))
    L = Y.shape[1]
    n = Xs.shape[0]
    y = Y.T.reshape(-1)
    Kmm = add_noise_mo(gram_mo(X, None, ls, F), E)
    Kmn = gram_mo(X, Xs, ls, F)
    Lm = np.linalg.cholesky(Kmm)
    A = sla.solve_triangular(Lm, Kmn, lower=True)
    knn_diag = np.repeat(np.diag(F), n)                       # diag of kernel(Xs, Xs): F[l,l] * exp(0)
    fvar = knn_diag - np.einsum('ki,ki->i', A, A)
    A = sla.solve_triangular(Lm.T, A, lower=False)
    fmean = A.T @ y
    mean, var = fmean.reshape(L, n).T, fvar.reshape(L, n).T
    if y_instead_of_f:
        var = var + np.diag(E)[None, :]
    return mean, var


def k_cho_mo(X, ls, F, E):
    """gpr/models.py:427-431,439 (covariant): chol(add_to(KXX)), shape (LN,LN)."""
    return np.linalg.cholesky(add_noise_mo(gram_mo(X, None, ls, F), E))


def k_inv_y_mo(X, Y, ls, F, E):
    """gpr/models.py:441-444 (covariant): cholesky_solve(K_cho, vec(Y^T)) reshaped (L,1,N)."""
    Y = np.asarray(Y, float)
    Lc = k_cho_mo(X, ls, F, E)
    return sla.cho_solve((Lc, True), Y.T.reshape(-1)).reshape(Y.shape[1], 1, Y.shape[0])


# ----------------------------------------------------------------------------------------------------------------
# variant path: gpflow GPR + kernels.RBF (SquaredExponential) + likelihoods.Gaussian
# ----------------------------------------------------------------------------------------------------------------
def gram_rbf(X, X2, ls, variance):
    """gpflow SquaredExponential: variance * exp(-1/2 square_distance(X/ls, X2/ls)), with
    square_distance = -2 X X2^T + |X|^2 + |X2|^2 (gpflow.utilities.ops.square_distance)."""
    X = np.asarray(X, float) / np.asarray(ls, float)
    if X2 is None:
        Xs = np.sum(X * X, axis=-1, keepdims=True)
        r2 = -2.0 * (X @ X.T) + Xs + Xs.T
    else:
        X2 = np.asarray(X2, float) / np.asarray(ls, float)
        r2 = -2.0 * (X @ X2.T) + np.sum(X * X, -1)[:, None] + np.sum(X2 * X2, -1)[None, :]
    return variance * np.exp(-0.5 * r2)


def lml_rbf(X, y, ls, variance, noise):
    """gpflow.models.GPR.log_marginal_likelihood for one output column y (N,)."""
    K = gram_rbf(X, None, ls, variance)
    K[np.diag_indices_from(K)] += noise
    return lml_from_K(K, np.asarray(y, float).reshape(-1))[0]


def lml_grad_rbf(X, y, ls, variance, noise, gemm_form=False):
    """LML and d/d(ls[m]), d/d(variance), d/d(noise) for one gpflow GPR (variant).
    gemm_form: sum_ab WK[a,b] (x_a - x_b)^2 = 2 sum_a x_a (rowsum(WK)[a] x_a - (WK X)[a]) for symmetric WK - the same sum without the
    (N,N,M) tensor, for N in the thousands (the cfg2 folds)."""
    X, y, ls = np.asarray(X, float), np.asarray(y, float).reshape(-1), np.broadcast_to(np.asarray(ls, float), (np.shape(X)[1],))
    Kf = gram_rbf(X, None, ls, variance)
    K = Kf.copy()
    K[np.diag_indices_from(K)] += noise
    lml, Lc, _ = lml_from_K(K, y)
    Kinv = sla.cho_solve((Lc, True), np.eye(len(y)))
    a = Kinv @ y
    W = np.outer(a, a) - Kinv
    WK = W * Kf
    if gemm_form:
        s = 2.0 * np.einsum('am,am->m', X, WK.sum(axis=1)[:, None] * X - WK @ X)
    else:
        s = np.einsum('ab,abm->m', WK, (X[:, None, :] - X[None, :, :]) ** 2)
    return {'lml': lml, 'dls': 0.5 * s / ls ** 3, 'dvariance': 0.5 * WK.sum() / variance, 'dnoise': 0.5 * np.trace(W)}


def predict_rbf(X, y, ls, variance, noise, Xs, y_instead_of_f=True):
    """gpflow.models.GPR.predict_f / predict_y (full_cov=False) for one output. Returns (mean (n,), var (n,))."""
    X, y, Xs = np.asarray(X, float), np.asarray(y, float).reshape(-1), np.asarray(Xs, float)
    Kmm = gram_rbf(X, None, ls, variance)
    Kmm[np.diag_indices_from(Kmm)] += noise
    Kmn = gram_rbf(X, Xs, ls, variance)
    Lm = np.linalg.cholesky(Kmm)
    A = sla.solve_triangular(Lm, Kmn, lower=True)
    fvar = variance - np.einsum('ki,ki->i', A, A)
    A = sla.solve_triangular(Lm.T, A, lower=False)
    mean = A.T @ y
    return mean, fvar + (noise if y_instead_of_f else 0.0)


def k_cho_rbf(X, ls, variance, noise):
    """gpr/models.py:432-439 (variant): stacked (L,N,N) Cholesky factors. ls (L,M), variance (L,), noise (L,)."""
    out = []
    for l in range(len(variance)):
        K = gram_rbf(X, None, ls[l], variance[l])
        K[np.diag_indices_from(K)] += noise[l]
        out.append(np.linalg.cholesky(K))
    return np.stack(out)


def k_inv_y_rbf(X, Y, ls, variance, noise):
    """gpr/models.py:441-444 (variant): (L,1,N)."""
    Y = np.asarray(Y, float)
    cho = k_cho_rbf(X, ls, variance, noise)
    return np.stack([sla.cho_solve((cho[l], True), Y[:, l]) for l in range(Y.shape[1])])[:, None, :]


def predict_gradient_rbf(X, Y, ls, variance, noise, xs):
    """gpr/models.py:386-415, variant branch (the covariant branch of the reference raises a TypeError at :405).
    mean (o,L,M) = d/dx of the predictive mean; var (O,o,L,M,m) = -W^T W with W = K_cho^-1 dK(X,x)/dx, plus, on the M == m diagonal,
    Lambda[L,M]^2 k_L(x_O, x_o) with Lambda = 1/lengthscale - exactly the reference's expression (it drops the (x-x')^2 term of the second
    derivative and fills only the diagonal in M)."""
    X, Y, xs = np.asarray(X, float), np.asarray(Y, float), np.asarray(xs, float)
    L, M, o = Y.shape[1], X.shape[1], xs.shape[0]
    cho = k_cho_rbf(X, ls, variance, noise)
    KiY = k_inv_y_rbf(X, Y, ls, variance, noise)
    mean = np.zeros((o, L, M))
    var = np.zeros((o, o, L, M, M))
    for l in range(L):
        K = gram_rbf(X, xs, ls[l], variance[l])                                            # (N,o)
        J = K[:, :, None] * (X[:, None, :] - xs[None, :, :]) / (ls[l] ** 2)[None, None, :]   # (N,o,M)
        mean[:, l, :] = np.einsum('NoM,N->oM', J, KiY[l, 0])
        W = sla.solve_triangular(cho[l], J.reshape(X.shape[0], o * M), lower=True).reshape(X.shape[0], o, M)
        var[:, :, l] = -np.einsum('NOM,Nom->OoMm', W, W)
        kxx = gram_rbf(xs, xs, ls[l], variance[l])
        lam = 1.0 / ls[l]
        dd = lam[None, None, :] * lam[None, None, :] * kxx[:, :, None]                     # (O,o,M)
        idx = np.arange(M)
        var[:, :, l, idx, idx] += dd
    return mean, var


# ----------------------------------------------------------------------------------------------------------------
# the timed CPU unit: one LML + gradient evaluation through LAPACK (potrf + potri), blocked gram
# ----------------------------------------------------------------------------------------------------------------
def dls_gemm_form(X, ls, F, W4, U4):
    """d LML / d ls[l,m] of lml_grad_mo in GEMM form (no (N,N,M) tensor), for sizes where the broadcast form does not fit:
         G_lj = W_lj * F[l,j] * U_lj,   dls[l,m] = sum_j sum_a X[a,m] ( rowsum(G_lj)[a] X[a,m]/ls[l,m] - (G_lj Xs_j)[a,m] ) / ls[l,m]^2
    with Xs_j = X / ls[j] - the same sum as 'ab,abm,am->m' over d = Xs_l[a] - Xs_j[b] (gpf/kernels.py:82 differentiated)."""
    L, N = W4.shape[0], W4.shape[1]
    Xs = X[None, :, :] / ls[:, None, :]
    dls = np.zeros_like(ls)
    for l in range(L):
        for j in range(L):
            if F[l, j] == 0.0:
                continue
            G = W4[l, :, j, :] * (F[l, j] * U4[l, :, j, :])
            dls[l] += np.einsum('am,am->m', X, G.sum(axis=1)[:, None] * Xs[l] - G @ Xs[j]) / ls[l] ** 2
    return dls


def lml_grad_mo_lapack(X, Y, ls, F, E, with_lengthscales=False, predict_at=None, want_kinvy=False):
    """Same numbers as lml_grad_mo (and predict_mo / k_inv_y_mo), organised the way a tuned CPU code would run it: ONE dpotrf shared by
    the LML, the predictions and K^-1 y, dpotri for the explicit inverse (OpenBLAS), gram built block by block.  bench.py's cpu_baseline
    (default arguments) and the full-size parity tests (tests/test_gpu_baseline_sizes.py).  Memory: ~4 n^2 doubles.

    predict_at: optional (n*, M) inputs -> out['mean'], out['var_f'] (n*, L) as predict_mo(y_instead_of_f=False).
    want_kinvy: out['KiY'] (L,1,N) as k_inv_y_mo."""
    X, Y, ls, F, E = (np.asarray(a, float) for a in (X, Y, np.atleast_2d(ls), F, E))
    N, M = X.shape
    L = Y.shape[1]
    n = L * N
    y = Y.T.reshape(-1)
    U = np.empty((n, n))
    Xs = X[None, :, :] / ls[:, None, :]
    sq = np.einsum('lnm,lnm->ln', Xs, Xs)
    for l in range(L):
        for j in range(L):
            r2 = sq[l][:, None] + sq[j][None, :] - 2.0 * (Xs[l] @ Xs[j].T)  # blocked BLAS form; fine for a timing baseline
            np.exp(-0.5 * r2, out=U[l * N:(l + 1) * N, j * N:(j + 1) * N])
    K = U * np.kron(F, np.ones((N, N)))
    K += np.kron(E, np.eye(N))
    c, info = sla.lapack.dpotrf(K, lower=1, overwrite_a=1)
    if info != 0:
        raise np.linalg.LinAlgError(f'dpotrf info={info}')
    alpha = sla.solve_triangular(c, y, lower=True)
    lml = float(-0.5 * alpha @ alpha - 0.5 * n * LOG2PI - np.log(np.diag(c)).sum())
    a = sla.solve_triangular(c, alpha, lower=True, trans='T')
    out = {'lml': lml}
    if want_kinvy:
        out['KiY'] = a.reshape(L, 1, N).copy()
    if predict_at is not None:
        Xn = np.asarray(predict_at, float)
        ns = Xn.shape[0]
        Kmn = gram_mo(X, Xn, ls, F)
        A = sla.solve_triangular(c, Kmn, lower=True)
        out['var_f'] = (np.repeat(np.diag(F), ns) - np.einsum('ki,ki->i', A, A)).reshape(L, ns).T
        out['mean'] = (Kmn.T @ a).reshape(L, ns).T
    Kinv, info = sla.lapack.dpotri(c, lower=1, overwrite_c=1)
    Kinv = np.tril(Kinv) + np.tril(Kinv, -1).T
    W = np.outer(a, a) - Kinv
    W4, U4 = W.reshape(L, N, L, N), U.reshape(L, N, L, N)
    out |= {'dF': 0.5 * np.einsum('anbm,anbm->ab', W4, U4), 'dE': 0.5 * np.einsum('anbn->ab', W4)}
    if with_lengthscales:
        out['dls'] = dls_gemm_form(X, ls, F, W4, U4)
    return out


# ----------------------------------------------------------------------------------------------------------------
# variant fit: what gf.optimizers.Scipy().minimize(gp.training_loss, gp.trainable_variables, options=...) does per output
# (gpr/models.py:359-361) - L-BFGS-B on the unconstrained variables in tf.Module order (kernel.lengthscales, kernel.variance,
# likelihood.variance), softplus transforms, noise floor 1e-6 (gpflow.likelihoods.Gaussian)
# ----------------------------------------------------------------------------------------------------------------
def fit_rbf(X, y, ls0, variance0, noise0, train_lengthscales=True, noise_floor=1e-6, **options):
    """-> dict(ls (M,), variance, noise, lml, nit, nfev).  ls0 of size 1 = isotropic kernel (one shared lengthscale)."""
    import scipy.optimize
    X, y = np.asarray(X, float), np.asarray(y, float).reshape(-1)
    M = X.shape[1]
    ls0 = np.atleast_1d(np.asarray(ls0, float))
    k = ls0.size

    def unpack(u):
        ls = softplus(u[:k]) if train_lengthscales else ls0
        off = k if train_lengthscales else 0
        return ls, float(softplus(u[off])), float(softplus(u[off + 1]) + noise_floor)

    def fun(u):
        ls, var, noise = unpack(u)
        r = lml_grad_rbf(X, y, np.broadcast_to(ls, (M,)), var, noise, gemm_form=X.shape[0] > 600)
        g = []
        if train_lengthscales:
            dls = r['dls'] if k == M else np.array([r['dls'].sum()])
            g.append(dls * sigmoid(u[:k]))
        off = k if train_lengthscales else 0
        g.append([r['dvariance'] * sigmoid(u[off])])
        g.append([r['dnoise'] * sigmoid(u[off + 1])])
        return -r['lml'], -np.concatenate([np.reshape(v, -1) for v in g])

    u0 = np.concatenate(([softplus_inverse(ls0)] if train_lengthscales else []) + [[softplus_inverse(variance0)], [softplus_inverse(noise0 - noise_floor)]])
    res = scipy.optimize.minimize(fun, np.asarray(u0, float).reshape(-1), jac=True, method='L-BFGS-B', options=options)
    ls, var, noise = unpack(res.x)
    return {'ls': np.broadcast_to(ls, (M,)).copy() if k == M else ls, 'variance': var, 'noise': noise, 'lml': -float(res.fun), 'nit': res.nit, 'nfev': res.nfev}
