"""Literal torch-CPU (float64) transliteration of the reference's TensorFlow broadcasting code.  TEST INFRASTRUCTURE ONLY.

Purpose: pin the simplified closed forms of ``oracle/gp.py`` / ``oracle/sobol.py`` against the *un-simplified* computation
the reference performs - same tensors, same axis insertions, same einsum strings, ``torch`` standing in for ``tf`` -
on shapes small enough to materialise the (l,L',N,j,J',n,m) intermediates.  Gradients come from autograd through
``torch.linalg.cholesky``, mirroring ``tf.GradientTape`` at romcomma/gpr/models.py:359-361.
"""
from __future__ import annotations

import math

import torch

T = torch.float64


def _t(a):
    return a if isinstance(a, torch.Tensor) else torch.as_tensor(a, dtype=T)


# ---- gpf ------------------------------------------------------------------------------------------------------
def softplus(u):
    return torch.nn.functional.softplus(u, threshold=1e9)


def variance_value(u_diag, lower, floor=1e-3):
    """gpf/base.py:42-55: ragged strict-lower rows + set_diag(softplus(u)+floor); value = C C^T."""
    L = u_diag.shape[0]
    C = torch.zeros((L, L), dtype=T)
    k = 0
    for i in range(1, L):                      # row lengths 0,1,..,L-1 (base.py:96)
        for j in range(i):
            C = C.index_put((torch.tensor(i), torch.tensor(j)), lower[k])
            k += 1
    C = C + torch.diag(softplus(u_diag) + floor)
    return C @ C.T


def scaled_difference_matrix(X, X2, ls):
    """gpflow AnisotropicStationary.scaled_difference_matrix with ls (L,1,M): (L,N,L,N2,M)."""
    A, B = X / ls, X2 / ls                      # (L,N,M), (L,N2,M)
    a, b = A.reshape(-1, A.shape[-1]), B.reshape(-1, B.shape[-1])
    d = a[:, None, :] - b[None, :, :]
    return d.reshape(tuple(A.shape[:-1]) + tuple(B.shape[:-1]) + (A.shape[-1],))


def K_mo(X, X2, ls, Fv):
    """gpf/kernels.py:82,154,94-104"""
    L = Fv.shape[0]
    d = scaled_difference_matrix(X, X2, ls.reshape(L, 1, -1))
    Ku = torch.exp(-0.5 * torch.einsum('...M,...M->...', d, d))
    sh = Ku.shape
    return (Fv.reshape(L, 1, L, 1) * Ku).reshape(sh[0] * sh[1], sh[2] * sh[3])


def add_to(K, Ev):
    """gpf/likelihoods.py:64-67, gpf/base.py:62-69"""
    L = Ev.shape[0]
    N = K.shape[-1] // L
    noise = Ev.reshape(L, 1, L, 1) * torch.eye(N, dtype=T)[None, :, None, :]
    return K + noise.reshape(K.shape)


def multivariate_normal(x, mu, Lc):
    """gpflow.logdensities.multivariate_normal"""
    alpha = torch.linalg.solve_triangular(Lc, x - mu, upper=False)
    return -0.5 * (alpha ** 2).sum(0) - 0.5 * x.shape[0] * math.log(2 * math.pi) - torch.log(torch.diagonal(Lc)).sum()


def lml_mo(X, Y, ls, Fv, Ev):
    """gpf/models.py:73-82"""
    y = Y.T.reshape(-1, 1)
    Lc = torch.linalg.cholesky(add_to(K_mo(X, X, ls, Fv), Ev))
    return multivariate_normal(y, torch.zeros_like(y), Lc).sum()


def lml_grad_unconstrained_mo(X, Y, u_ls, uF_d, F_low, uE_d, E_low):
    """-> (lml, grads wrt [u_ls, uF_d, F_low, uE_d, E_low]) : what tf.GradientTape hands gpflow's Scipy optimizer."""
    vs = [_t(v).clone().requires_grad_(True) for v in (u_ls, uF_d, F_low, uE_d, E_low)]
    ls = softplus(vs[0])
    val = lml_mo(_t(X), _t(Y), ls, variance_value(vs[1], vs[2]), variance_value(vs[3], vs[4]))
    grads = torch.autograd.grad(val, vs, allow_unused=True)
    return val.item(), [None if g is None else g.numpy() for g in grads]


def base_conditional(Kmn, Kmm, Knn, f):
    """gpflow.conditionals.base_conditional(full_cov=True, white=False)"""
    Lm = torch.linalg.cholesky(Kmm)
    A = torch.linalg.solve_triangular(Lm, Kmn, upper=False)
    fvar = Knn - A.T @ A
    A = torch.linalg.solve_triangular(Lm.T, A, upper=True)
    return A.T @ f, fvar


def predict_mo(X, Y, ls, Fv, Ev, Xn, y_instead_of_f=True):
    """gpf/models.py:84-111 and gpf/likelihoods.py:80-89 (rank-2 branch)"""
    X, Y, ls, Fv, Ev, Xn = (_t(a) for a in (X, Y, ls, Fv, Ev, Xn))
    L, n = Y.shape[1], Xn.shape[0]
    fm, fv = base_conditional(K_mo(X, Xn, ls, Fv), add_to(K_mo(X, X, ls, Fv), Ev), K_mo(Xn, Xn, ls, Fv), Y.T.reshape(-1, 1))
    fm = fm.reshape(L, n)
    fv = torch.einsum('LNLn->LNn', fv.reshape(L, n, L, n))
    fv = torch.einsum('...NN->...N', fv)
    mean, var = fm.T, fv.T
    if y_instead_of_f:
        var = var + torch.diagonal(Ev).reshape(1, L)
    return mean.numpy(), var.numpy()


# ---- gsa ------------------------------------------------------------------------------------------------------
class Gaussian:
    """gsa/base.py:52-126 (diagonal-variance branch only; the reference path never uses the other)."""

    def __init__(self, mean, variance, ordinate=None, LBunch=2):
        ordinate = torch.zeros((), dtype=T) if ordinate is None else ordinate
        cho = torch.sqrt(variance)
        if tuple(ordinate.shape) == tuple(mean.shape):
            shape = list(ordinate.shape)
            fill = [1] * (len(shape) - 1)
            ordinate = ordinate.reshape(shape[:-1] + fill + [shape[-1]])
            mean = mean.reshape(fill + shape)
        ordinate = ordinate - mean
        insertions = cho.dim() - 1
        insertions -= insertions % LBunch
        for axis in range(insertions, 0, -LBunch):
            cho = cho.unsqueeze(axis)
        z = ordinate / torch.broadcast_to(cho, tuple(cho.shape[:-2]) + tuple(ordinate.shape[-2:]))
        self.exponent = -0.5 * torch.einsum('...o,...o->...', z, z)
        self.cho_diag = cho

    def expand_dims(self, axes):
        out = Gaussian.__new__(Gaussian)
        out.exponent, out.cho_diag = self.exponent, self.cho_diag
        for axis in sorted(axes, reverse=True):
            out.exponent = out.exponent.unsqueeze(axis)
            out.cho_diag = out.cho_diag.unsqueeze((axis - 1) if axis < 0 else axis)
        return out

    def __truediv__(self, other):
        out = Gaussian.__new__(Gaussian)
        out.exponent = self.exponent - other.exponent
        out.cho_diag = self.cho_diag / other.cho_diag
        return out

    @property
    def pdf(self):
        return torch.exp(self.exponent) / torch.prod(self.cho_diag, dim=-1)


def closed_sobol_literal(X, Lambda, F, K_inv_Y, is_F_diagonal=True):
    """gsa/calibrators.py:82-143 followed by _V for arbitrary slices. Returns dict(g0, g0KY, G, Phi, V0, V=callable)."""
    X, Lambda, F, K_inv_Y = (_t(a) for a in (X, Lambda, F, K_inv_Y))
    L = K_inv_Y.shape[0]
    M = X.shape[1]
    if is_F_diagonal:
        F = torch.atleast_2d(F)
        F = (F if F.shape[0] == 1 else torch.diagonal(F)).reshape(L, 1)
    else:
        K_inv_Y = K_inv_Y.permute(1, 0, 2)
    Lambda = torch.broadcast_to(Lambda, (L, M))
    base = torch.einsum('lM,lM->lM', Lambda, Lambda)[:, None, :] if is_F_diagonal else torch.einsum('lM,LM->lLM', Lambda, Lambda)
    plus = tuple(base + j for j in range(3))
    Lambda2 = {1: plus, -1: tuple(v ** (-1) for v in plus)}
    pre_factor = torch.sqrt(torch.prod(Lambda2[1][0] * Lambda2[-1][1], dim=-1)) * F
    g0 = torch.exp(Gaussian(mean=X[None, None, ...], variance=Lambda2[1][1]).exponent) * pre_factor[..., None]
    g0KY = g0 * K_inv_Y
    g0KY = g0KY - torch.einsum('lLN->l', g0KY)[..., None, None] / float(g0KY.shape[1] * g0KY.shape[2])
    G = torch.einsum('lLM,NM->lLNM', Lambda2[-1][1], X)
    Phi = Lambda2[-1][1]

    def V(m0, m1, left=g0KY, right=g0KY):
        g, phi = G[..., m0:m1], Phi[..., m0:m1]
        Gamma = 1 - phi
        Psi = Gamma[:, :, None, None, :] + Gamma[None, None, ...]
        Psi = Psi - torch.einsum('lLM,jJM->lLjJM', Gamma, Gamma)
        PsiPhi = torch.einsum('lLjJM,lLM->lLjJM', Psi, phi)
        PhiG = torch.einsum('lLM,jJnM->lLjJnM', phi, g).unsqueeze(2)
        PhiGauss = Gaussian(mean=g, variance=phi)
        H = Gaussian(mean=PhiG, variance=PsiPhi, ordinate=g[..., None, None, None, :])
        H = H / PhiGauss.expand_dims([-1, -2, -3])
        return torch.einsum('lLN,lLNjJn,jJn->lj', left, H.pdf, right).numpy()

    return {'g0': g0.numpy(), 'g0KY': g0KY.numpy(), 'g0KY_uncentred': (g0 * K_inv_Y).numpy(), 'G': G.numpy(), 'Phi': Phi.numpy(),
            'V': V, 'V0': V(0, M)}
