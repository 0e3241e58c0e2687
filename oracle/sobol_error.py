"""Oracle (float64 numpy) for the standard errors of the closed Sobol indices.  TEST INFRASTRUCTURE ONLY.

Restates ``ClosedSobolWithError`` (romcomma/gsa/calibrators.py:146-402) for the only configuration the reference supports and
its scripts use: diagonal kernel variance F (``:380-381`` raises otherwise) and ``is_T_partial=True`` (``META``, ``:149-157``).
The reference builds rank-8 broadcast tensors (``liLNjkJM``) and takes their diagonals with ``_equateRanks``; with a diagonal F
the axes L', k, J have size one and the two DIAGONAL rank equations reduce to (j = l, any i) and (l = i = j).  Expanding the
Gaussian-ratio chains of ``_psi_factor`` (``:290-309``), ``_UpsilonGaussian`` (``:244-257``), ``_OmegaGaussian`` (``:214-242``) and
``_mu_phi_mu`` (``:259-288``) gives, for a marginal slice s and outputs l, i (phi = Phi[.,m], ups = Upsilon[.,m] = 1/(Lambda^2+2),
g = 1 - phi, x = X[N,m], y = X[n,m]):

  psi   u_li[n]  = sum_N c_l[N] H_li[N,n]                                   H = the kernel of ClosedSobol._V (oracle/sobol.py), so V[l,i] = c_i . u_li
        psi_li   = K_cho^-1 (g0_i * u_li)                                    (covariant GP: embedded in block i of an LN-vector, ``:304-305``)
        mu_psi_mu[l,i] = |psi_li|^2 * (2 if l == i)                                                                  (``:311-322``)
  phi   a = phi_i phi_l^2 (1-ups_i)/(1-phi_l ups_i),  v = g_l phi_l + phi_l^2 g_i + phi_i^2 phi_l^2 (1-ups_i) g_l/(1-phi_l ups_i),
        b = ups_i phi_l^2/(1-ups_i phi_l)
        Q_li[N,n] = prod_m (v (1-ups_i phi_l)/phi_l)^-1/2 * exp( sum_m -1/2 (a x - phi_l y)^2/v + 1/2 phi_l y^2 - 1/2 b x^2 )
        mu_phi_mu[l,i] = pre[i] * sum_{N,n} c_l[N] Q_li[N,n] c_l[n] * (2 if l == i),   pre[i] = F_i sqrt(prod_{all m} Lambda_i^2/(Lambda_i^2+2))
  W = (mu_phi_mu - mu_psi_mu) + transpose,    T = sqrt(|W| / V[2]^2)                                                   (``:324-346``)

(Pi of ``:227-228`` simplifies to 1 - phi and Omega of ``:234-235`` to phi_i phi_j.)  Pinned by tests/golden/ref_*.npz, which
hold W and T produced by executing the reference's own file.
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np
import scipy.linalg

from .sobol import ClosedSobol, m_slices, FIRST_ORDER, CLOSED, TOTAL  # noqa: F401


def pair_kernel(X, A, B, C, logk, rows=slice(None)):
    """exp( sum_m ( -1/2 A_m x_Nm^2 - 1/2 B_m x_nm^2 + C_m x_Nm x_nm + logk_m ) ) for N in ``rows`` and all n."""
    u = -0.5 * (X[rows] ** 2) @ A
    v = -0.5 * (X ** 2) @ B
    return np.exp((X[rows] * C) @ X.T + u[:, None] + v[None, :] + np.sum(logk))


def psi_coefficients(phi_l, phi_i):
    psi = 1.0 - phi_l * phi_i
    gamma = phi_l * phi_i / psi
    return gamma * phi_l, gamma * phi_i, gamma, -0.5 * np.log(psi)


def omega_coefficients(phi_l, phi_i, ups_i):
    g_l, g_i = 1.0 - phi_l, 1.0 - phi_i
    r = (1.0 - ups_i) / (1.0 - phi_l * ups_i)
    a = phi_i * phi_l * phi_l * r
    v = g_l * phi_l + phi_l * phi_l * g_i + phi_i * phi_i * phi_l * phi_l * r * g_l
    b = ups_i * phi_l * phi_l / (1.0 - ups_i * phi_l)
    return a * a / v + b, phi_l * phi_l / v - phi_l, a * phi_l / v, -0.5 * np.log(v * (1.0 - ups_i * phi_l) / phi_l)


class ClosedSobolWithError(ClosedSobol):
    """K_cho: (LN, LN) lower factor of a covariant GP or (L, N, N) of a variant one (gpr/models.py:427-439)."""

    def __init__(self, X, Lambda, F, K_inv_Y, K_cho, block=1024):
        super().__init__(X, Lambda, F, K_inv_Y, True, block)
        self.K_cho = np.asarray(K_cho, float)
        self.Upsilon = self.Lambda2[-1][2]                                         # (L,1,M)
        self.V[4] = self.V[2] * self.V[2]
        self.pre_factor = (np.sqrt(np.prod(self.Lambda2[1][0] * self.Lambda2[-1][2], axis=-1)) * self.F).reshape(-1)
        self.W = self._W(0, self.M)

    def _psi(self, l, i, m0, m1):
        """(u_li, psi_li)"""
        s = slice(m0, m1)
        A, B, C, logk = psi_coefficients(self.Phi[l, 0, s], self.Phi[i, 0, s])
        Xs = self.X[:, s]
        u = np.zeros(self.N)
        for r0 in range(0, self.N, self.block):
            rows = slice(r0, min(self.N, r0 + self.block))
            u += self.g0KY[l, 0, rows] @ pair_kernel(Xs, A, B, C, logk, rows)
        f = self.g0[i, 0] * u
        if self.K_cho.ndim == 2:
            rhs = np.zeros(self.L * self.N)
            rhs[i * self.N:(i + 1) * self.N] = f
            return u, scipy.linalg.solve_triangular(self.K_cho, rhs, lower=True)
        return u, scipy.linalg.solve_triangular(self.K_cho[i], f, lower=True)

    def _phi(self, l, i, m0, m1):
        s = slice(m0, m1)
        A, B, C, logk = omega_coefficients(self.Phi[l, 0, s], self.Phi[i, 0, s], self.Upsilon[i, 0, s])
        Xs = self.X[:, s]
        acc = 0.0
        for r0 in range(0, self.N, self.block):
            rows = slice(r0, min(self.N, r0 + self.block))
            acc += self.g0KY[l, 0, rows] @ (pair_kernel(Xs, A, B, C, logk, rows) @ self.g0KY[l, 0])
        return self.pre_factor[i] * acc

    def _W(self, m0, m1):
        L = self.L
        W = np.zeros((L, L))
        for l in range(L):
            for i in range(L):
                _, psi = self._psi(l, i, m0, m1)
                W[l, i] = (self._phi(l, i, m0, m1) - psi @ psi) * (2.0 if l == i else 1.0)
        return W + W.T

    def marginalize(self, m: Tuple[int, int]) -> Dict[str, np.ndarray]:
        result = super().marginalize(m)
        W = self._W(int(m[0]), int(m[1]))
        return result | {'W': W, 'T': np.sqrt(np.abs(W) / self.V[4])}


def sobol_kind_with_error(cal: ClosedSobolWithError, kind: int, m: int = -1) -> Dict[str, np.ndarray]:
    """gsa/models.py:117-137,207-214 with is_T_partial: V, S get the full-model column appended, W and T do not."""
    res = [cal.marginalize(s) for s in m_slices(kind, cal.M, m)]
    out = {key: np.stack([r[key] for r in res], axis=-1) for key in ('V', 'S', 'W', 'T')}
    out['V'] = np.concatenate([out['V'], cal.V[0][..., None]], axis=-1)
    if kind == TOTAL:
        out['S'] = cal.S[..., None] - out['S']
    out['S'] = np.concatenate([out['S'], cal.S[..., None]], axis=-1)
    return out
