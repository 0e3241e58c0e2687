"""Oracle (float64 numpy) for the standard errors of the closed Sobol indices.  TEST INFRASTRUCTURE ONLY.

Restates ``ClosedSobolWithError`` (romcomma/gsa/calibrators.py:146-402) for a diagonal kernel variance F (``:380-381`` raises otherwise),
first for ``is_T_partial=True`` (``META``, ``:149-157``), then for the non-partial form the reference's scripts run.
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

(Pi of ``:227-228`` simplifies to 1 - phi and Omega of ``:234-235`` to phi_i phi_j.)

``is_T_partial=False`` (the default of all three reference scripts: installation_test.py:51, csv_script.py:48, benchmark_script.py:144) adds the
MIXED rank equation ``(l='k', i='k', j='j', k='i')`` (``:169-170``): the covariance W[Mm] between the FULL model and the marginal s.  In the
rank-8 tensors it equates the first output with i and leaves j = l' free, so with the general (l, i, j) form of the Omega Gaussian
(a = phi_i phi_j phi_l r_li, v = g_j phi_j + phi_j^2 g_i + phi_i^2 phi_j^2 r_li g_l, r_li = (1-ups_i)/(1-phi_l ups_i)) evaluated at l = i, j = l':

  phi   a' = phi_i^2 phi_l r_ii,  v' = g_l phi_l + phi_l^2 g_i + phi_i^2 phi_l^2 r_ii g_i,  b' = ups_i phi_i^2/(1-ups_i phi_i)
        Q'_il[N,n] = prod_{m in s} (v'/phi_l)^-1/2 exp( sum_{m in s} -1/2 (a' x - phi_l y)^2/v' + 1/2 phi_l y^2 )
                     * prod_{ALL m} (1-ups_i phi_i)^-1/2 exp( sum_{ALL m} -1/2 b' x^2 )
        (``marginalize`` pairs the marginal Omega Gaussian with the FULL-model Upsilon Gaussian ``self.UpsilonGaussians.MIXED``, ``:369``, which is
        why the Upsilon factor runs over all M inputs and becomes a per-sample weight of c_i)
        mu_phi_mu_MIXED[l,i] = pre[i] * sum_{N,n} c_i[N] Q'_il[N,n] c_l[n] * (2 if l == i)                                (``:278-281``)
  psi   mu_psi_mu_MIXED[l,i] = psi^FULL_ii . psi^s_li * (2 if l == i)                                       (``:318-322``, first factor ``self.psi_factor``)
  W[Mm] = (mu_phi_mu_MIXED - mu_psi_mu_MIXED) + transpose
  Q     = q_l + q_i + 2 delta_li q_l,  q = diag(W[MM]_MIXED) / (4 V[1]^2)                                                 (``:400-401``)
  T     = sqrt(| W[mm] - 2 V_m W[Mm] / V[1] + V_m^2 Q | / V[2]^2)    (``:343-346``; V[1] is (L,) and broadcasts along the LAST axis, so T is
          not symmetric - replicated as is)
  full model: self.W = (DIAGONAL, MIXED) at s = all inputs, self.T = T(W.DIAGONAL, W.MIXED, V[0])                          (``:395-402``)

Pinned by tests/golden/ref_*.npz, which hold W and T for both settings produced by executing the reference's own file.
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


def mixed_coefficients(phi_i, phi_l, ups_i):
    """Marginal part of the MIXED Omega kernel (x side = output i at l = i, y side = output l); the Upsilon factor is ``mixed_weight``."""
    g_i, g_l = 1.0 - phi_i, 1.0 - phi_l
    r = (1.0 - ups_i) / (1.0 - phi_i * ups_i)
    a = phi_i * phi_i * phi_l * r
    v = g_l * phi_l + phi_l * phi_l * g_i + phi_i * phi_i * phi_l * phi_l * r * g_i
    return a * a / v, phi_l * phi_l / v - phi_l, a * phi_l / v, -0.5 * np.log(v / phi_l)


def mixed_weight(X, phi_i, ups_i):
    """prod_{ALL m} (1 - ups_i phi_i)^-1/2 exp(-1/2 b' x^2): the full-model Upsilon Gaussian as a weight per sample (N,)."""
    b = ups_i * phi_i * phi_i / (1.0 - ups_i * phi_i)
    return np.exp(-0.5 * (X * X) @ b - 0.5 * np.sum(np.log(1.0 - ups_i * phi_i)))


class ClosedSobolWithError(ClosedSobol):
    """K_cho: (LN, LN) lower factor of a covariant GP or (L, N, N) of a variant one (gpr/models.py:427-439)."""

    def __init__(self, X, Lambda, F, K_inv_Y, K_cho, block=1024, is_T_partial=True):
        super().__init__(X, Lambda, F, K_inv_Y, True, block)
        self.K_cho = np.asarray(K_cho, float)
        self.is_T_partial = is_T_partial
        self.Upsilon = self.Lambda2[-1][2]                                         # (L,1,M)
        self.V[4] = self.V[2] * self.V[2]
        self.pre_factor = (np.sqrt(np.prod(self.Lambda2[1][0] * self.Lambda2[-1][2], axis=-1)) * self.F).reshape(-1)
        if is_T_partial:
            self.W = self._W(0, self.M)
        else:
            self.psi_full = [self._psi(i, i, 0, self.M)[1] for i in range(self.L)]
            self.W_DIAGONAL = self._W(0, self.M)
            self.W_MIXED, mixed_scale = self._W_mixed(0, self.M, True)
            self.W = (self.W_DIAGONAL, self.W_MIXED)
            q = np.diag(self.W_MIXED) / (4.0 * self.V[1] * self.V[1])
            self.Q = q[None, :] + q[:, None] + 2.0 * np.diag(q)
            qs = np.diag(mixed_scale) / (4.0 * self.V[1] * self.V[1])
            self.Q_scale = qs[None, :] + qs[:, None] + 2.0 * np.diag(qs)
            self.T = self._T(self.W_DIAGONAL, self.W_MIXED, self.V[0])

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

    def _W(self, m0, m1, with_scale=False):
        """W[mm]; with_scale also |mu_phi_mu| + |mu_psi_mu| (+ transpose): W is the difference of two nearly equal terms, and the rounding
        error of any float64 evaluation scales with THEM (times eps cond(K), through K^-1 y), not with |W| - tests on fitted (ill-conditioned)
        models state their tolerance against this scale."""
        L = self.L
        W, S = np.zeros((L, L)), np.zeros((L, L))
        for l in range(L):
            for i in range(L):
                _, psi = self._psi(l, i, m0, m1)
                phi, twice = self._phi(l, i, m0, m1), (2.0 if l == i else 1.0)
                W[l, i] = (phi - psi @ psi) * twice
                S[l, i] = (abs(phi) + psi @ psi) * twice
        return (W + W.T, S + S.T) if with_scale else W + W.T

    def _phi_mixed(self, l, i, m0, m1):
        s = slice(m0, m1)
        A, B, C, logk = mixed_coefficients(self.Phi[i, 0, s], self.Phi[l, 0, s], self.Upsilon[i, 0, s])
        left = self.g0KY[i, 0] * mixed_weight(self.X, self.Phi[i, 0], self.Upsilon[i, 0])
        Xs = self.X[:, s]
        acc = 0.0
        for r0 in range(0, self.N, self.block):
            rows = slice(r0, min(self.N, r0 + self.block))
            acc += left[rows] @ (pair_kernel(Xs, A, B, C, logk, rows) @ self.g0KY[l, 0])
        return self.pre_factor[i] * acc

    def _W_mixed(self, m0, m1, with_scale=False):
        L = self.L
        W, S = np.zeros((L, L)), np.zeros((L, L))
        for l in range(L):
            for i in range(L):
                _, psi = self._psi(l, i, m0, m1)
                phi, dot, twice = self._phi_mixed(l, i, m0, m1), self.psi_full[i] @ psi, (2.0 if l == i else 1.0)
                W[l, i] = (phi - dot) * twice
                S[l, i] = (abs(phi) + abs(dot)) * twice
        return (W + W.T, S + S.T) if with_scale else W + W.T

    def _T(self, Wmm, WMm=None, Vm=None):
        """calibrators.py:335-346."""
        Q = Wmm if self.is_T_partial else Wmm - 2.0 * Vm * WMm / self.V[1] + Vm * Vm * self.Q
        return np.sqrt(np.abs(Q) / self.V[4])

    def marginalize(self, m: Tuple[int, int]) -> Dict[str, np.ndarray]:
        result = super().marginalize(m)
        W, W_scale = self._W(int(m[0]), int(m[1]), True)
        if self.is_T_partial:
            return result | {'W': W, 'T': self._T(W), 'W_scale': W_scale}
        WMm, WMm_scale = self._W_mixed(int(m[0]), int(m[1]), True)
        # Q_m = W[mm] - 2 V W[Mm]/V1 + V^2 Q, and the size of what cancels in it (for tolerances)
        V = result['V']
        Q_scale = W_scale + 2.0 * np.abs(V) * WMm_scale / self.V[1] + V * V * self.Q_scale
        return result | {'W': W, 'WMm': WMm, 'T': self._T(W, WMm, V), 'W_scale': W_scale, 'WMm_scale': WMm_scale, 'Q_scale': Q_scale}


def sobol_kind_with_error(cal: ClosedSobolWithError, kind: int, m: int = -1) -> Dict[str, np.ndarray]:
    """gsa/models.py:117-137,207-214: V, S get the full-model column appended; with is_T_partial W and T do not, without it T is
    post-processed like S (TOTAL: T_full + T, sic) and gets the full-model column (``:211-213``)."""
    res = [cal.marginalize(s) for s in m_slices(kind, cal.M, m)]
    out = {key: np.stack([r[key] for r in res], axis=-1) for key in ('V', 'S', 'W', 'T')}
    out['V'] = np.concatenate([out['V'], cal.V[0][..., None]], axis=-1)
    if kind == TOTAL:
        out['S'] = cal.S[..., None] - out['S']
    out['S'] = np.concatenate([out['S'], cal.S[..., None]], axis=-1)
    if not cal.is_T_partial:
        if kind == TOTAL:
            out['T'] = cal.T[..., None] + out['T']
        out['T'] = np.concatenate([out['T'], cal.T[..., None]], axis=-1)
    return out
