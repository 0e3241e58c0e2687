"""Oracle (float64 numpy) for the closed-form Sobol contractions.  TEST INFRASTRUCTURE ONLY.

Restates romcomma/gsa/calibrators.py:49-143 (ClosedSobol) and romcomma/gsa/models.py:77-90,117-137,207-214 (slice lists,
sweep, post-processing).  ``_V`` is evaluated in the algebraically simplified form of SURVEY App. A.4 - obtained by
expanding the reference's Gaussian-ratio chain Psi / PsiPhi / PhiG / PhiGauss / H (calibrators.py:69-79,
gsa/base.py:92-126) - and row-blocked so the (l,L',N,j,J',n,m) tensor the reference materialises never exists.
``oracle/literal.py`` holds the un-simplified transliteration used to validate this form.

With p = Phi[l,L',m], q = Phi[j,J',m], psi = 1 - p q, gamma = p q / psi, x = X[N,m], y = X[n,m]:

    H[(l L' N),(j J' n)] = prod_m psi^-1/2 * exp( sum_m ( -1/2 gamma p x^2 - 1/2 gamma q y^2 + gamma x y ) )
    V[l,j]             = sum_{L',J'} g0KY[l,L',:]^T H g0KY[j,J',:]
"""
from __future__ import annotations

import itertools
from typing import Dict, Sequence, Tuple

import numpy as np

FIRST_ORDER, CLOSED, TOTAL = 1, 2, 3      # gsa/models.py:38-42 (IntEnum auto())


def H_block(X, p, q, m0, m1, rows=slice(None)):
    """H[(a,N),(b,n)] for one (a,b) = ((l,L'),(j,J')) pair over the slice [m0:m1]; p, q are the Phi rows (M,)."""
    s = slice(m0, m1)
    p, q = p[s], q[s]
    psi = 1.0 - p * q
    gamma = p * q / psi
    Xs = X[:, s]
    u = -0.5 * (Xs[rows] ** 2) @ (gamma * p)
    v = -0.5 * (Xs ** 2) @ (gamma * q)
    e = (Xs[rows] * gamma) @ Xs.T + u[:, None] + v[None, :]   # summed before exp: the exponent can be positive
    return np.exp(e) / np.sqrt(np.prod(psi))


def V_bilinear(X, Phi, c_left, c_right, m0, m1, block=1024, workers=1):
    """sum_{L',J'} c_left[l,L',:]^T H c_right[j,J',:]  ->  (L,L).  Phi (L,L',M); c (L,L',N).
    workers > 1: the (pair, row block) terms are evaluated by a thread pool (numpy releases the GIL in exp and matmul) and added in
    the same fixed order as the serial loop - the full-size parity tests need seconds, not minutes."""
    L, Lp, _ = Phi.shape
    N = X.shape[0]
    tasks = [(l, a, j, b, r0) for l, a, j, b in itertools.product(range(L), range(Lp), range(L), range(Lp)) for r0 in range(0, N, block)]

    def term(t):
        l, a, j, b, r0 = t
        rows = slice(r0, min(N, r0 + block))
        return c_left[l, a, rows] @ (H_block(X, Phi[l, a], Phi[j, b], m0, m1, rows) @ c_right[j, b])
    if workers > 1:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=workers) as ex:
            terms = list(ex.map(term, tasks))
    else:
        terms = [term(t) for t in tasks]
    V = np.zeros((L, L))
    for (l, a, j, b, r0), v in zip(tasks, terms):
        V[l, j] += v
    return V


class ClosedSobol:
    """gsa/calibrators.py:31-143.  Public attributes as in the reference: V (dict 0,1,2), S, G, Phi, g0, g0KY, Lambda2."""

    def __init__(self, X, Lambda, F, K_inv_Y, is_F_diagonal=True, block=1024, centre=True, workers=1):
        X = np.asarray(X, float)
        self.N, self.M = X.shape
        self.X = X
        F = np.asarray(F, float)
        K_inv_Y = np.asarray(K_inv_Y, float)
        self.L = K_inv_Y.shape[0]
        L, M = self.L, self.M
        self.is_F_diagonal = is_F_diagonal
        self.block, self.workers = block, workers
        # calibrators.py:134-138
        if is_F_diagonal:
            F = np.atleast_2d(F)
            F = (F if F.shape[0] == 1 else np.diag(F)).reshape(L, 1)
        else:
            K_inv_Y = np.transpose(K_inv_Y, (1, 0, 2))
        self.F, self.K_inv_Y = F, K_inv_Y
        self.Lambda = np.broadcast_to(np.asarray(Lambda, float), (L, M))
        # calibrators.py:99-109
        base = (self.Lambda * self.Lambda)[:, None, :] if is_F_diagonal else self.Lambda[:, None, :] * self.Lambda[None, :, :]
        plus = tuple(base + j for j in range(3))
        self.Lambda2 = {1: plus, -1: tuple(1.0 / v for v in plus)}
        # calibrators.py:82-97
        pre_factor = np.sqrt(np.prod(self.Lambda2[1][0] * self.Lambda2[-1][1], axis=-1)) * F       # (L,L')
        expo = -0.5 * np.einsum('nm,lLm->lLn', X * X, self.Lambda2[-1][1])                           # Gaussian(mean=X, variance=L2+1).exponent
        self.g0 = np.exp(expo) * pre_factor[..., None]
        self.g0KY = self.g0 * K_inv_Y
        if centre:
            self.g0KY = self.g0KY - self.g0KY.sum(axis=(1, 2), keepdims=True) / (self.g0KY.shape[1] * self.g0KY.shape[2])
        self.Phi = self.Lambda2[-1][1]
        self.G = np.einsum('lLm,nm->lLnm', self.Phi, X)
        V0 = self._V(0, M)
        self.V = {0: V0, 1: np.diag(V0).copy()}
        r = np.sqrt(self.V[1])
        self.V[2] = np.outer(r, r)
        self.S = self.V[0] / self.V[2]

    def _V(self, m0, m1):
        return V_bilinear(self.X, self.Phi, self.g0KY, self.g0KY, m0, m1, self.block, self.workers)

    def marginalize(self, m: Tuple[int, int]) -> Dict[str, np.ndarray]:
        """calibrators.py:49-58"""
        V = self._V(int(m[0]), int(m[1]))
        return {'V': V, 'S': V / self.V[2]}


def m_slices(kind: int, M: int, m: int = -1):
    """gsa/models.py:77-90 - the slice list of one kind."""
    ms = range(M) if m < 0 else [m]
    if kind == FIRST_ORDER:
        return [(i, i + 1) for i in ms]
    if kind == CLOSED:
        return [(0, i + 1) for i in ms]
    if kind == TOTAL:
        return [(i + 1, M) for i in ms]
    raise ValueError(kind)


def sobol_kind(cal: ClosedSobol, kind: int, m: int = -1) -> Dict[str, np.ndarray]:
    """gsa/models.py:117-137 + 207-214: V, S of shape (L,L,len(slices)+1); last column is the full model."""
    res = [cal.marginalize(s) for s in m_slices(kind, cal.M, m)]
    V = np.stack([r['V'] for r in res], axis=-1)
    S = np.stack([r['S'] for r in res], axis=-1)
    V = np.concatenate([V, cal.V[0][..., None]], axis=-1)
    if kind == TOTAL:
        S = cal.S[..., None] - S
    S = np.concatenate([S, cal.S[..., None]], axis=-1)
    return {'V': V, 'S': S}


def sweep(X, Lambda, F, K_inv_Y, is_F_diagonal=True, kinds: Sequence[int] = (FIRST_ORDER, CLOSED, TOTAL), block=1024):
    """One 'Sobol sweep' with the reference's semantics (user/run.py:152-153): a fresh calibrator per kind (quirk Q5)."""
    return {k: sobol_kind(ClosedSobol(X, Lambda, F, K_inv_Y, is_F_diagonal, block), k) for k in kinds}


def subset_V(X, Lambda, F, K_inv_Y, subset: Sequence[int], is_F_diagonal=True):
    """Closed index of an arbitrary input subset: permute the columns so the subset is a prefix, then marginalize([0,|S|])
    (the reference API itself only takes contiguous slices, gsa/models.py:84-89)."""
    M = np.shape(X)[1]
    rest = [i for i in range(M) if i not in set(subset)]
    perm = list(subset) + rest
    cal = ClosedSobol(np.asarray(X)[:, perm], np.broadcast_to(np.asarray(Lambda, float), (np.shape(K_inv_Y)[0], M))[:, perm],
                      F, K_inv_Y, is_F_diagonal)
    return cal.marginalize((0, len(subset)))
