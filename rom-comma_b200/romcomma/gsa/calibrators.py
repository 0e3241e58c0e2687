""" Closed Sobol indices of a GP by closed-form Gaussian integrals (reference romcomma/gsa/calibrators.py:31-143) on the B200 path.

The precomputation (Phi, g0, mean-centred g0KY) is rc_sobol_prepare; every marginal variance V(m) is the fused
integrand + contraction kernel rc_sobol_contract, which takes a whole list of marginal subsets per launch (``marginalize_many``).
``ClosedSobolWithError`` (the T/W error terms, reference :146-402) adds rc_sobol_error / rc_sobol_error_mixed (diagonal F; is_T_partial or not)."""
from __future__ import annotations

from romcomma.base.definitions import *
from romcomma.gpr.models import GPR
from romcomma.gsa.base import Calibrator
from romcomma import _capi, distributed
from romcomma._tensors import DeviceTensor, HostTensor, as_device


class ClosedSobol(gf.Module, Calibrator):
    """ Calculates closed Sobol Indices."""

    @classproperty
    def META(cls) -> Dict[str, Any]:
        return {}

    def __init__(self, gp: GPR, **kwargs: Any):
        """
        Args:
            gp: The gp to analyze.
            **kwargs: The calculation meta to override META (``is_F_diagonal`` is honoured; unknown keys are kept in ``self.meta``).
        """
        super().__init__()
        self.gp = gp
        self.meta = self.META | kwargs
        self.L, self.M, self.N = self.gp.L, self.gp.M, self.gp.N
        self.Ms = (0, self.M)
        F = np.asarray(self.gp.kernel.data.frames.variance.np, dtype=FLOAT())
        # Is F diagonal?  Defaults to "the GP did not train a kernel covariance" (quirk Q3: True even for covariant GPs by default).
        self.is_F_diagonal = self.meta.pop('is_F_diagonal', None)
        if self.is_F_diagonal is None:
            gp_options = self.gp.read_meta() if self.gp._meta_json.exists() else self.gp.META
            self.is_F_diagonal = not gp_options.pop('kernel', {}).pop('covariance', False)
        self._K_inv_Y = self.gp.K_inv_Y                                   # (L,1,N) device
        if self.is_F_diagonal:
            F = (F if F.shape[0] == 1 else np.diag(F)).reshape(self.L, 1)
            self.K_inv_Y = self._K_inv_Y
        else:
            self.K_inv_Y = DeviceTensor.wrap(self._K_inv_Y.permute(1, 0, 2).contiguous())
        self.F = HostTensor(F)
        self.Lambda = HostTensor(np.broadcast_to(np.asarray(self.gp.kernel.data.frames.lengthscales.np, dtype=FLOAT()), (self.L, self.M)).copy())
        self.Lambda2 = self._Lambda2()
        self._calibrate()

    @property
    def K_cho(self) -> DeviceTensor:
        """ The reference caches gp.K_cho here (calibrators.py:126) although ClosedSobol never reads it; computed on demand instead."""
        return self.gp.K_cho

    def _Lambda2(self) -> Dict[int, Tuple[HostTensor, ...]]:
        """ {1: <Lambda^2 + J>, -1: <Lambda^2 + J>^(-1)} for J in {0,1,2}; shape (L,1,M) if F is diagonal else (L,L,M)."""
        Lam = self.Lambda.numpy()
        base = (Lam * Lam)[:, None, :] if self.is_F_diagonal else Lam[:, None, :] * Lam[None, :, :]
        plus = tuple(HostTensor(base + j) for j in range(3))
        return {1: plus, -1: tuple(HostTensor(1.0 / v) for v in plus)}

    def _calibrate(self):
        """ Everything that does not depend on the marginal subset, then V[0] (full model), V[1], V[2] and S."""
        self._Xd = as_device(self.gp._X)
        Lp = 1 if self.is_F_diagonal else self.L
        KiY = self._K_inv_Y.reshape(self.L, self.N).contiguous()
        Phi, g0, g0KY = _capi.sobol_prepare(self._Xd, _capi.dev(self.Lambda.numpy()), _capi.dev(self.F.numpy().reshape(-1) if self.is_F_diagonal
                                                                                                 else self.F.numpy()), KiY, self.is_F_diagonal)
        self._Phi_d, self._g0KY_d = Phi, g0KY
        self.Phi = DeviceTensor.wrap(Phi.reshape(self.L, Lp, self.M))
        self.g0 = DeviceTensor.wrap(g0.reshape(self.L, Lp, self.N))
        self.g0KY = DeviceTensor.wrap(g0KY.reshape(self.L, Lp, self.N))
        self._parts = None
        V0 = self._V_many([_capi.slice_mask(0, self.M)])[0]
        self.V = {0: HostTensor(V0), 1: HostTensor(np.diag(V0).copy())}
        root = np.sqrt(self.V[1].numpy())
        self.V |= {2: HostTensor(np.outer(root, root))}
        self.S = HostTensor(self.V[0].numpy() / self.V[2].numpy())

    @property
    def G(self) -> DeviceTensor:
        """ G[l,L',n,m] = Phi[l,L',m] * X[n,m] (reference calibrators.py:91); only materialised if asked for."""
        return DeviceTensor.wrap(torch.einsum('lLM,NM->lLNM', self.Phi.as_subclass(torch.Tensor), self._Xd))

    def _V_many(self, masks: Sequence[int]) -> np.ndarray:
        """ V for a list of subsets given as bit masks over the inputs -> host array (len(masks), L, L)."""
        # With several ranks (torchrun) the (N, n) pair space is split by row tile and the partial sums are added with one all-reduce.
        w = distributed.world_size() if self.meta.get('shard_pair_space', False) else 1
        V = _capi.sobol_contract(self._Xd, self._Phi_d, self._g0KY_d, self.L, self.is_F_diagonal, masks, None, distributed.rank() if w > 1 else 0, w)
        if w > 1:
            distributed.all_reduce_sum_tensor(V)
        return V.cpu().numpy()

    def marginalize_many(self, slices: Sequence[Sequence[int]]) -> List[Dict[str, HostTensor]]:
        """ ``[marginalize(m) for m in slices]`` in one kernel launch."""
        V = self._V_many([_capi.slice_mask(int(m[0]), int(m[1])) for m in slices])
        return [{'V': HostTensor(v), 'S': HostTensor(v / self.V[2].numpy())} for v in V]

    def _subset_masks(self, subsets: Sequence[Sequence[int]], total: bool) -> List[int]:
        """ Bit masks of the subsets - of their COMPLEMENTS for total indices: the total index of S is S_full - S_closed(complement of S), the rule
        gsa.models.Sobol applies to its slices (reference gsa/models.py:86-89,207-210: total of [0:m+1] from the closed index of [m+1:M])."""
        full = (1 << self.M) - 1
        masks = [sum(1 << int(i) for i in set(s)) for s in subsets]
        return [full & ~m for m in masks] if total else masks

    def marginalize_subsets(self, subsets: Sequence[Sequence[int]], total: bool = False) -> List[Dict[str, HostTensor]]:
        """ Closed (or, with ``total``, total) indices of arbitrary - not necessarily contiguous - input subsets: the all-subsets sweep of cfg5.
        As in ``Sobol._post_calibrate`` a total result carries the V of the complement it was derived from and S = S_full - S_closed(complement)."""
        V = self._V_many(self._subset_masks(subsets, total))
        V2, S0 = self.V[2].numpy(), self.S.numpy()
        return [{'V': HostTensor(v), 'S': HostTensor(S0 - v / V2 if total else v / V2)} for v in V]

    def marginalize(self, m: TF.Slice) -> Dict[str, HostTensor]:
        """ The closed Sobol index of the input slice [m[0]:m[1]]: {'V': (L,L), 'S': V / V[2]}."""
        return self.marginalize_many([m])[0]


class ClosedSobolWithError(ClosedSobol):
    """ Closed Sobol indices with their standard errors T and the covariances W behind them (reference calibrators.py:146-402).

    As in the reference the kernel variance F must be diagonal (``:380-381``).  ``is_T_partial=True`` (the class ``META``) asserts that the full
    model is variance free: only W[mm] is formed (rc_sobol_error).  ``is_T_partial=False`` - what installation_test.py:51, csv_script.py:48 and
    benchmark_script.py:144 pass - adds the MIXED rank equation (``:169-170``): W[Mm], Q and the non-partial T (``:342-346,358-372,388-402``), one
    more family of pairwise kernels in rc_sobol_error_mixed.  Either way one C-ABI call evaluates, for every requested marginal subset, the
    _psi_factor / Upsilon / Omega Gaussian chains as fused pairwise kernels and the triangular solve with K_cho as a single TRSM; V is a by-product.
    """

    class RankEquations(NamedTuple):
        DIAGONAL: Any
        MIXED: Any

    @classproperty
    def META(cls) -> Dict[str, Any]:
        """ ``is_T_partial`` forces W[Mm] = W[MM] = 0."""
        return {'is_T_partial': True}

    def _calibrate(self):
        super()._calibrate()
        if not self.is_F_diagonal:
            raise NotImplementedError('If the MOGP kernel covariance is not diagonal, the Sobol error calculation is unstable.')
        self._mixed = not self.meta.get('is_T_partial', True)
        self.Upsilon = self.Lambda2[-1][2]
        self.V |= {4: HostTensor(self.V[2].numpy() * self.V[2].numpy())}
        pre = np.sqrt(np.prod(self.Lambda2[1][0].numpy() * self.Lambda2[-1][2].numpy(), axis=-1)) * self.F.numpy()
        self.mu_phi_mu = {'pre-factor': HostTensor(pre.reshape(-1))}
        self._fac = self.gp._factorize()[0]                              # K_cho with its solve workspace, resident for the whole sweep
        self._Lam_d, self._F_d = _capi.dev(self.Lambda.numpy()), _capi.dev(self.F.numpy().reshape(-1))
        self._g0_d = self.g0.as_subclass(torch.Tensor).reshape(self.L, self.N).contiguous()
        self._full = None                                                # (W[MM] DIAGONAL, W[MM] MIXED or None) of the full model

    # -- the full model: evaluated on first use, and for free when a ``marginalize_*`` call comes first (the full-model subset rides along in
    #    that call's single triangular solve instead of paying for a launch-latency-bound solve of its own) -------------------------------
    def _full_model(self) -> Tuple[np.ndarray, Optional[np.ndarray]]:
        if self._full is None:
            self._VW_many([])
        return self._full

    @property
    def W(self):
        """ The covariance W[MM] of the full model: an (L,L) tensor if is_T_partial, else RankEquations(DIAGONAL, MIXED) (reference :388-399)."""
        diagonal, mixed = self._full_model()
        return HostTensor(diagonal) if not self._mixed else self.RankEquations(DIAGONAL=HostTensor(diagonal), MIXED=HostTensor(mixed))

    @property
    def Q(self) -> HostTensor:
        """ q_l + q_i + 2 delta_li q_l with q = diag(W.MIXED) / (4 V[1]^2) (reference :400-401); only defined when not is_T_partial."""
        if not self._mixed:
            raise AttributeError('Q is only calculated when is_T_partial is False.')
        q = np.diag(self._full_model()[1]) / (4.0 * self.V[1].numpy() * self.V[1].numpy())
        return HostTensor(q[None, :] + q[:, None] + 2.0 * np.diag(q))

    @property
    def T(self) -> HostTensor:
        """ The uncertainty of the full model's index (reference :402); only defined when not is_T_partial."""
        if not self._mixed:
            raise AttributeError('T of the full model is only calculated when is_T_partial is False.')
        diagonal, mixed = self._full_model()
        return HostTensor(self._T(diagonal, mixed, self.V[0].numpy()))

    def _T(self, Wmm: np.ndarray, WMm: Optional[np.ndarray] = None, Vm: Optional[np.ndarray] = None) -> np.ndarray:
        """ reference :335-346.  V[1] is an (L,) vector: it divides along the LAST axis, as in the reference (T is not symmetric)."""
        Q = Wmm if not self._mixed else Wmm - 2.0 * Vm * WMm / self.V[1].numpy() + Vm * Vm * self.Q.numpy()
        return np.sqrt(np.abs(Q) / self.V[4].numpy())

    def _VW_many(self, masks: Sequence[int]):
        """ -> (V, W[mm], W[Mm] or None), host arrays (len(masks), L, L)."""
        masks = [int(m) for m in masks]
        ride_along = self._full is None               # rc_sobol_error chunks long lists itself
        if ride_along:
            masks = masks + [_capi.slice_mask(0, self.M)]
        out = _capi.sobol_error(self._Xd, self._Lam_d, self._F_d, self._Phi_d, self._g0_d, self._g0KY_d, self._fac, masks, mixed=self._mixed)
        V, W = out[0].cpu().numpy(), out[1].cpu().numpy()
        WMm = out[2].cpu().numpy() if self._mixed else None
        if ride_along:
            self._full = (W[-1], WMm[-1] if self._mixed else None)
            V, W, WMm = V[:-1], W[:-1], (WMm[:-1] if self._mixed else None)
        return V, W, WMm

    def _results(self, V: np.ndarray, W: np.ndarray, WMm: Optional[np.ndarray]) -> List[Dict[str, HostTensor]]:
        V2 = self.V[2].numpy()
        if not self._mixed:
            return [{'V': HostTensor(v), 'S': HostTensor(v / V2), 'W': HostTensor(w), 'T': HostTensor(self._T(w))} for v, w in zip(V, W)]
        return [{'V': HostTensor(v), 'S': HostTensor(v / V2), 'W': HostTensor(w), 'T': HostTensor(self._T(w, wm, v))} for v, w, wm in zip(V, W, WMm)]

    def marginalize_many(self, slices: Sequence[Sequence[int]]) -> List[Dict[str, HostTensor]]:
        return self._results(*self._VW_many([_capi.slice_mask(int(m[0]), int(m[1])) for m in slices]))

    def marginalize_subsets(self, subsets: Sequence[Sequence[int]], total: bool = False) -> List[Dict[str, HostTensor]]:
        """ With ``total`` the results follow ``Sobol._post_calibrate``: S = S_full - S_closed(complement) and, when not is_T_partial,
        T = T_full + T(complement) (reference gsa/models.py:207-213); V and W are those of the complement."""
        results = self._results(*self._VW_many(self._subset_masks(subsets, total)))
        if total:
            S0 = self.S.numpy()
            for r in results:
                r['S'] = HostTensor(S0 - r['S'].numpy())
                if self._mixed:
                    r['T'] = HostTensor(self.T.numpy() + r['T'].numpy())
        return results
