""" Closed Sobol indices of a GP by closed-form Gaussian integrals (reference romcomma/gsa/calibrators.py:31-143) on the B200 path.

The precomputation (Phi, g0, mean-centred g0KY) is rc_sobol_prepare; every marginal variance V(m) is the fused
integrand + contraction kernel rc_sobol_contract, which takes a whole list of marginal subsets per launch (``marginalize_many``).
``ClosedSobolWithError`` (the T/W error terms, reference :146-402) adds rc_sobol_error (diagonal F, is_T_partial)."""
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

    def marginalize_subsets(self, subsets: Sequence[Sequence[int]]) -> List[Dict[str, HostTensor]]:
        """ Closed indices of arbitrary (not necessarily contiguous) input subsets - the all-subsets sweep of cfg5."""
        V = self._V_many([sum(1 << int(i) for i in set(s)) for s in subsets])
        return [{'V': HostTensor(v), 'S': HostTensor(v / self.V[2].numpy())} for v in V]

    def marginalize(self, m: TF.Slice) -> Dict[str, HostTensor]:
        """ The closed Sobol index of the input slice [m[0]:m[1]]: {'V': (L,L), 'S': V / V[2]}."""
        return self.marginalize_many([m])[0]


class ClosedSobolWithError(ClosedSobol):
    """ Closed Sobol indices with their standard errors T and the covariances W behind them (reference calibrators.py:146-402).

    As in the reference the kernel variance F must be diagonal (``:380-381``).  ``is_T_partial=True`` (the default ``META`` and what all of
    the reference's scripts run) asserts that the full model is variance free; the non-partial variant (MIXED rank equations, ``WMm``, ``Q``)
    is not implemented on the device path.  One rc_sobol_error call evaluates, for every requested marginal subset, the _psi_factor /
    Upsilon / Omega Gaussian chains as fused pairwise kernels and the triangular solve with K_cho as a single TRSM; V comes out as a by-product.
    """

    @classproperty
    def META(cls) -> Dict[str, Any]:
        """ ``is_T_partial`` forces W[Mm] = W[MM] = 0."""
        return {'is_T_partial': True}

    def _calibrate(self):
        super()._calibrate()
        if not self.is_F_diagonal:
            raise NotImplementedError('If the MOGP kernel covariance is not diagonal, the Sobol error calculation is unstable.')
        if not self.meta.get('is_T_partial', True):
            raise NotImplementedError('is_T_partial=False (the MIXED rank equations) is not implemented on the B200 path.')
        self.Upsilon = self.Lambda2[-1][2]
        self.V |= {4: HostTensor(self.V[2].numpy() * self.V[2].numpy())}
        pre = np.sqrt(np.prod(self.Lambda2[1][0].numpy() * self.Lambda2[-1][2].numpy(), axis=-1)) * self.F.numpy()
        self.mu_phi_mu = {'pre-factor': HostTensor(pre.reshape(-1))}
        self._fac = self.gp._factorize()[0]                              # K_cho with its solve workspace, resident for the whole sweep
        self._Lam_d, self._F_d = _capi.dev(self.Lambda.numpy()), _capi.dev(self.F.numpy().reshape(-1))
        self._g0_d = self.g0.as_subclass(torch.Tensor).reshape(self.L, self.N).contiguous()
        self._W_full = None

    @property
    def W(self) -> HostTensor:
        """ The covariance W of the full model (reference calibrators.py: computed eagerly in ``_calibrate``).  Here it is evaluated on first
        use, and for free when a ``marginalize_*`` call comes first: the full-model subset rides along in that call's single triangular
        solve (16 more right-hand sides) instead of paying for a launch-latency-bound solve of its own."""
        if self._W_full is None:
            self._W_full = HostTensor(self._VW_many([_capi.slice_mask(0, self.M)])[1][0])
        return self._W_full

    def _VW_many(self, masks: Sequence[int]) -> Tuple[np.ndarray, np.ndarray]:
        masks = [int(m) for m in masks]
        ride_along = self._W_full is None             # rc_sobol_error chunks long lists itself
        if ride_along:
            masks = masks + [_capi.slice_mask(0, self.M)]
        V, W = _capi.sobol_error(self._Xd, self._Lam_d, self._F_d, self._Phi_d, self._g0_d, self._g0KY_d, self._fac, masks)
        V, W = V.cpu().numpy(), W.cpu().numpy()
        if ride_along:
            self._W_full = HostTensor(W[-1])
            V, W = V[:-1], W[:-1]
        return V, W

    def _results(self, V: np.ndarray, W: np.ndarray) -> List[Dict[str, HostTensor]]:
        V2, V4 = self.V[2].numpy(), self.V[4].numpy()
        return [{'V': HostTensor(v), 'S': HostTensor(v / V2), 'W': HostTensor(w), 'T': HostTensor(np.sqrt(np.abs(w) / V4))} for v, w in zip(V, W)]

    def marginalize_many(self, slices: Sequence[Sequence[int]]) -> List[Dict[str, HostTensor]]:
        return self._results(*self._VW_many([_capi.slice_mask(int(m[0]), int(m[1])) for m in slices]))

    def marginalize_subsets(self, subsets: Sequence[Sequence[int]]) -> List[Dict[str, HostTensor]]:
        return self._results(*self._VW_many([sum(1 << int(i) for i in set(s)) for s in subsets]))
