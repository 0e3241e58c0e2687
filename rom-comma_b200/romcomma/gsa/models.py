""" User interface for GSA on a MOGP (reference romcomma/gsa/models.py:35-214): slice lists per kind, the sweep, post-processing
(total = full - closed of the complement) and the S/V csv output under ``<gp>/gsa/<kind>[.<m>]/`` with '%.6f' formatting."""
from __future__ import annotations

from enum import IntEnum, auto

from romcomma.base.definitions import *
from romcomma.base.classes import Model, Data, Frame
from romcomma.gpr.models import GPR
from romcomma.gsa.base import Calibrator
from romcomma.gsa.calibrators import ClosedSobol, ClosedSobolWithError


class GSA(Model):
    """ A generic sensitivity calculation over a list of marginal slices."""

    class Kind(IntEnum):
        FIRST_ORDER = auto()
        CLOSED = auto()
        TOTAL = auto()

    @classproperty
    def ALL_KINDS(cls) -> List['GSA.Kind']:
        return [kind for kind in cls.Kind]

    def __init__(self, gp: GPR, kind: 'GSA.Kind', m: int = -1, is_error_calculated: bool = False, **kwargs: Any):
        """
        Args:
            gp: The underlying Gaussian Process.
            kind: first order, closed or total.
            m: A single input ``0 <= m < gp.M``, or anything else for all of ``range(M)``.
            is_error_calculated: Whether to calculate the standard error T (and the covariances W) of the index.
            **kwargs: The calculation meta to override META.
        """
        self.gp, self.kind, self.is_error_calculated = gp, kind, is_error_calculated
        m = m if 0 <= m < gp.M else -1
        name = kind.name.lower() if m == -1 else f'{kind.name.lower()}.{m}'
        folder = gp.folder / 'gsa' / name
        super().__init__(folder, read_data=False)
        self.meta = {'folder': str(folder), 'm': m, 'M': gp.M} | self.META | kwargs
        self.write_meta(self.meta)

    @staticmethod
    def _columns(M: int, m_cols: int, m_list: List[int]) -> pd.Index:
        if m_cols > len(m_list):
            m_list = m_list + [M]
        if m_cols > len(m_list):
            m_list = [-1] + m_list
        return pd.Index(m_list, name='m')

    @staticmethod
    def _index(shape: List[int]) -> pd.MultiIndex:
        axes = [list(range(l)) for l in shape[:-1]]
        return pd.MultiIndex.from_product(axes, names=[f'l.{l}' for l in range(len(axes))])

    @property
    def _m_dataset(self) -> List[Tuple[int, int]]:
        """ The slices [m0:m1] to marginalize: first order [m,m+1], closed [0,m+1], total [m+1,M]."""
        m, M = self.meta['m'], self.meta['M']
        ms = range(M) if m < 0 else [m]
        if self.kind == GSA.Kind.FIRST_ORDER:
            return [(i, i + 1) for i in ms]
        if self.kind == GSA.Kind.CLOSED:
            return [(0, i + 1) for i in ms]
        if self.kind == GSA.Kind.TOTAL:
            return [(i + 1, M) for i in ms]
        return []

    @property
    @abstractmethod
    def calibrator(self) -> Calibrator:
        raise NotImplementedError('This is a base class.')

    @abstractmethod
    def _post_calibrate(self, calibrator: Calibrator, results: Dict[str, Any]) -> Dict[str, Any]:
        raise NotImplementedError('This is a base class.')

    def _compose_and_save(self, results: Dict[str, Any]):
        m, M = self.meta['m'], self.meta['M']
        m_list = list(range(M)) if m < 0 else [m]
        for key, value in self.data.asdict().items():
            result = results.get(key, None)
            if result is not None:
                result = np.asarray(result)
                shape = list(result.shape)
                df = pd.DataFrame(result.reshape(-1, shape[-1]), columns=GSA._columns(M, shape[-1], m_list), index=GSA._index(shape))
                Frame(value.csv, df, float_format='%.6f')

    def calibrate(self, method: str = None, **kwargs) -> Dict[str, Any]:
        """ Run the sweep: all slices of this kind go to the device in ONE launch (the reference loops marginalize)."""
        calibrator = self.calibrator
        per_slice = calibrator.marginalize_many(self._m_dataset)
        results = {key: np.stack([np.asarray(r[key]) for r in per_slice], axis=-1) for key in per_slice[0].keys()}
        self.results = self._post_calibrate(calibrator, results)
        self._compose_and_save(self.results)
        return self.meta


class Sobol(GSA):
    """ Sobol indices of the three kinds."""

    class Data(Data):
        class NamedTuple(NamedTuple):
            """ S: the Sobol index, T: its standard deviation, V: the conditional variances, W: the covariances behind T."""
            S: Any = np.atleast_2d(None)
            T: Any = np.atleast_2d(None)
            V: Any = np.atleast_2d(None)
            W: Any = np.atleast_2d(None)

    @classproperty
    def META(cls) -> Dict[str, Any]:
        return ClosedSobolWithError.META

    @property
    def calibrator(self) -> ClosedSobol:
        return ClosedSobolWithError(self.gp, **self.meta) if self.is_error_calculated else ClosedSobol(self.gp, **self.meta)

    def _post_calibrate(self, calibrator: ClosedSobol, results: Dict[str, Any]) -> Dict[str, Any]:
        V0, S0 = np.asarray(calibrator.V[0])[..., None], np.asarray(calibrator.S)[..., None]
        results['V'] = np.concatenate([results['V'], V0], axis=-1)
        S = S0 - results['S'] if self.kind == GSA.Kind.TOTAL else results['S']
        results['S'] = np.concatenate([S, S0], axis=-1)
        if 'T' in results and not self.meta['is_T_partial']:     # reference models.py:211-213 (TOTAL: the full model's T is ADDED, sic)
            T0 = np.asarray(calibrator.T)[..., None]
            T = T0 + results['T'] if self.kind == GSA.Kind.TOTAL else results['T']
            results['T'] = np.concatenate([T, T0], axis=-1)
        return results
