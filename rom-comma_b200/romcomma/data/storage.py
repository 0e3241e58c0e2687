""" Repository / Fold / Normalization with the reference's on-disk layout (romcomma/data/storage.py:38-558).

<repo>/data.csv (two header rows: level 0 = X | Y heading), meta.json, normalization.csv, fold.<k>/{data.csv,test.csv,meta.json,normalization.csv}.
The K-fold index assignment is integer work driven by Python's global ``random`` state and is bit-exact with the reference for the
same seed (storage.py:180-203); the probit normalisation of X and standardisation of Y follow storage.py:469-485,532-558.
"""
from __future__ import annotations

import itertools
import json
import random
import shutil
from copy import deepcopy
from enum import IntEnum, auto

import scipy.stats

from romcomma.base.definitions import *


class Frame:
    """ A pd.DataFrame backed by a csv with a two-level column header."""

    @classproperty
    def CSV_OPTIONS(cls) -> Dict[str, Any]:
        return {'sep': ',', 'header': [0, 1], 'index_col': 0, }

    def __init__(self, csv: Path | str = Path(), df: pd.DataFrame = pd.DataFrame(), **kwargs):
        self._csv = Path(csv)
        if self.is_empty:
            assert df.empty, 'csv is an empty path, but df is not an empty pd.DataFrame.'
            self.df = df
        elif df.empty:
            self.df = pd.read_csv(self._csv, **{**Frame.CSV_OPTIONS, **kwargs})
        else:
            self.df = df
            self.write()

    @property
    def csv(self) -> Path:
        return self._csv

    @property
    def is_empty(self) -> bool:
        return 0 == len(self._csv.parts)

    def write(self):
        assert not self.is_empty, 'Cannot write when frame.is_empty.'
        self.df.to_csv(path_or_buf=self._csv, sep=Frame.CSV_OPTIONS['sep'], index=True)

    def __repr__(self) -> str:
        return str(self._csv)

    def __str__(self) -> str:
        return self._csv.name


def k_fold_indices(N: int, K: int, shuffle_before_folding: bool = False) -> Dict[int, Tuple[List[int], List[int]]]:
    """ {k: (train rows, test rows)} exactly as ``Repository.into_K_folds`` assigns them (reference storage.py:180-203).

    Consumes the global ``random`` stream in the reference's order: one optional shuffle of range(N), then one shuffle per block
    of the indicator (floor(N/K) copies of range(K) and one range(N % K)).  K > 0 adds the improper fold K = all rows."""
    if not (1 <= abs(K) <= N):
        raise IndexError(f'K={K:d} does not lie between 1 and N={N:d} inclusive.')
    rows = list(range(N))
    if shuffle_before_folding:
        random.shuffle(rows)
    result = {}
    if K > 0:
        result[K] = (list(rows), list(rows))
    K = abs(K)
    blocks = [list(range(K)) for _ in range(int(N / K))] + [list(range(N % K))]
    for block in blocks:
        random.shuffle(block)
    owner = list(itertools.chain(*blocks))
    for k in range(K):
        test = [row for row, o in zip(rows, owner) if o == k]
        train = [row for row, o in zip(rows, owner) if o != k]
        result[k] = (train if train else test, test)
    return result


class Repository:
    """ A folder holding ``data.csv`` + ``meta.json``: the global data set, to be split into Folds."""

    class _InitMode(IntEnum):
        READ_META_ONLY = auto()
        READ = auto()
        CREATE = auto()

    def __init__(self, folder: Path | str, **kwargs):
        self._folder = Path(folder)
        self._meta_json = self._folder / 'meta.json'
        self._csv = self._folder / 'data.csv'
        self._data = None
        init_mode = kwargs.get('init_mode', Repository._InitMode.READ)
        if init_mode <= Repository._InitMode.READ:
            self._meta = self.read_meta()
            if init_mode is Repository._InitMode.READ:
                self._data = Frame(self._csv)
        else:
            shutil.rmtree(self._folder, ignore_errors=True)
            self._folder.mkdir(mode=0o777, parents=True, exist_ok=False)

    @classproperty
    def META(cls) -> Dict[str, Any]:
        return {'csv_kwargs': Frame.CSV_OPTIONS, 'data': {}, 'K': 0, 'shuffle before folding': False}

    @classproperty
    def CSV_OPTIONS(cls) -> Dict[str, Any]:
        return {'skiprows': None, 'index_col': 0}

    @classmethod
    def from_df(cls, folder: Path | str, df: pd.DataFrame, meta: Dict | None = None) -> 'Repository':
        repo = Repository(folder, init_mode=Repository._InitMode.CREATE)
        repo._meta = cls.META | ({} if meta is None else meta)
        repo._data = Frame(repo._csv, df)
        repo._update_meta()
        return repo

    @classmethod
    def from_csv(cls, folder: Path | str, csv: Path | str, meta: Dict = None, **kwargs) -> 'Repository':
        csv = Path(csv)
        origin_csv_kwargs = cls.CSV_OPTIONS | kwargs
        data = Frame(csv, **origin_csv_kwargs)
        meta = cls.META if meta is None else cls.META | meta
        meta['origin'] = {'csv': str(csv.absolute()), 'origin_csv_kwargs': origin_csv_kwargs}
        return cls.from_df(folder, data.df, meta)

    @property
    def folder(self) -> Path:
        return self._folder

    @property
    def data(self) -> Frame:
        return self._data

    @property
    def X(self) -> pd.DataFrame:
        """ The input X, as an (N,M) design Matrix with column headings."""
        return self._data.df[self._meta['data']['X_heading']]

    @property
    def Y(self) -> pd.DataFrame:
        """ The output Y as an (N,L) Matrix with column headings."""
        return self._data.df[self._meta['data']['Y_heading']]

    def read_meta(self) -> Dict[str, Any]:
        with open(self._meta_json, mode='r') as file:
            return json.load(file)

    def write_meta(self):
        with open(self._meta_json, mode='w') as file:
            json.dump(self._meta, file, indent=8)

    @property
    def meta(self) -> Dict[str, Any]:
        return self._meta

    def _update_meta(self):
        columns = self._data.df.columns.values
        self._meta.update({'data': {'X_heading': columns[0][0], 'Y_heading': columns[-1][0]}})
        self._meta['data'].update({'N': self.data.df.shape[0], 'M': self.X.shape[1], 'L': self.Y.shape[1]})
        self.write_meta()

    @property
    def N(self) -> int:
        return self._meta['data']['N']

    @property
    def M(self) -> int:
        return self._meta['data']['M']

    @property
    def L(self) -> int:
        return self._meta['data']['L']

    @property
    def K(self) -> int:
        return self._meta['K']

    @property
    def folds(self) -> range:
        """ The indices of the folds contained in this Repository (the improper fold K included when present)."""
        if isinstance(self, Fold) or self.K < 1:
            return range(0, 0)
        return range(self.K + (1 if self.meta['has_improper_fold'] else 0))

    def fold_folder(self, k: int) -> Path:
        return self._folder / f'fold.{k:d}'

    def into_K_folds(self, K: int, shuffle_before_folding: bool = False, normalization: Optional[Path | str] = None,
                     is_normalization_applicable: bool = True) -> 'Repository':
        """ Fold this repo into |K| Folds indexed by range(|K|); K > 0 adds the improper fold K (all data train and test)."""
        data = self.data.df
        N = data.shape[0]
        if not (1 <= abs(K) <= N):
            raise IndexError(f'K={K:d} does not lie between 1 and N={N:d} inclusive.')
        for k in range(max(abs(K), self.K) + 1):
            shutil.rmtree(self.fold_folder(k), ignore_errors=True)
        assignment = k_fold_indices(N, K, shuffle_before_folding)      # consumes `random` before any file is written, as the reference
        self._meta.update({'K': abs(K), 'shuffle before folding': shuffle_before_folding, 'has_improper_fold': K > 0})
        self.write_meta()
        normalization = Normalization(self, self._data.df).csv if normalization is None else normalization
        for k in ([K] if K > 0 else []) + list(range(abs(K))):
            train, test = assignment[k]
            Fold.from_dfs(parent=self, k=k, data=data.iloc[train], test_data=data.iloc[test], normalization=normalization,
                          is_normalization_applicable=is_normalization_applicable)
        return self

    def rotate_folds(self, rotation: NP.Matrix | None) -> 'Repository':
        """ The reference can rotate the input basis of every fold (an OUT-OF-SCOPE data-preparation feature).  Its scripts
        call this with ``None`` - the identity - which is all that is supported here."""
        if rotation is not None:
            raise NotImplementedError('input rotations are not part of the B200 hot path; rotate the data before building the Repository.')
        return self

    def __repr__(self) -> str:
        return str(self._folder)

    def __str__(self) -> str:
        return self._folder.name


class Fold(Repository):
    """ A Repository equipped with test data (``test.csv``) and a Normalization."""

    def __init__(self, parent: Repository, k: int, **kwargs):
        init_mode = kwargs.get('init_mode', Repository._InitMode.READ)
        super().__init__(parent.fold_folder(k), init_mode=init_mode)
        self._test_csv = self.folder / 'test.csv'
        if init_mode == Repository._InitMode.READ:
            self._test_data = Frame(self._test_csv)
            self._normalization = Normalization(self)

    @classmethod
    def from_dfs(cls, parent: Repository, k: int, data: pd.DataFrame, test_data: pd.DataFrame, normalization: Optional[Path | str] = None,
                 is_normalization_applicable: bool = True) -> 'Fold':
        fold = cls(parent, k, init_mode=Repository._InitMode.CREATE)
        fold._meta = cls.META | parent.meta | {'k': k}
        fold._meta['data'] = dict(parent.meta.get('data', {}))
        fold._normalization = Normalization(fold, data, is_normalization_applicable)
        if normalization is not None:
            # As in the reference (storage.py:431-435): the file on disk becomes the repo-wide normalization, while the statistics
            # already held in memory (computed from this fold's training rows) are the ones applied below.
            shutil.copy(Path(normalization), fold._normalization.csv)
        fold._data = Frame(fold._csv, fold.normalization.apply_to(data))
        fold._test_data = Frame(fold._test_csv, fold.normalization.apply_to(test_data))
        fold._update_meta()
        return fold

    @property
    def normalization(self) -> 'Normalization':
        return self._normalization

    @property
    def test_csv(self) -> Path:
        return self._test_csv

    @property
    def test_data(self) -> Frame:
        return self._test_data

    @property
    def test_x(self) -> pd.DataFrame:
        return self._test_data.df[self._meta['data']['X_heading']]

    @property
    def test_y(self) -> pd.DataFrame:
        return self._test_data.df[self._meta['data']['Y_heading']]


class Normalization:
    """ X is taken as uniform on [mean - sqrt(3) std, mean + sqrt(3) std], mapped to U[0,1] then through the probit to N(0,1);
    Y is standardised to zero mean and unit variance."""

    @classproperty
    def UNIFORM_MARGIN(cls) -> float:
        return 1.0E-12

    def __init__(self, fold: Repository, data: Optional[pd.DataFrame] = None, is_applicable: bool = True):
        self._fold = fold
        self._is_applicable = is_applicable
        if self.csv.exists():
            self._frame = Frame(self.csv)
        elif data is None:
            self._frame = None
        else:
            mean, std = data.mean().rename('mean'), data.std().rename('std')
            semi_range = (std * np.sqrt(3)).rename('rng')
            rows = (mean, std, (2 * semi_range).rename('rng'), (mean - semi_range).rename('min'), (mean + semi_range).rename('max'))
            self._frame = Frame(self.csv, pd.concat(rows, axis=1).T)

    @property
    def csv(self) -> Path:
        return self._fold.folder / 'normalization.csv'

    @property
    def frame(self) -> Frame:
        self._frame = Frame(self.csv) if self._frame is None else self._frame
        return self._frame

    @property
    def is_applicable(self) -> bool:
        return self._is_applicable

    @property
    def _relevant_stats(self):
        df, M = self.frame.df, self._fold.M
        return df.loc['min'].iloc[:M], df.loc['rng'].iloc[:M], df.loc['mean'].iloc[M:], df.loc['std'].iloc[M:]

    def apply_to(self, df: pd.DataFrame) -> pd.DataFrame:
        if not self._is_applicable:
            return df
        X_min, X_rng, Y_mean, Y_std = self._relevant_stats
        M = self._fold.M
        X, Y = df.iloc[:, :M].copy(deep=True), df.iloc[:, M:].copy(deep=True)
        U = ((X.values - X_min.values) / X_rng.values).clip(self.UNIFORM_MARGIN, 1 - self.UNIFORM_MARGIN)
        X.iloc[:, :] = scipy.stats.norm.ppf(U, loc=0, scale=1)
        Y.iloc[:, :] = (Y.values - Y_mean.values) / Y_std.values
        return pd.concat((X, Y), axis=1)

    def undo_from(self, df: pd.DataFrame) -> pd.DataFrame:
        """ The inverse of ``apply_to``: N(0,1) inputs back through the normal cdf onto [min, min + rng], outputs back to mean + std * y."""
        if not self._is_applicable:
            return df
        lo, span, centre, spread = (stat.to_numpy(dtype=float) for stat in self._relevant_stats)
        M = self._fold.M
        values = df.to_numpy(dtype=float, copy=True)
        values[:, :M] = lo + span * scipy.stats.norm.cdf(values[:, :M])
        values[:, M:] = centre + spread * values[:, M:]
        return pd.DataFrame(values, index=df.index, columns=df.columns)

    # ---- the same maps on the device (rc_column_stats / rc_normalize): samples that are already in HBM - synthetic designs, predictions on their
    #      way back to physical units - are normalised where they are.  Device only: these raise without the CUDA library.
    def stats_tensor(self):
        """ The (5, M + L) statistics of this Normalization (rows mean, std, rng, min, max) as a device tensor."""
        from romcomma._tensors import as_device
        df = self.frame.df
        return as_device(np.stack([df.loc[row].to_numpy(dtype=float) for row in ('mean', 'std', 'rng', 'min', 'max')]))

    @staticmethod
    def stats_of_tensor(data):
        """ The statistics ``__init__`` computes from ``data`` (N, M + L), on the device (reference data/storage.py:544-558)."""
        from romcomma import _capi
        return _capi.column_stats(data)

    def apply_to_tensor(self, data, stats=None):
        """ ``apply_to`` for a device tensor (N, M + L); ``stats`` defaults to this Normalization's (reference data/storage.py:469-485)."""
        from romcomma import _capi
        if not self._is_applicable:
            return data
        return _capi.normalize(data, self._fold.M, self.stats_tensor() if stats is None else stats, self.UNIFORM_MARGIN)

    def undo_from_tensor(self, data, stats=None):
        """ ``undo_from`` for a device tensor (N, M + L) (reference data/storage.py:487-503)."""
        from romcomma import _capi
        if not self._is_applicable:
            return data
        return _capi.normalize(data, self._fold.M, self.stats_tensor() if stats is None else stats, self.UNIFORM_MARGIN, undo=True)

    def unscale_Y(self, dfY: pd.DataFrame) -> pd.DataFrame:
        """ Standard deviations of normalised outputs in the units of the raw outputs (a scale, so no shift)."""
        if not self._is_applicable:
            return dfY
        return dfY * self._relevant_stats[3].to_numpy(dtype=float)

    def __repr__(self) -> str:
        return str(self.csv)

    def __str__(self) -> str:
        return self.csv.name
