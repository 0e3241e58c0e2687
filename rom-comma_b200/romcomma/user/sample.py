""" Sampling of test functions into Repositories (reference romcomma/user/sample.py:41-254): DOE, GaussianNoise, Function."""
from __future__ import annotations

import scipy.stats

from romcomma.base.definitions import *
from romcomma.data.storage import Frame, Repository, Fold
from romcomma.user import functions


def permute_axes(new_order: Sequence | None) -> NP.Matrix | None:
    """ The rotation matrix that reorders the input axes to ``new_order`` (None -> None)."""
    return None if new_order is None else np.eye(len(new_order))[new_order, :]


class DOE:
    """ Designs of experiment on the unit cube."""

    Method = Callable[..., NP.Matrix]

    @staticmethod
    def latin_hypercube(N: int, M: int, is_centered: bool = True, **kwargs):
        """ An (N,M) latin hypercube; ``is_centered`` puts each sample at the centre of its cell. kwargs (e.g. ``seed``) go to scipy."""
        return scipy.stats.qmc.LatinHypercube(M, scramble=not is_centered, **kwargs).random(N)

    @staticmethod
    def full_factorial(N: int, M: int):
        NM = N // M
        N1 = N - M * NM
        return np.concatenate([1 / (2 * N1) + np.linspace(0, 1, N1, False), ] + (M - 1) * [1 / (2 * NM) + np.linspace(0, 1, NM, False), ], axis=1)


class GaussianNoise:
    """ An (N,L) sample of zero-mean homoskedastic Gaussian noise, drawn once at construction."""

    class Variance:
        """ An (L,L) noise covariance of given magnitude (StdDev): deterministic or random, diagonal or not."""

        def __init__(self, L: int, magnitude: float, is_covariant: bool = False, is_determined: bool = True):
            self.magnitude, self.is_covariant, self.is_determined = magnitude, is_covariant, is_determined
            if self.is_determined:     # sic: the reference draws the random matrix in this branch (sample.py:139-143)
                matrix = 2 * np.random.random_sample((L, L)) - np.ones((L, L))
                matrix = np.matmul(matrix, matrix.transpose())
                matrix /= np.trace(matrix) / L
            else:
                matrix = np.array([[(-1) ** (i - j) / (1.0 + abs(i - j)) for i in range(L)] for j in range(L)])
            if not self.is_covariant:
                matrix = np.diag(np.diag(matrix))
            self._matrix = matrix * self.magnitude ** 2

        @property
        def matrix(self) -> NP.Matrix:
            return self._matrix

        @property
        def meta(self) -> Dict[str, Any]:
            return {'generator': 'determined' if self.is_determined else 'undetermined',
                    'is_covariant': 'covariance' if self.is_covariant else 'variance', 'magnitude': self.magnitude}

        def __call__(self) -> NP.Matrix:
            return self._matrix

        def __format__(self, format_spec: Any) -> str:
            return f'{"d." if self.is_determined else "u."}{"c." if self.is_covariant else "v."}{100 * self.magnitude:.2f}'

    def __init__(self, N: int, variance: NP.MatrixLike):
        self._variance = np.atleast_2d(variance)
        if self._variance.ndim == 2 and self._variance.shape[0] == 1:
            self._variance = np.diagflat(self._variance)
        elif self._variance.ndim > 2 or self._variance.shape[0] != self._variance.shape[1]:
            raise IndexError(f'variance.shape = {self._variance.shape} should be (L,) or (L,L).')
        self._rvs = np.reshape(scipy.stats.multivariate_normal.rvs(mean=None, cov=self._variance, size=N), (N, self._variance.shape[1]))

    @property
    def variance(self) -> NP.Matrix:
        return self._variance

    def __call__(self, repo: Repository | None = None) -> NP.Matrix:
        if repo is not None:
            repo.data.df.iloc[:, :] = np.concatenate((repo.X, repo.Y + self._rvs), axis=1)
            repo.data.write()
        return self._rvs


class Function:
    """ A Repository sampled from a test function vector: ``(X, f(X) + std(f) * noise)``, named
    ``<function>.M.<M>.<noise>.N.<N>[.<ext>]`` under ``root``."""

    def __init__(self, root: Path | str, doe: DOE.Method, function_vector: functions.Vector, N: int, M: int,
                 noise_variance: GaussianNoise.Variance, ext: str | None = None, overwrite_existing: bool = False, **kwargs: Any):
        self._N, self._noise_variance = N, noise_variance
        folder = Path(root) / f'{function_vector.name}.M.{M:d}.{self._noise_variance}.N.{N:d}{"" if ext is None else "." + ext}'
        if folder.is_dir() and not overwrite_existing:
            self._repo = Repository(folder)
        else:
            X = doe(N, M, **kwargs)
            Y = function_vector(X)
            Y += np.reshape(np.std(Y, axis=0), (1, -1)) * GaussianNoise(N, self._noise_variance())(repo=None)
            columns = [('X', f'X.{i:d}') for i in range(X.shape[1])] + [('Y', f'Y.{i:d}') for i in range(Y.shape[1])]
            df = pd.DataFrame(np.concatenate((X, Y), axis=1), columns=pd.MultiIndex.from_tuples(columns), dtype=float)
            origin = {'DOE': doe.__name__, 'function_vector': function_vector.meta, 'noise': self._noise_variance.meta}
            self._repo = Repository.from_df(folder=folder, df=df, meta={'origin': origin})
            pd.DataFrame(self._noise_variance()).to_csv(folder / 'likelihood.variance.csv')

    @property
    def repo(self) -> Repository:
        return self._repo

    def collection(self, sub_folder: Union[Path, str]) -> Dict[str, Any]:
        return {'folder': self._repo.folder / sub_folder, 'N': self._N, 'noise': self._noise_variance.magnitude}

    def into_K_folds(self, K: int, shuffle_before_folding: bool = False, normalization=None, is_normalization_applicable: bool = True) -> 'Function':
        self._repo.into_K_folds(K, shuffle_before_folding, normalization, is_normalization_applicable)
        return self

    def rotate_folds(self, rotation: NP.Matrix | None) -> 'Function':
        self._repo.rotate_folds(rotation)
        return self
