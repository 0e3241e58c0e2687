""" Base classes and helpers underpinning GSA (reference romcomma/gsa/base.py:33-151).

``Gaussian`` keeps the reference's broadcasting contract for callers that use it directly on small tensors; the production Sobol
path does not go through it - the whole Gaussian-ratio chain is fused inside rc_sobol_contract."""
from __future__ import annotations

from abc import ABC

from romcomma.base.definitions import *
from romcomma._tensors import as_device


def diag_det(tensor):
    """ Determinant of a diagonal tensor whose last axis holds the diagonal: ``[...,m] -> [...]``."""
    return torch.prod(as_device(tensor), dim=-1)


class Calibrator(ABC):
    """ Interface to a GSA calibrator."""

    @abstractmethod
    def marginalize(self, m: TF.Slice) -> Dict[str, Any]:
        raise NotImplementedError('This is an abstract class.')


class Gaussian:
    """ An un-normalised Gaussian pdf: ``exponent`` = -z^T Sigma^-1 z / 2 and ``cho_diag`` = diag(chol(Sigma)); the 2 pi factor is left out."""

    def __init__(self, mean, variance, is_variance_diagonal: bool, ordinate=None, LBunch: int = 2):
        """
        Args:
            mean: Population mean, of adequate rank to broadcast the L axes.
            variance: Population variance; an M-vector per batch entry if ``is_variance_diagonal`` else an (M,M) matrix.
            ordinate: The z-value (default 0).
            LBunch: How many consecutive output (L) axes to count before an N axis is inserted for broadcasting.
        """
        mean, variance = as_device(mean), as_device(variance)
        ordinate = torch.zeros((), dtype=torch.float64, device=mean.device) if ordinate is None else as_device(ordinate)
        cho = torch.sqrt(variance) if is_variance_diagonal else torch.linalg.cholesky(variance)
        if tuple(ordinate.shape) == tuple(mean.shape):
            shape = list(ordinate.shape)
            ones = [1] * (len(shape) - 1)
            ordinate, mean = ordinate.reshape(shape[:-1] + ones + [shape[-1]]), mean.reshape(ones + shape)
        z = ordinate - mean
        insertions = cho.dim() - (1 if is_variance_diagonal else 2)
        insertions -= insertions % LBunch
        for axis in range(insertions, 0, -LBunch):
            cho = cho.unsqueeze(axis)
        if is_variance_diagonal:
            z = z / torch.broadcast_to(cho, tuple(cho.shape[:-2]) + tuple(z.shape[-2:]))
            self.cho_diag = cho
        else:
            z = torch.linalg.solve_triangular(cho, z.unsqueeze(-1), upper=False).squeeze(-1)
            self.cho_diag = torch.diagonal(cho, dim1=-2, dim2=-1)
        self.exponent = -0.5 * torch.einsum('...o,...o->...', z, z)

    @property
    def det(self):
        """ The sqrt of the determinant of the covariance."""
        return torch.prod(self.cho_diag, dim=-1)

    @property
    def pdf(self):
        return torch.exp(self.exponent) / self.det

    def _clone(self) -> 'Gaussian':
        out = Gaussian.__new__(Gaussian)
        out.exponent, out.cho_diag = self.exponent, self.cho_diag
        return out

    def expand_dims(self, axes: Sequence[int]) -> 'Gaussian':
        out = self._clone()
        for axis in sorted(axes, reverse=True):
            out.exponent = out.exponent.unsqueeze(axis)
            out.cho_diag = out.cho_diag.unsqueeze((axis - 1) if axis < 0 else axis)
        return out

    def __truediv__(self, other: 'Gaussian') -> 'Gaussian':
        out = self._clone()
        out.exponent = self.exponent - other.exponent
        out.cho_diag = self.cho_diag / other.cho_diag
        return out


def sym_check(tensor, transposition: List[int]):
    tensor = as_device(tensor)
    return torch.sum((tensor - tensor.permute(transposition)) ** 2)


def mean(tensor):
    tensor = as_device(tensor)
    return torch.sum(tensor) / tensor.numel()


def sos(tensor, ein: str = 'lijk, lijk'):
    tensor = as_device(tensor)
    return torch.einsum(ein, tensor, tensor)


def ms(tensor, ein: str = 'lijk, lijk'):
    return sos(tensor, ein) / as_device(tensor).numel()


def rms(tensor, ein: str = 'lijk, lijk'):
    return torch.sqrt(ms(tensor, ein))
