""" User interface to GPR and GSA (reference romcomma/user/run.py:35-158): recursion over folds and over the
(variant|covariant) x (isotropic|anisotropic) model hierarchy, model names ``<name>.{v|c}.{i|a}``.

Folds are independent, so under torchrun (one process per GPU) fold k runs on rank k % world_size; rank 0 collects the csvs after a
barrier.  Single-process behaviour is the reference's sequential loop."""
from __future__ import annotations

import shutil

from romcomma.base.definitions import *
from romcomma.data.storage import Repository, Fold
from romcomma.gpr.kernels import Kernel
from romcomma.gpr.models import GPR, MOGP
from romcomma.gsa.models import GSA, Sobol
from romcomma.user import contexts, results
from romcomma import distributed


def gpr(name: str, repo: Repository, is_read: bool | None, is_covariant: bool | None, is_isotropic: bool | None, ignore_exceptions: bool = False,
        kernel_parameters: Kernel.Data | None = None, likelihood_variance: NP.Matrix | None = None,
        is_calibrated: bool = True, is_tested: bool = True, **kwargs) -> List[str]:
    """ Undertake GPR on a Fold, or recursively across the Folds in a Repository.

    Args:
        name: The MOGP name.
        repo: A Fold to house the MOGP, or a Repository containing Folds to house the GPs.
        is_read: True reads kernel and likelihood data from ``fold.folder/name``; False uses defaults; None seeds the model from its nearest
            ancestor in the (variant -> covariant, isotropic -> anisotropic) hierarchy, constructing that first if necessary.
        is_covariant: Whether the outputs are dependent. None runs variant then covariant.
        is_isotropic: Whether the kernel is isotropic. None runs isotropic then anisotropic.
        ignore_exceptions: Whether to continue when a fold throws.
        kernel_parameters, likelihood_variance: replace the defaults when given.
        is_calibrated, is_tested: Whether to calibrate / test each MOGP.
        kwargs: passed straight to MOGP.calibrate().
    Returns: The names of the GPs which have been constructed.
    """
    if not isinstance(repo, Fold):
        names = []
        for k in distributed.shard(list(repo.folds)):
            names = gpr(name, Fold(repo, k), is_read, is_covariant, is_isotropic, ignore_exceptions, kernel_parameters, likelihood_variance,
                        is_calibrated, is_tested, **kwargs)
        names = _agree_on_names(name, names, is_covariant, is_isotropic)
        distributed.barrier()
        if distributed.rank() == 0:
            if is_tested:
                results.Collect({'test': {'header': [0, 1]}, 'test_summary': {'header': [0, 1], 'index_col': 0}},
                                {name: {} for name in names}, ignore_exceptions).from_folds(repo, True)
            results.Collect({'variance': {}, 'log_marginal': {}}, {f'{name}/likelihood': {} for name in names}, ignore_exceptions).from_folds(repo, True)
            results.Collect({'variance': {}, 'lengthscales': {}}, {f'{name}/kernel': {} for name in names}, ignore_exceptions).from_folds(repo, True)
        distributed.barrier()
        return names
    if is_covariant is None:
        names = gpr(name, repo, is_read, False, is_isotropic, ignore_exceptions, kernel_parameters, likelihood_variance, is_calibrated, is_tested, **kwargs)
        return names + gpr(name, repo, None, True, False if is_isotropic is None else is_isotropic, ignore_exceptions,
                           kernel_parameters, likelihood_variance, is_calibrated, is_tested, **kwargs)
    full_name = name + ('.c' if is_covariant else '.v')
    if is_isotropic is None:
        names = gpr(name, repo, is_read, is_covariant, True, ignore_exceptions, kernel_parameters, likelihood_variance, is_calibrated, is_tested, **kwargs)
        return names + gpr(name, repo, None, is_covariant, False, ignore_exceptions, kernel_parameters, likelihood_variance, is_calibrated, is_tested, **kwargs)
    full_name = full_name + ('.i' if is_isotropic else '.a')
    if is_read is None:
        if not (repo.folder / full_name).exists():
            nearest_name = name + '.v' + full_name[-2:]
            if not (is_covariant and (repo.folder / nearest_name).exists()):
                nearest_name = full_name[:-2] + '.i'
                if not (repo.folder / nearest_name).exists():
                    return gpr(name, repo, False, is_covariant, is_isotropic, ignore_exceptions, kernel_parameters, likelihood_variance,
                               is_calibrated, is_tested, **kwargs)
            GPR.Data.copy(src_folder=repo.folder / nearest_name, dst_folder=repo.folder / full_name)
        return gpr(name, repo, True, is_covariant, is_isotropic, ignore_exceptions, kernel_parameters, likelihood_variance, is_calibrated, is_tested, **kwargs)
    with contexts.Timer(f'fold.{repo.meta["k"]} {full_name} GPR'):
        try:
            if is_read:
                gp = MOGP(full_name, repo, is_read, is_covariant, is_isotropic)
            else:
                gp = MOGP(full_name, repo, is_read, is_covariant, is_isotropic, kernel_parameters, likelihood_variance)
            if is_calibrated:
                gp.calibrate(**kwargs)
            if is_tested:
                gp.test()
        except BaseException as exception:
            if not ignore_exceptions:
                raise exception
    return [full_name]


def _agree_on_names(name: str, names: List[str], is_covariant, is_isotropic) -> List[str]:
    """ A rank that owns no fold still needs the list of model names for Collect: derive it from the flags (same rule as the recursion)."""
    if names:
        return names
    cov = [False, True] if is_covariant is None else [is_covariant]
    out = []
    for c in cov:
        iso = ([True, False] if is_isotropic is None else [is_isotropic]) if not (is_covariant is None and c) else [False if is_isotropic is None else is_isotropic]
        out += [name + ('.c' if c else '.v') + ('.i' if i else '.a') for i in iso]
    return out


def gsa(name: str, repo: Repository, is_covariant: Optional[bool], is_isotropic: Optional[bool],
        kinds: GSA.Kind | Sequence[GSA.Kind] = GSA.ALL_KINDS, m: int = -1,
        ignore_exceptions: bool = False, is_error_calculated: bool = False, **kwargs) -> List[Path]:
    """ Undertake GSA on a Fold, or recursively across the Folds in a Repository.

    Args:
        name: The GP name.
        repo: A Fold, or a Repository containing Folds.
        is_covariant, is_isotropic: select the model as in ``gpr``; None runs both.
        kinds: first_order, closed or total; a Sequence runs consecutively.
        m: a single input ``0 <= m < M``, or anything else for all of them.
        is_error_calculated: also compute the standard errors T and covariances W of the indices (ClosedSobolWithError).
        kwargs: calculation options which update the GSA META.
    Returns: The calculation folders which have been written, relative to repo.folder.
    """
    kinds = (kinds,) if isinstance(kinds, GSA.Kind) else kinds
    if not isinstance(repo, Fold):
        names = []
        for k in distributed.shard(list(repo.folds)):
            names = gsa(name, Fold(repo, k), is_covariant, is_isotropic, kinds, m, ignore_exceptions, is_error_calculated, **kwargs)
        distributed.barrier()
        if distributed.rank() == 0 and len(repo.folds) > 0:
            if not names:
                first = Fold(repo, repo.folds.start)
                names = sorted(p.parent.relative_to(first.folder) for p in first.folder.glob(f'{name}.*/gsa/*/S.csv'))
            results.Collect({'S': {}, 'V': {}} | ({'T': {}, 'W': {}} if is_error_calculated else {}),
                            {name: {} for name in names}, ignore_exceptions).from_folds(repo, True)
            for name in names:
                shutil.copyfile(repo.fold_folder(repo.folds.start) / name / 'meta.json', repo.folder / name / 'meta.json')
        distributed.barrier()
        return names
    if is_covariant is None:
        names = gsa(name, repo, False, is_isotropic, kinds, m, ignore_exceptions, is_error_calculated, **kwargs)
        return names + gsa(name, repo, True, False if is_isotropic is None else is_isotropic, kinds, m, ignore_exceptions, is_error_calculated, **kwargs)
    full_name = name + ('.c' if is_covariant else '.v')
    if is_isotropic is None:
        names = gsa(name, repo, is_covariant, True, kinds, m, ignore_exceptions, is_error_calculated, **kwargs)
        return names + gsa(name, repo, is_covariant, False, kinds, m, ignore_exceptions, is_error_calculated, **kwargs)
    full_name = full_name + ('.i' if is_isotropic else '.a')
    names = []
    with contexts.Timer(f'fold.{repo.meta["k"]} {full_name} GSA'):
        try:
            gp = MOGP(full_name, repo, is_read=True, is_covariant=is_covariant, is_isotropic=is_isotropic)
            for kind in kinds:
                folder = Sobol(gp, kind, m, is_error_calculated, **kwargs).calibrate().get('folder')
                names += [Path(folder).relative_to(repo.folder)]
        except BaseException as exception:
            if not ignore_exceptions:
                raise exception
    return names
