""" ``user.run.gpr`` / ``user.run.gsa``: fit, test and analyse GPs over one Fold or over all Folds of a Repository, for one model or for the
whole (variant|covariant) x (isotropic|anisotropic) hierarchy.  Same call signatures, model names (``<name>.{v|c}.{i|a}``), seeding rules,
files and return values as the reference (romcomma/user/run.py:35-158); organised differently:

* the requested hierarchy is first flattened into an ordered list of ``Stage``s, which are then carried out in turn - instead of the reference's
  recursion through ``None`` flags;
* the folds of a repository are independent (reference :60-61, :132-133 loops over them).  The folds owned by this process are fitted side by
  side with their LML evaluations batched (``romcomma.lockstep``); with several processes (torchrun, one per GPU) fold k belongs to rank
  k mod world_size, the ranks agree on success or failure before anyone waits at a barrier, and rank 0 collects the csv files.  A shared file
  system between the ranks is assumed (single node).
"""
from __future__ import annotations

import shutil

from romcomma.base.definitions import *
from romcomma.data.storage import Repository, Fold
from romcomma.gpr.kernels import Kernel
from romcomma.gpr.models import GPR, MOGP
from romcomma.gsa.models import GSA, Sobol
from romcomma.user import contexts, results
from romcomma import distributed, lockstep


LOCKSTEP_MAX_N: int = 4096    #: Folds of repositories with at most this many (output, sample) pairs are fitted side by side.


class Stage(NamedTuple):
    """ One model of the hierarchy: where its starting hyper-parameters come from (``is_read``: True = its own folder, False = the defaults /
    the given parameters, None = its nearest fitted ancestor) and which model it is."""
    is_covariant: bool
    is_isotropic: bool
    is_read: bool | None

    def model_name(self, name: str) -> str:
        return f'{name}.{"c" if self.is_covariant else "v"}.{"i" if self.is_isotropic else "a"}'


def stages(is_read: bool | None, is_covariant: bool | None, is_isotropic: bool | None) -> List[Stage]:
    """ The models a call stands for, in execution order.  ``is_covariant=None``: the variant model(s) first, then the covariant one seeded from
    them (anisotropic unless isotropy was asked for explicitly); ``is_isotropic=None``: isotropic first, then anisotropic seeded from it."""
    if is_covariant is None:
        return stages(is_read, False, is_isotropic) + stages(None, True, False if is_isotropic is None else is_isotropic)
    if is_isotropic is None:
        return [Stage(is_covariant, True, is_read), Stage(is_covariant, False, None)]
    return [Stage(is_covariant, is_isotropic, is_read)]


def _seed_from_ancestor(fold: Fold, name: str, stage: Stage) -> bool:
    """ ``is_read=None``: make sure ``fold/<model>`` holds starting values and say whether it does.  An existing folder is used as it is; otherwise
    the variant model of the same isotropy (for a covariant model), else the isotropic model of the same covariance, is copied into place; with
    no ancestor the model starts from the defaults."""
    target = fold.folder / stage.model_name(name)
    if target.exists():
        return True
    ancestors = [Stage(False, stage.is_isotropic, None)] if stage.is_covariant else []
    ancestors.append(Stage(stage.is_covariant, True, None))
    for ancestor in ancestors:
        source = fold.folder / ancestor.model_name(name)
        if source != target and source.exists():
            GPR.Data.copy(src_folder=source, dst_folder=target)
            return True
    return False


def _each_owned_fold(repo: Repository, job: Callable[[Fold], Any], side_by_side: bool) -> List[Any]:
    """ ``job(Fold(repo, k))`` for every fold this process owns - side by side with batched evaluations where the job fits GPs - and agreement
    between the processes on whether everybody succeeded, so that no rank is left waiting at a barrier for one that raised."""
    mine = distributed.shard(list(repo.folds))
    error, done = None, []
    jobs = [(lambda k=k: job(Fold(repo, k))) for k in mine]
    try:
        # Side by side only where one fit leaves the GPU idle: past n = L N ~ 4096 a single evaluation fills it (and K workspaces would not be free)
        done = lockstep.run_together(jobs) if side_by_side and repo.L * repo.N <= LOCKSTEP_MAX_N else [j() for j in jobs]
    except BaseException as exception:
        error = exception
    somebody_failed = distributed.all_reduce_max(0.0 if error is None else 1.0) > 0.0
    if error is not None:
        raise error
    if somebody_failed:
        raise RuntimeError('another process of this job failed on one of its folds; see its log.')
    return done


def gpr(name: str, repo: Repository, is_read: bool | None, is_covariant: bool | None, is_isotropic: bool | None, ignore_exceptions: bool = False,
        kernel_parameters: Kernel.Data | None = None, likelihood_variance: NP.Matrix | None = None,
        is_calibrated: bool = True, is_tested: bool = True, **kwargs) -> List[str]:
    """ Undertake GPR on a Fold, or across the Folds in a Repository.

    Args:
        name: The MOGP name.
        repo: A Fold to house the MOGP, or a Repository containing Folds to house the GPs.
        is_read: True reads kernel and likelihood data from ``fold.folder/<model>``; False uses defaults (or the parameters given here); None seeds
            the model from its nearest ancestor in the (variant -> covariant, isotropic -> anisotropic) hierarchy.
        is_covariant: Whether the outputs are dependent. None runs variant then covariant.
        is_isotropic: Whether the kernel is isotropic. None runs isotropic then anisotropic.
        ignore_exceptions: Whether to continue when a model throws.
        kernel_parameters, likelihood_variance: replace the defaults when given.
        is_calibrated, is_tested: Whether to calibrate / test each MOGP.
        kwargs: passed straight to MOGP.calibrate().
    Returns: The names of the GPs which have been constructed.
    """
    plan = stages(is_read, is_covariant, is_isotropic)
    names = [stage.model_name(name) for stage in plan]

    def fit(fold: Fold):
        for stage in plan:
            model = stage.model_name(name)
            with contexts.Timer(f'fold.{fold.meta["k"]} {model} GPR'):
                try:
                    from_file = _seed_from_ancestor(fold, name, stage) if stage.is_read is None else stage.is_read
                    given = () if from_file else (kernel_parameters, likelihood_variance)
                    gp = MOGP(model, fold, from_file, stage.is_covariant, stage.is_isotropic, *given)
                    if is_calibrated:
                        gp.calibrate(**kwargs)
                    if is_tested:
                        gp.test()
                except BaseException:
                    if not ignore_exceptions:
                        raise

    if isinstance(repo, Fold):
        fit(repo)
        return names
    _each_owned_fold(repo, fit, side_by_side=is_calibrated)
    distributed.barrier()
    if distributed.rank() == 0:
        per_model = {'': ({'test': {'header': [0, 1]}, 'test_summary': {'header': [0, 1], 'index_col': 0}} if is_tested else {}),
                     '/likelihood': {'variance': {}, 'log_marginal': {}}, '/kernel': {'variance': {}, 'lengthscales': {}}}
        for sub_folder, csvs in per_model.items():
            if csvs:
                results.Collect(csvs, {model + sub_folder: {} for model in names}, ignore_exceptions).from_folds(repo, True)
    distributed.barrier()
    return names


def gsa(name: str, repo: Repository, is_covariant: Optional[bool], is_isotropic: Optional[bool],
        kinds: GSA.Kind | Sequence[GSA.Kind] = GSA.ALL_KINDS, m: int = -1,
        ignore_exceptions: bool = False, is_error_calculated: bool = False, **kwargs) -> List[Path]:
    """ Undertake GSA on a Fold, or across the Folds in a Repository.

    Args:
        name: The GP name.
        repo: A Fold, or a Repository containing Folds.
        is_covariant, is_isotropic: select the model as in ``gpr``; None runs both.
        kinds: first_order, closed or total; a Sequence runs consecutively.
        m: a single input ``0 <= m < M``, or anything else for all of them.
        is_error_calculated: also compute the standard errors T and covariances W of the indices (ClosedSobolWithError).
        kwargs: calculation options which update the GSA META (e.g. ``is_T_partial``).
    Returns: The calculation folders which have been written, relative to repo.folder.
    """
    kinds = (kinds,) if isinstance(kinds, GSA.Kind) else tuple(kinds)
    plan = stages(True, is_covariant, is_isotropic)

    def analyse(fold: Fold) -> List[Path]:
        written = []
        for stage in plan:
            model = stage.model_name(name)
            with contexts.Timer(f'fold.{fold.meta["k"]} {model} GSA'):
                try:
                    gp = MOGP(model, fold, is_read=True, is_covariant=stage.is_covariant, is_isotropic=stage.is_isotropic)
                    for kind in kinds:
                        folder = Sobol(gp, kind, m, is_error_calculated, **kwargs).calibrate().get('folder')
                        written.append(Path(folder).relative_to(fold.folder))
                except BaseException:
                    if not ignore_exceptions:
                        raise
        return written

    if isinstance(repo, Fold):
        return analyse(repo)
    per_fold = _each_owned_fold(repo, analyse, side_by_side=False)      # nothing to batch: threads would only contend for the interpreter
    distributed.barrier()
    written = per_fold[-1] if per_fold else []
    if distributed.rank() == 0 and len(repo.folds) > 0:
        if not written:                # this rank owns no fold: read the calculation folders off the first fold
            first = Fold(repo, repo.folds.start)
            written = sorted(p.parent.relative_to(first.folder) for p in first.folder.glob(f'{name}.*/gsa/*/S.csv'))
        csvs = {'S': {}, 'V': {}} | ({'T': {}, 'W': {}} if is_error_calculated else {})
        results.Collect(csvs, {calculation: {} for calculation in written}, ignore_exceptions).from_folds(repo, True)
        for calculation in written:
            shutil.copyfile(repo.fold_folder(repo.folds.start) / calculation / 'meta.json', repo.folder / calculation / 'meta.json')
    distributed.barrier()
    return written
