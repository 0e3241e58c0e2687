""" Collecting - concatenating - result csvs across folders or folds (reference romcomma/user/results.py:31-128)."""
from __future__ import annotations

from shutil import rmtree

from romcomma.base.definitions import *
from romcomma.base.classes import Data
from romcomma.data.storage import Repository, Fold


def copy(src: Path | str, dst: Path | str) -> Path:
    Data.copy(src, dst)
    return dst


class Collect:
    """ ``csvs``: {csv stem: pd.read_csv kwargs}; ``folders``: {folder: {column name: value} inserted right-to-left}."""

    write_options: Dict[str, Any] = {'index': False, 'float_format': '%.6f'}

    def __init__(self, csvs: Dict[str, Dict[str, Any]] = None, folders: Dict[str, Dict[str, Any]] = None, ignore_missing: bool = False, **kwargs: Any):
        self.csvs = {} if csvs is None else csvs
        self.folders = {} if folders is None else folders
        self.ignore_missing = ignore_missing
        self.write_options = self.write_options | kwargs

    def __call__(self, dst: Union[Repository, Path, str], is_existing_deleted=False, **kwargs: Any):
        return self.from_folds(dst, is_existing_deleted, **kwargs) if isinstance(dst, Repository) else self.from_folders(dst, is_existing_deleted, **kwargs)

    def _labelled(self, stem: str, read_options: Dict[str, Any]):
        """ The ``<stem>.csv`` of every source folder, each with that folder's label columns put in front (the last label outermost)."""
        for source, labels in self.folders.items():
            path = Path(source) / f'{stem}.csv'
            if self.ignore_missing and not path.exists():
                continue
            table = pd.read_csv(path, **read_options)
            for name, value in labels.items():
                table.insert(0, name, np.full(table.shape[0], value), True)
            yield table

    def from_folders(self, dst: Union[Path, str], is_existing_deleted=False, **kwargs: Any) -> 'Collect':
        target = Path(dst)
        if is_existing_deleted:
            rmtree(target, ignore_errors=True)
        target.mkdir(mode=0o777, parents=True, exist_ok=True)
        options = self.write_options | kwargs
        for stem, read_options in self.csvs.items():
            tables = list(self._labelled(stem, read_options))
            if tables:
                pd.concat(tables, axis=0, ignore_index=True).to_csv(target / f'{stem}.csv', **options)
        return self

    def from_folds(self, dst: Repository, is_existing_deleted=False, **kwargs: Any) -> 'Collect':
        if isinstance(dst, Fold):
            raise NotADirectoryError('dst is a Fold, which cannot contain other Folds, so cannot be Collected from.')
        # fold number and N come from each fold's meta.json: the data are not read for this
        folds = [Fold(dst, k, init_mode=Repository._InitMode.READ_META_ONLY) for k in dst.folds]
        for sub_folder, labels in self.folders.items():
            per_fold = {fold.folder / sub_folder: {'fold': fold.meta['k'], 'N': fold.N} | labels for fold in folds}
            Collect(self.csvs, per_fold, self.ignore_missing).from_folders(dst.folder / sub_folder, is_existing_deleted, **kwargs)
        return self
