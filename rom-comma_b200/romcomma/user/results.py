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

    def from_folders(self, dst: Union[Path, str], is_existing_deleted=False, **kwargs: Any) -> 'Collect':
        dst = Path(dst)
        if is_existing_deleted:
            rmtree(dst, ignore_errors=True)
        dst.mkdir(mode=0o777, parents=True, exist_ok=True)
        for csv, read_options in self.csvs.items():
            pieces = []
            for folder, columns in self.folders.items():
                file = Path(folder) / f'{csv}.csv'
                if file.exists() or not self.ignore_missing:
                    piece = pd.read_csv(file, **read_options)
                    for key, value in columns.items():
                        piece.insert(0, key, np.full(piece.shape[0], value), True)
                    pieces.append(piece)
            if pieces:
                pd.concat(pieces, axis=0, ignore_index=True).to_csv(dst / f'{csv}.csv', **(self.write_options | kwargs))
        return self

    def from_folds(self, dst: Repository, is_existing_deleted=False, **kwargs: Any) -> 'Collect':
        if isinstance(dst, Fold):
            raise NotADirectoryError('dst is a Fold, which cannot contain other Folds, so cannot be Collected from.')
        folds = tuple(Fold(dst, k, init_mode=Repository._InitMode.READ_META_ONLY) for k in dst.folds)    # fold number and N: meta.json, not the data
        for sub_folder, extra_columns in self.folders.items():
            folders = {fold.folder / sub_folder: {'fold': fold.meta['k'], 'N': fold.N} | extra_columns for fold in folds}
            Collect(self.csvs, folders, self.ignore_missing).from_folders(dst.folder / sub_folder, is_existing_deleted, **kwargs)
        return self
