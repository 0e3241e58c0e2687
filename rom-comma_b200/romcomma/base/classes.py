""" Storage-backed base classes: ``Frame`` (a DataFrame mirrored in one csv), ``Data`` (a folder of Frames) and ``Model``
(a folder holding Data plus a meta.json).  Keeps the on-disk layout of reference romcomma/base/classes.py:34-321:
one ``<field>.csv`` per matrix with pandas defaults (index column 0), ``meta.json`` with indent=8."""
from __future__ import annotations

import json
import shutil
from abc import ABC

from romcomma.base.definitions import *


def _host(value):
    """numpy view of anything carrying numbers (Frame, torch tensor, HostTensor, ...)."""
    if isinstance(value, Frame):
        return value.np
    if isinstance(value, torch.Tensor):
        return value.detach().cpu().numpy()
    if hasattr(value, 'numpy') and not isinstance(value, np.ndarray):
        return value.numpy()
    return value


class Frame:
    """ A pandas DataFrame backed by ``<csv>.csv``: every mutation is written through."""

    def __init__(self, csv: Path | str, data=None, index=None, columns=None, dtype=None, copy=None, **kwargs):
        """ ``data is None`` reads the csv (kwargs -> pd.read_csv, index_col=0 by default), otherwise the data are written (kwargs -> to_csv)."""
        self.csv = Path(csv)
        self._write_options = {}
        if data is None:
            self._df = pd.read_csv(self._file, **({'index_col': 0} | kwargs))
        else:
            self._df = pd.DataFrame(_host(data), index, columns, dtype, copy)
            self.write(**kwargs)

    @property
    def _file(self) -> Path:
        return self.csv.with_suffix(f'{self.csv.suffix}.csv')

    @property
    def df(self) -> pd.DataFrame:
        return self._df

    @df.setter
    def df(self, value: pd.DataFrame):
        self._df = value

    @property
    def np(self) -> NP.Matrix:
        return self._df.values

    @np.setter
    def np(self, value):
        self._df.iloc[:, :] = _host(value)
        self.write()

    @property
    def tf(self):
        """ The matrix as a device tensor (the reference returns a tf.Tensor here)."""
        from romcomma._tensors import DeviceTensor, as_device
        return DeviceTensor.wrap(as_device(self.np.astype(float)))

    @tf.setter
    def tf(self, value):
        self.np = value

    def write(self, **kwargs: Any) -> 'Frame':
        self._write_options |= kwargs
        self._df.to_csv(self._file, **self._write_options)
        return self

    def broadcast_value(self, target_shape: Tuple[int, int], is_diagonal: bool = True) -> 'Frame':
        """ Broadcast to ``target_shape``; a square target keeps only its diagonal when ``is_diagonal``. Raises IndexError on shrinkage."""
        try:
            values = np.array(np.broadcast_to(self.np, target_shape))
        except ValueError:
            raise IndexError(f'{repr(self)} has shape {self.df.shape} which cannot be broadcast to {target_shape}.')
        if is_diagonal and target_shape[0] > 1:
            values = np.diag(np.diagonal(values))
        self._df = pd.DataFrame(values)
        return self.write()

    def __call__(self, *args, **kwargs):
        return self.np

    def __repr__(self) -> str:
        return str(self.csv)

    def __str__(self) -> str:
        return self.csv.name


def _wipe(folder: Path | str) -> Path:
    """ Remove ``folder`` and everything below it (missing is fine); returns it as a Path."""
    target = Path(folder)
    shutil.rmtree(target, ignore_errors=True)
    return target


def _fresh(folder: Path | str) -> Path:
    """ An existing, empty ``folder``."""
    target = _wipe(folder)
    target.mkdir(mode=0o777, parents=True, exist_ok=False)
    return target


class Data(ABC):
    """ A NamedTuple of Frames living in one folder. Subclasses provide ``NamedTuple`` (fields with default matrices)."""

    Matrix = Any

    class NamedTuple(NamedTuple):
        NotImplemented: Any = np.atleast_2d('NotImplemented')

    # ---- construction --------------------------------------------------------------------------------------------------------------------
    def __init__(self, folder: Path | str, **kwargs):
        """ Fields missing from ``kwargs`` take the defaults of ``NamedTuple``; a folder that does not exist yet is created empty."""
        where = Path(folder)
        self._folder = where if where.exists() else _fresh(where)
        self._frames = None
        self.replace(**self.NamedTuple(**kwargs)._asdict())

    @classmethod
    def read(cls, folder: Path | str, **kwargs) -> 'Data':
        """ The Data stored in ``folder``; a field given in ``kwargs`` is written instead of read."""
        where = Path(folder)
        stored = {name: Frame(where / name, kwargs.get(name)) for name in cls.fields}
        return cls(where, **stored)

    @classmethod
    def make(cls, iterable: Iterable):
        return cls.NamedTuple._make(iterable)

    # ---- the field set -------------------------------------------------------------------------------------------------------------------
    @classproperty
    def fields(cls) -> Tuple[str, ...]:
        return cls.NamedTuple._fields

    @classproperty
    def field_defaults(cls) -> Dict[str, Any]:
        return cls.NamedTuple._field_defaults

    @property
    def frames(self):
        return self._frames

    def __call__(self, *args, **kwargs):
        return self._frames

    def asdict(self) -> Dict[str, Any]:
        return self._frames._asdict()

    def replace(self, **kwargs) -> 'Data':
        """ Overwrite the named fields (anything matrix-like, or a Frame) - each is written to ``<folder>/<field>.csv``."""
        as_frames = {name: (value if isinstance(value, Frame) else Frame(self._folder / name, np.atleast_2d(_host(value)))) for name, value in kwargs.items()}
        self._frames = self._frames._replace(**as_frames) if self._frames is not None else self.NamedTuple(**as_frames)
        return self

    # ---- the folder ----------------------------------------------------------------------------------------------------------------------
    @property
    def folder(self) -> Path:
        return self._folder

    def move(self, dst_folder: Path | str) -> 'Data':
        """ Re-home this Data in ``dst_folder`` (emptied first)."""
        self._folder = type(self)(_fresh(dst_folder), **self.asdict()).folder
        return self

    @staticmethod
    def delete(folder: Path | str) -> Path:
        return _wipe(folder)

    @staticmethod
    def empty(folder: Path | str) -> Path:
        return _fresh(folder)

    @staticmethod
    def copy(src_folder: Path | str, dst_folder: Path | str) -> Path:
        target = _wipe(dst_folder)
        shutil.copytree(src=src_folder, dst=target)
        return target

    def __repr__(self) -> str:
        return str(self._folder)

    def __str__(self) -> str:
        return self._folder.name


class Model(ABC):
    """ Anything that lives in a folder with a ``Data`` set and a ``meta.json`` and can be calibrated."""

    class Data(Data):
        class NamedTuple(NamedTuple):
            NotImplemented: Any = np.atleast_2d('NotImplemented')

    @classproperty
    def META(cls) -> Dict[str, Any]:
        """ Default meta data."""
        return {}

    @abstractmethod
    def __init__(self, folder: Path | str, read_data: bool = False, **kwargs):
        """ ``read_data``: take the Data from ``folder`` (then apply ``kwargs``) instead of starting from the defaults."""
        self._folder = Path(folder)
        self._implementation = None
        if not read_data:
            self._folder.mkdir(mode=0o777, parents=True, exist_ok=True)
        self._data = self.Data.read(self._folder).replace(**kwargs) if read_data else self.Data(self._folder, **kwargs)

    @abstractmethod
    def calibrate(self, method: str, **kwargs) -> Dict[str, Any]:
        raise NotImplementedError('base.calibrate() must never be called.')

    # ---- meta.json (indent 8, as the reference writes it) ---------------------------------------------------------------------------------
    @property
    def _meta_json(self) -> Path:
        return self._folder / 'meta.json'

    def read_meta(self) -> Dict[str, Any]:
        return json.loads(self._meta_json.read_text())

    def write_meta(self, meta: Dict[str, Any]):
        self._meta_json.write_text(json.dumps(meta, indent=8))

    # ---- accessors -----------------------------------------------------------------------------------------------------------------------
    @property
    def folder(self) -> Path:
        return self._folder

    @property
    def data(self) -> Data:
        return self._data

    @data.setter
    def data(self, value: Data):
        self._data = value

    def __repr__(self) -> str:
        return str(self._folder)

    def __str__(self) -> str:
        return self._folder.name
