"""CPU: the C-ABI shared library loads and exports every symbol include/romcomma_b200.h declares (no compute calls without a GPU)."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / 'include' / 'romcomma_b200.h'


def declared_symbols():
    text = re.sub(r'/\*.*?\*/', '', HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r'\b(rc_[a-z0-9_]+)\s*\(', text)))


def test_header_declares_the_expected_surface():
    names = declared_symbols()
    for required in ('rc_gram', 'rc_potrf', 'rc_potri', 'rc_trsv', 'rc_trsm_fwd', 'rc_lml_grad', 'rc_predict_reduce', 'rc_sobol_prepare',
                     'rc_sobol_contract', 'rc_last_error'):
        assert required in names


def test_library_exports_every_declared_symbol():
    from romcomma import _capi
    path = _capi.library_path()
    assert path.exists(), f'{path} not built: run __graft_entry__.build()'
    handle = ctypes.CDLL(str(path))
    missing = [name for name in declared_symbols() if not hasattr(handle, name)]
    assert not missing, f'missing exports: {missing}'


def test_python_binding_covers_the_header():
    from romcomma import _capi
    assert sorted(_capi.EXPORTED_SYMBOLS) == declared_symbols()
    lib = _capi.lib()
    assert lib.rc_version() >= 100
    assert lib.rc_padded(1) == 128 and lib.rc_padded(128) == 128 and lib.rc_padded(1843) == 1920
    assert lib.rc_lml_grad_stride(4, 8) == 1 + 2 * 16 + 32
    assert lib.rc_potrf_bufsize(256, 2) >= 2 * 2 * 128 * 128 * 8
    assert lib.rc_lml_grad_bufsize(4096, 8, 4, 1, 1) > 2 * 16384 * 16384 * 8


def test_no_cpu_fallback():
    """The product path must fail loudly without a GPU / with host tensors rather than compute on the CPU."""
    import numpy as np
    import torch
    from romcomma import _capi
    with pytest.raises(_capi.RomcommaB200Error):
        _capi.ptr(torch.zeros(3, dtype=torch.float64))
    if not torch.cuda.is_available():
        from romcomma._tensors import as_device
        with pytest.raises(RuntimeError):
            as_device(np.zeros(3))


def test_product_never_imports_the_oracle():
    for py in (ROOT / 'rom-comma_b200').rglob('*.py'):
        text = py.read_text()
        assert not re.search(r'^\s*(from|import)\s+oracle\b', text, flags=re.M), f'{py} imports the oracle'
