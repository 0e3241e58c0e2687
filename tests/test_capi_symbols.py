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


@pytest.mark.parametrize('tiles_m,tiles_n,lower,kmode,sel', [
    (1, 1, 0, 0, 0), (5, 3, 0, 0, 0), (13, 25, 0, 1, 0), (64, 64, 0, 1, 0), (64, 64, 0, 2, 0), (37, 12, 0, 2, 0), (12, 37, 0, 4, 0), (30, 7, 0, 3, 0),
    (1, 1, 1, 0, 0), (11, 11, 1, 0, 0), (12, 12, 1, 0, 0), (13, 13, 1, 0, 0), (124, 124, 1, 0, 0), (25, 25, 1, 3, 0), (128, 128, 1, 3, 4096), (48, 48, 1, 3, 1536)])
def test_gemm_tile_order_visits_every_tile_once(tiles_m, tiles_n, lower, kmode, sel):
    """The L2-blocked raster of the GEMM's dynamic tile scheduler (csrc/gemm_dmma.cuh: gemm_decode_tile, host-callable through the test hook
    rc_debug_tile_order): every tile of the (full or lower-triangular) list exactly once, ragged super-tiles included; the K range of each
    tile is what its kmode prescribes; tiles come in 12 x 12 super-tiles (any 144 consecutive tiles of a full list touch at most 48 tile rows and
    48 tile columns - four super-tiles when a ragged one lies in between); lists whose K range varies start with the longest ranges."""
    import numpy as np
    from romcomma import _capi
    lib = _capi.lib()
    M, N, K = 128 * tiles_m, 128 * tiles_n, 128 * (max(tiles_m, tiles_n) if kmode else 4)
    count = tiles_m * (tiles_m + 1) // 2 if lower else tiles_m * tiles_n
    out = np.full((count, 4), -7, dtype=np.int32)
    got = lib.rc_debug_tile_order(M, N, K, lower, kmode, sel, out.ctypes.data_as(ctypes.c_void_p))
    assert got == count
    m0, n0, kb, nk = out.T
    assert np.all(m0 % 128 == 0) and np.all(n0 % 128 == 0) and np.all((0 <= m0) & (m0 < M)) and np.all((0 <= n0) & (n0 < N))
    if lower:
        assert np.all(n0 <= m0)
    assert len({(a, b) for a, b in zip(m0.tolist(), n0.tolist())}) == count, 'a tile is handed out twice'
    kb_ref = np.where(kmode == 1, n0, np.where(kmode == 3, m0, 0))
    ke_ref = np.where(kmode == 2, np.minimum(K, m0 + 128), np.where(kmode == 4, np.minimum(K, n0 + 128), K))
    live = nk >= 0
    if sel:
        assert np.array_equal(~live, m0 // sel > (n0 + 127) // sel)
    assert np.array_equal(kb[live], kb_ref[live]) and np.array_equal(nk[live] * 16, (ke_ref - kb_ref)[live])
    if not lower and count >= 144:
        for start in range(0, count - 143, 97):
            window = slice(start, start + 144)
            # 144 consecutive tiles touch at most four super-tiles (a ragged one in between): <= 48 tile rows / columns, not all of them
            assert len(set(m0[window].tolist())) <= 48 and len(set(n0[window].tolist())) <= 48
    if kmode and not sel:
        lengths = nk * 16
        assert lengths[0] == lengths.max(), 'the list must start with the longest K range'
        assert lengths[-1] <= lengths[: max(1, count // 4)].min(), 'the shortest ranges belong at the end'


@pytest.mark.parametrize('tiles,cols', [(20, 4), (7, 2), (5, 8), (9, 1), (128, 4)])
def test_gemm_column_limited_lower_list(tiles, cols):
    """The look-ahead panel update of the factorisation (chol.cu: potrf_lookahead, U_g) runs over the first `cols` tile columns of a lower
    triangle (negative sel_block selects the list in the test hook): every such tile exactly once, full K, the triangle on the diagonal first."""
    import numpy as np
    from romcomma import _capi
    lib = _capi.lib()
    w = min(cols, tiles)
    count = w * (w + 1) // 2 + (tiles - w) * w
    out = np.full((count, 4), -7, dtype=np.int32)
    assert lib.rc_debug_tile_order(128 * tiles, 128 * tiles, 512, 1, 0, -cols, out.ctypes.data_as(ctypes.c_void_p)) == count
    m0, n0, kb, nk = out.T
    want = {(tm * 128, tn * 128) for tm in range(tiles) for tn in range(min(tm + 1, w))}
    assert {(a, b) for a, b in zip(m0.tolist(), n0.tolist())} == want and len(want) == count
    assert np.all(kb == 0) and np.all(nk == 32)
    assert np.all(np.diff(m0) >= 0), 'row-major: a strip of tiles shares its row panel'
