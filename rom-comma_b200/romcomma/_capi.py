"""ctypes binding of ``librc_b200.so`` (include/romcomma_b200.h) with torch CUDA float64 tensors as the carrier.

There is deliberately NO CPU fallback: if the shared library is missing, or a tensor is not a CUDA float64 tensor, the call
raises.  Tensors handed back are ordinary ``torch`` CUDA tensors, so ``torch.utils.dlpack.to_dlpack`` gives the DLPack
capsule a TensorFlow/GPflow caller would consume (``tf.experimental.dlpack.from_dlpack``) without leaving the device.
"""
from __future__ import annotations

import ctypes
import os
from pathlib import Path
from typing import Optional, Sequence

import torch

_LIB_PATH = Path(__file__).resolve().parent.parent / 'csrc' / 'librc_b200.so'
_lib: Optional[ctypes.CDLL] = None

c_double_p = ctypes.c_void_p
RC_GRAD_NONE, RC_GRAD_VARIANCE, RC_GRAD_LENGTHSCALES, RC_GRAD_F_DIAGONAL, RC_NO_OVERLAP = 0, 1, 2, 4, 8

_SIGNATURES = {
    'rc_version': (ctypes.c_int, []),
    'rc_last_error': (ctypes.c_char_p, []),
    'rc_padded': (ctypes.c_int, [ctypes.c_int]),
    'rc_launch_count': (ctypes.c_long, []),
    'rc_measure_dmma_tflops': (ctypes.c_int, [c_double_p, ctypes.c_void_p]),
    'rc_measure_exp_gexps': (ctypes.c_int, [c_double_p, ctypes.c_void_p]),
    'rc_measure_exp_tab_gexps': (ctypes.c_int, [c_double_p, ctypes.c_void_p]),
    'rc_debug_exp': (ctypes.c_int, [c_double_p, c_double_p, ctypes.c_long, ctypes.c_int, ctypes.c_void_p]),
    'rc_debug_tile_order': (ctypes.c_int, [ctypes.c_int] * 6 + [ctypes.c_void_p]),
    'rc_trsm_sbinv_bufsize': (ctypes.c_size_t, [ctypes.c_int, ctypes.c_int]),
    'rc_trsm_sbinv_prepare': (ctypes.c_int, [c_double_p, ctypes.c_int, ctypes.c_long, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    'rc_trsm_fwd_sbinv': (ctypes.c_int, [c_double_p, ctypes.c_int, ctypes.c_long, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, c_double_p, ctypes.c_int,
                                         ctypes.c_long, ctypes.c_void_p]),
    'rc_column_stats': (ctypes.c_int, [c_double_p, ctypes.c_int, ctypes.c_int, c_double_p, ctypes.c_void_p]),
    'rc_normalize': (ctypes.c_int, [c_double_p, ctypes.c_long, ctypes.c_int, ctypes.c_int, c_double_p, ctypes.c_double, ctypes.c_int, c_double_p,
                                    ctypes.c_void_p]),
    'rc_test_metrics': (ctypes.c_int, [c_double_p] * 3 + [ctypes.c_int, ctypes.c_int] + [c_double_p] * 3 + [ctypes.c_void_p]),
    'rc_profile_begin': (ctypes.c_int, []),
    'rc_profile_end': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    'rc_gram': (ctypes.c_int, [c_double_p, ctypes.c_int, c_double_p, ctypes.c_int, ctypes.c_int, c_double_p, ctypes.c_int, c_double_p, c_double_p,
                               c_double_p, ctypes.c_long, ctypes.c_long, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                               ctypes.c_void_p]),
    'rc_apply_variance_noise': (ctypes.c_int, [c_double_p, ctypes.c_long, c_double_p, c_double_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                               c_double_p, ctypes.c_long, ctypes.c_int, ctypes.c_void_p]),
    'rc_potrf_bufsize': (ctypes.c_size_t, [ctypes.c_int, ctypes.c_int]),
    'rc_potrf': (ctypes.c_int, [c_double_p, ctypes.c_int, ctypes.c_long, ctypes.c_long, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                ctypes.c_void_p]),
    'rc_logdet': (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, c_double_p, ctypes.c_void_p]),
    'rc_trsv': (ctypes.c_int, [c_double_p, ctypes.c_int, ctypes.c_long, ctypes.c_long, ctypes.c_int, ctypes.c_void_p, c_double_p, c_double_p,
                               ctypes.c_long, ctypes.c_int, ctypes.c_void_p]),
    'rc_trsm_fwd': (ctypes.c_int, [c_double_p, ctypes.c_int, ctypes.c_long, ctypes.c_long, ctypes.c_int, ctypes.c_void_p, c_double_p, ctypes.c_int,
                                   ctypes.c_long, ctypes.c_long, ctypes.c_void_p]),
    'rc_potri': (ctypes.c_int, [c_double_p, ctypes.c_int, ctypes.c_long, ctypes.c_long, ctypes.c_int, ctypes.c_void_p, c_double_p, ctypes.c_long,
                                ctypes.c_long, ctypes.c_void_p]),
    'rc_pad_identity': (ctypes.c_int, [c_double_p, ctypes.c_int, ctypes.c_long, c_double_p, ctypes.c_int, ctypes.c_long, ctypes.c_long, ctypes.c_int,
                                       ctypes.c_void_p]),
    'rc_extract_lower': (ctypes.c_int, [c_double_p, ctypes.c_long, ctypes.c_long, c_double_p, ctypes.c_int, ctypes.c_long, ctypes.c_int,
                                        ctypes.c_int, ctypes.c_void_p]),
    'rc_lml_grad_stride': (ctypes.c_int, [ctypes.c_int, ctypes.c_int]),
    'rc_lml_grad_bufsize': (ctypes.c_size_t, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    'rc_lml_grad': (ctypes.c_int, [c_double_p, c_double_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_double_p, c_double_p,
                                   c_double_p, c_double_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, c_double_p, ctypes.c_void_p,
                                   ctypes.c_void_p]),
    'rc_lml_grad_multi': (ctypes.c_int, [c_double_p, c_double_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_double_p,
                                         c_double_p, c_double_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, c_double_p, ctypes.c_void_p,
                                         ctypes.c_void_p]),
    'rc_predict_bufsize': (ctypes.c_size_t, [ctypes.c_int, ctypes.c_int]),
    'rc_predict_reduce': (ctypes.c_int, [c_double_p, ctypes.c_long, ctypes.c_long, c_double_p, ctypes.c_long, ctypes.c_int, ctypes.c_int,
                                         ctypes.c_int, ctypes.c_int, ctypes.c_int, c_double_p, c_double_p, ctypes.c_void_p, c_double_p, c_double_p,
                                         ctypes.c_void_p]),
    'rc_syrk_tn': (ctypes.c_int, [c_double_p, ctypes.c_int, ctypes.c_int, ctypes.c_long, ctypes.c_long, ctypes.c_int, ctypes.c_double, ctypes.c_double,
                                  c_double_p, ctypes.c_long, ctypes.c_long, ctypes.c_void_p]),
    'rc_predict_gradient_jacobian': (ctypes.c_int, [c_double_p, ctypes.c_int, ctypes.c_int, c_double_p, ctypes.c_int, c_double_p, c_double_p, c_double_p,
                                                    ctypes.c_int, c_double_p, ctypes.c_long, ctypes.c_long, c_double_p, ctypes.c_void_p]),
    'rc_predict_gradient_finish': (ctypes.c_int, [c_double_p, ctypes.c_long, ctypes.c_long, c_double_p, ctypes.c_int, ctypes.c_int, c_double_p, c_double_p,
                                                  ctypes.c_int, c_double_p, ctypes.c_void_p]),
    'rc_sobol_bufsize': (ctypes.c_size_t, [ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    'rc_sobol_prepare': (ctypes.c_int, [c_double_p, ctypes.c_int, ctypes.c_int, c_double_p, c_double_p, c_double_p, ctypes.c_int, ctypes.c_int,
                                        c_double_p, c_double_p, c_double_p, ctypes.c_void_p]),
    'rc_sobol_contract': (ctypes.c_int, [c_double_p, ctypes.c_int, ctypes.c_int, c_double_p, c_double_p, ctypes.c_int, ctypes.c_int,
                                         ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, c_double_p, ctypes.c_void_p]),
    'rc_sobol_contract_part': (ctypes.c_int, [c_double_p, ctypes.c_int, ctypes.c_int, c_double_p, c_double_p, ctypes.c_int, ctypes.c_int,
                                              ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, c_double_p, ctypes.c_void_p]),
    'rc_sobol_error_bufsize': (ctypes.c_size_t, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    'rc_sobol_error': (ctypes.c_int, [c_double_p, ctypes.c_int, ctypes.c_int, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p,
                                      ctypes.c_int, c_double_p, ctypes.c_int, ctypes.c_long, ctypes.c_long, ctypes.c_int, ctypes.c_void_p,
                                      ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, c_double_p, c_double_p, ctypes.c_void_p]),
    'rc_sobol_error_mixed_bufsize': (ctypes.c_size_t, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    'rc_sobol_error_mixed': (ctypes.c_int, [c_double_p, ctypes.c_int, ctypes.c_int, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p,
                                            ctypes.c_int, c_double_p, ctypes.c_int, ctypes.c_long, ctypes.c_long, ctypes.c_int, ctypes.c_void_p,
                                            ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, c_double_p, c_double_p, c_double_p,
                                            ctypes.c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


class RomcommaB200Error(RuntimeError):
    """A C-ABI call returned a non-zero status."""


def library_path() -> Path:
    return Path(os.environ.get('ROMCOMMA_B200_LIB', _LIB_PATH))


def lib() -> ctypes.CDLL:
    """Load (once) and return the shared library; raises if it has not been built (``__graft_entry__.build()``)."""
    global _lib
    if _lib is None:
        path = library_path()
        if not path.exists():
            raise RomcommaB200Error(f'{path} is missing: build it with `make -C {path.parent}` - there is no CPU fallback.')
        handle = ctypes.CDLL(str(path))
        for name, (restype, argtypes) in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = restype, argtypes
        _lib = handle
    return _lib


def check(status: int, what: str):
    if status != 0:
        raise RomcommaB200Error(f'{what} failed with status {status}: {lib().rc_last_error().decode()}')


def ptr(t: Optional[torch.Tensor]):
    """Device pointer of a contiguous CUDA float64 tensor (None -> NULL)."""
    if t is None:
        return None
    if not (t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()):
        raise RomcommaB200Error(f'expected a contiguous CUDA float64 tensor, got {t.dtype} on {t.device} (contiguous={t.is_contiguous()})')
    return ctypes.c_void_p(t.data_ptr())


def raw_ptr(t: torch.Tensor):
    if not t.is_cuda:
        raise RomcommaB200Error('expected a CUDA tensor')
    return ctypes.c_void_p(t.data_ptr())


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


_REPLAYED_LAUNCHES = 0     # kernels launched by CUDA-graph replays (the C counter only sees eager launches and captures)
_PROFILING = False         # rc_profile_begin/end brackets launches with host-side event records: incompatible with graph replay


def launch_count() -> int:
    """Kernels of this library launched so far in this process, eager or replayed from a captured graph."""
    return int(lib().rc_launch_count()) + _REPLAYED_LAUNCHES


def measure_peaks() -> dict:
    """Live FP64 roofline denominators of the current device: DMMA TFLOP/s and exp evaluations (1e9/s)."""
    scratch = torch.zeros(8, dtype=torch.float64, device='cuda')
    tf, ge = ctypes.c_double(0.0), ctypes.c_double(0.0)
    torch.cuda.synchronize()
    check(lib().rc_measure_dmma_tflops(ptr(scratch), ctypes.byref(tf)), 'rc_measure_dmma_tflops')
    check(lib().rc_measure_exp_gexps(ptr(scratch), ctypes.byref(ge)), 'rc_measure_exp_gexps')
    gt = ctypes.c_double(0.0)
    check(lib().rc_measure_exp_tab_gexps(ptr(scratch), ctypes.byref(gt)), 'rc_measure_exp_tab_gexps')
    return {'dmma_tflops': tf.value, 'exp_gexps': ge.value, 'exp_tab_gexps': gt.value}


def debug_exp(x: torch.Tensor, form: int) -> torch.Tensor:
    """exp(x) through the device exp of the pairwise kernels (form 0: polynomial, 1: table) - a test hook."""
    x = x.contiguous()
    y = torch.empty_like(x)
    check(lib().rc_debug_exp(ptr(x), ptr(y), x.numel(), form, stream_ptr()), 'rc_debug_exp')
    return y


class gemm_profile:
    """Context manager around rc_profile_begin/end: ``with gemm_profile() as p: ...`` then p.ms, p.flops, p.launches, p.tflops."""

    def __enter__(self):
        global _PROFILING
        _PROFILING = True
        check(lib().rc_profile_begin(), 'rc_profile_begin')
        return self

    def __exit__(self, *exc):
        global _PROFILING
        _PROFILING = False
        ms, fl, cnt = ctypes.c_double(0.0), ctypes.c_double(0.0), ctypes.c_long(0)
        check(lib().rc_profile_end(ctypes.byref(ms), ctypes.byref(fl), ctypes.byref(cnt)), 'rc_profile_end')
        self.ms, self.flops, self.launches = ms.value, fl.value, cnt.value
        self.tflops = fl.value / (ms.value * 1e-3) * 1e-12 if ms.value > 0 else 0.0
        return False


def padded(n: int) -> int:
    return int(lib().rc_padded(int(n)))


def dev(a, device=None) -> torch.Tensor:
    """Host array / tensor -> contiguous CUDA float64 tensor on the current (or given) device."""
    if isinstance(a, torch.Tensor):
        return a.to(device=device or 'cuda', dtype=torch.float64).contiguous()
    import numpy as np
    return torch.as_tensor(np.require(a, dtype=np.float64, requirements=['C', 'W'])).to(device or 'cuda')


def workspace(nbytes: int, device=None) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 8), dtype=torch.uint8, device=device or 'cuda')


# ------------------------------------------------------------------------------------------------------------------
# thin, allocation-explicit wrappers (one per C entry point)
# ------------------------------------------------------------------------------------------------------------------
def gram(X, X2, ls, F=None, E=None, *, batch=1, lower_only=False, pad_to: Optional[int] = None, pad_identity=False, out=None) -> torch.Tensor:
    """-> (batch, rows_pad, cols_pad) tensor; ls is (batch*L, M); F, E are (batch, L, L) or None."""
    N, M = X.shape
    N2 = N if X2 is None else X2.shape[0]
    L = ls.shape[0] // batch
    rows = L * N if pad_to is None else max(pad_to, L * N)
    cols = L * N2 if pad_to is None else max(pad_to, L * N2)
    rows_pad, cols_pad = (rows + 63) // 64 * 64, (cols + 63) // 64 * 64
    if pad_to is not None:
        rows_pad, cols_pad = padded(rows), padded(cols)
    if out is None:
        out = torch.empty((batch, rows_pad, cols_pad), dtype=torch.float64, device=X.device)
    check(lib().rc_gram(ptr(X), N, ptr(X2), N2, M, ptr(ls), L, ptr(F), ptr(E), ptr(out), cols_pad, rows_pad * cols_pad, rows_pad, cols_pad,
                        int(lower_only), int(pad_identity), batch, stream_ptr()), 'rc_gram')
    return out


class Factorization:
    """Lower Cholesky factors of `batch` padded n_pad x n_pad matrices plus the workspace the solves need."""

    def __init__(self, A: torch.Tensor):
        assert A.dim() == 3 and A.shape[1] == A.shape[2]
        self.A, self.batch, self.n_pad = A, A.shape[0], A.shape[1]
        self.work = workspace(lib().rc_potrf_bufsize(self.n_pad, self.batch), A.device)
        self.info = torch.zeros(self.batch, dtype=torch.int32, device=A.device)
        check(lib().rc_potrf(ptr(A), self.n_pad, self.n_pad, self.n_pad * self.n_pad, self.batch, raw_ptr(self.work), raw_ptr(self.info),
                             stream_ptr()), 'rc_potrf')

    def raise_if_failed(self):
        info = self.info.cpu()
        if int(info.abs().max()) != 0:
            raise RomcommaB200Error(f'Cholesky decomposition was not successful: first non-positive pivot at (1-based) {info.tolist()}. '
                                    'The input might not be valid.')

    def logdet_half(self) -> torch.Tensor:
        out = torch.empty(self.batch, dtype=torch.float64, device=self.A.device)
        check(lib().rc_logdet(raw_ptr(self.work), self.n_pad, self.batch, ptr(out), stream_ptr()), 'rc_logdet')
        return out

    def trsv(self, v: torch.Tensor, transpose=False) -> torch.Tensor:
        """v: (batch, n_pad). Returns L^-1 v or L^-T v."""
        w = v.clone().contiguous()
        x = torch.empty_like(w)
        check(lib().rc_trsv(ptr(self.A), self.n_pad, self.n_pad, self.n_pad * self.n_pad, self.batch, raw_ptr(self.work), ptr(w), ptr(x),
                            self.n_pad, int(transpose), stream_ptr()), 'rc_trsv')
        return x

    SBINV_MIN_N, SBINV_MAX_RHS = 4096, int(os.environ.get('RC_TRSM_SBINV_MAX', '4096'))

    def trsm_fwd_(self, B: torch.Tensor) -> torch.Tensor:
        """B: (batch, n_pad, c_pad) with c_pad a multiple of 128; overwritten by L^-1 B.  One large factor and few columns: through the
        inverted diagonal super-blocks (rc_trsm_fwd_sbinv), which are formed on first use and kept with the factor."""
        if self.batch == 1 and self.n_pad >= self.SBINV_MIN_N and B.shape[2] <= self.SBINV_MAX_RHS and os.environ.get('RC_TRSM_SBINV', '1') != '0':
            if getattr(self, '_sbwork', None) is None:
                self._sbwork = workspace(lib().rc_trsm_sbinv_bufsize(self.n_pad, self.SBINV_MAX_RHS), self.A.device)
                check(lib().rc_trsm_sbinv_prepare(ptr(self.A), self.n_pad, self.n_pad, raw_ptr(self.work), raw_ptr(self._sbwork), stream_ptr()),
                      'rc_trsm_sbinv_prepare')
            check(lib().rc_trsm_fwd_sbinv(ptr(self.A), self.n_pad, self.n_pad, raw_ptr(self.work), raw_ptr(self._sbwork), self.SBINV_MAX_RHS, ptr(B),
                                          B.shape[2], B.shape[2], stream_ptr()), 'rc_trsm_fwd_sbinv')
            return B
        check(lib().rc_trsm_fwd(ptr(self.A), self.n_pad, self.n_pad, self.n_pad * self.n_pad, self.batch, raw_ptr(self.work), ptr(B), B.shape[2],
                                B.shape[2], B.shape[1] * B.shape[2], stream_ptr()), 'rc_trsm_fwd')
        return B

    def inverse_(self) -> torch.Tensor:
        """Overwrites the factor by L^-1 and returns K^-1 (lower 128-tiles valid)."""
        Kinv = torch.empty_like(self.A)
        self._sbwork = None                       # the factor is about to be overwritten: its inverted super-blocks go with it
        check(lib().rc_potri(ptr(self.A), self.n_pad, self.n_pad, self.n_pad * self.n_pad, self.batch, raw_ptr(self.work), ptr(Kinv), self.n_pad,
                             self.n_pad * self.n_pad, stream_ptr()), 'rc_potri')
        return Kinv

    def lower(self, n: int) -> torch.Tensor:
        """(batch, n, n) dense lower-triangular factor (zeros above the diagonal)."""
        return extract_lower(self.A, n)


def pad_identity(K: torch.Tensor) -> torch.Tensor:
    """(batch, n, n) -> (batch, n_pad, n_pad) with identity padding."""
    K = K if K.dim() == 3 else K[None]
    b, n, _ = K.shape
    n_pad = padded(n)
    out = torch.empty((b, n_pad, n_pad), dtype=torch.float64, device=K.device)
    check(lib().rc_pad_identity(ptr(K.contiguous()), n, n * n, ptr(out), n_pad, n_pad, n_pad * n_pad, b, stream_ptr()), 'rc_pad_identity')
    return out


def extract_lower(A: torch.Tensor, n: int, symmetrize=False) -> torch.Tensor:
    b, n_pad, _ = A.shape
    out = torch.empty((b, n, n), dtype=torch.float64, device=A.device)
    check(lib().rc_extract_lower(ptr(A), n_pad, n_pad * n_pad, ptr(out), n, n * n, b, int(symmetrize), stream_ptr()), 'rc_extract_lower')
    return out


def apply_variance_noise(Kunit: torch.Tensor, F: torch.Tensor, E: Optional[torch.Tensor], L: int, N: int, *, pad=False, lower_only=False):
    n = L * N
    n_pad = padded(n) if pad else n
    out = torch.empty((n_pad, n_pad), dtype=torch.float64, device=Kunit.device)
    check(lib().rc_apply_variance_noise(ptr(Kunit), Kunit.shape[-1], ptr(F), ptr(E), L, N, n_pad, ptr(out), n_pad, int(lower_only), stream_ptr()),
          'rc_apply_variance_noise')
    return out


class LmlGradPlan:
    """Pre-allocated workspace for repeated LML(+gradient) evaluations of one model shape (the optimiser's hot loop).

    The evaluation is a fixed sequence of a few hundred kernel launches on static buffers, so it can be replayed as one CUDA graph
    (``use_graph=True`` or ``ROMCOMMA_B200_GRAPHS=1``: the first call runs eagerly - it also performs the library's one-off attribute /
    scratch set-up - the second call is captured, later calls replay).  Measured on the B200: capture + instantiation costs ~150 ms once
    and saves 0.8 ms of a 120 ms evaluation at cfg3 (launch gaps are already hidden by the CPU running ahead), so it is OFF by default and
    only pays for long optimiser runs on small models.  Hyper-parameters are copied into the plan's own device buffers before every launch.
    """

    def __init__(self, X: torch.Tensor, Y: torch.Tensor, L: int, batch: int, flags: int, use_graph: Optional[bool] = None):
        self.X, self.Y = X.contiguous(), Y.contiguous()
        self.N, self.M = X.shape
        self.L, self.batch, self.flags = L, batch, flags
        assert Y.shape == (self.N, L * batch)
        self.stride = int(lib().rc_lml_grad_stride(L, self.M))
        self.nbytes = int(lib().rc_lml_grad_bufsize(self.N, self.M, L, batch, flags))
        self.work = workspace(self.nbytes, X.device)
        self.out = torch.empty((batch, self.stride), dtype=torch.float64, device=X.device)
        self.info = torch.zeros(batch, dtype=torch.int32, device=X.device)
        self._ls = torch.empty((batch * L, self.M), dtype=torch.float64, device=X.device)
        self._F = torch.empty((batch, L, L), dtype=torch.float64, device=X.device)
        self._E = torch.empty((batch, L, L), dtype=torch.float64, device=X.device)
        self._use_graph = (os.environ.get('ROMCOMMA_B200_GRAPHS', '0') == '1') if use_graph is None else bool(use_graph)
        self._calls, self._graph, self._graph_kunit = 0, None, None

    def _launch(self, Kunit: Optional[torch.Tensor]):
        check(lib().rc_lml_grad(ptr(self.X), ptr(self.Y), self.N, self.M, self.L, self.batch, ptr(self._ls), ptr(self._F), ptr(self._E), ptr(Kunit),
                                self.flags, raw_ptr(self.work), self.nbytes, ptr(self.out), raw_ptr(self.info), stream_ptr()), 'rc_lml_grad')

    def __call__(self, ls: torch.Tensor, F: torch.Tensor, E: torch.Tensor, Kunit: Optional[torch.Tensor] = None) -> torch.Tensor:
        """ls (batch*L, M), F, E (batch, L, L) device tensors. Returns the device tensor (batch, 1 + 2 L^2 + L M)."""
        self._ls.copy_(ls.reshape(self._ls.shape))
        self._F.copy_(F.reshape(self._F.shape))
        self._E.copy_(E.reshape(self._E.shape))
        self._calls += 1
        global _REPLAYED_LAUNCHES
        kunit_key = None if Kunit is None else Kunit.data_ptr()
        if _PROFILING:
            self._launch(Kunit)
        elif self._graph is not None and self._graph_kunit == kunit_key:
            self._graph.replay()
            _REPLAYED_LAUNCHES += self._graph_launches
        elif self._use_graph and self._calls >= 2 and self._graph is None:
            before = int(lib().rc_launch_count())
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                self._launch(Kunit)
            self._graph_launches = int(lib().rc_launch_count()) - before        # counted at capture, executed at every replay
            self._graph, self._graph_kunit, self._kunit_ref = graph, kunit_key, Kunit
            graph.replay()
        else:
            self._launch(Kunit)
        return self.out

    def unpack(self, out_host):
        """host (batch, stride) array -> list of dicts(lml, dF, dE, dls)."""
        L, M = self.L, self.M
        res = []
        for row in out_host:
            res.append({'lml': float(row[0]), 'dF': row[1:1 + L * L].reshape(L, L), 'dE': row[1 + L * L:1 + 2 * L * L].reshape(L, L),
                        'dls': row[1 + 2 * L * L:].reshape(L, M)})
        return res


class LmlGradMultiPlan:
    """LML(+gradient) of `batch` problems that do not share their data (folds, or any set of independent small GPs) in ONE call:
    rc_lml_grad_multi.  Xs[z] (N_z, M), Ys[z] (N_z, L) are packed once into (batch, Nmax, .) device buffers; every call takes the
    hyper-parameters of all problems: ls (batch*L, M), F, E (batch, L, L).  Result layout as LmlGradPlan."""

    def __init__(self, Xs: Sequence[torch.Tensor], Ys: Sequence[torch.Tensor], L: int, flags: int):
        self.batch, self.L, self.flags = len(Xs), L, flags
        self.M = Xs[0].shape[1]
        self.Ns = [int(x.shape[0]) for x in Xs]
        self.Nmax = max(self.Ns)
        device = Xs[0].device
        self.X = torch.zeros((self.batch, self.Nmax, self.M), dtype=torch.float64, device=device)
        self.Y = torch.zeros((self.batch, self.Nmax, L), dtype=torch.float64, device=device)
        for z, (x, y) in enumerate(zip(Xs, Ys)):
            assert x.shape[1] == self.M and tuple(y.shape) == (x.shape[0], L)
            self.X[z, :x.shape[0]].copy_(x)
            self.Y[z, :x.shape[0]].copy_(y)
        self.Ns_dev = torch.tensor(self.Ns, dtype=torch.int32, device=device)
        self.stride = int(lib().rc_lml_grad_stride(L, self.M))
        self.nbytes = int(lib().rc_lml_grad_bufsize(self.Nmax, self.M, L, self.batch, flags))
        self.work = workspace(self.nbytes, device)
        self.out = torch.empty((self.batch, self.stride), dtype=torch.float64, device=device)
        self.info = torch.zeros(self.batch, dtype=torch.int32, device=device)

    def __call__(self, ls: torch.Tensor, F: torch.Tensor, E: torch.Tensor) -> torch.Tensor:
        ls, F, E = ls.reshape(self.batch * self.L, self.M).contiguous(), F.reshape(self.batch, self.L, self.L).contiguous(), \
            E.reshape(self.batch, self.L, self.L).contiguous()
        self._hold = (ls, F, E)       # read by the kernels until the evaluation ends - possibly on another stream than the one they were made on
        check(lib().rc_lml_grad_multi(ptr(self.X), ptr(self.Y), raw_ptr(self.Ns_dev), self.Nmax, self.M, self.L, self.batch, ptr(ls), ptr(F), ptr(E),
                                      self.flags, raw_ptr(self.work), self.nbytes, ptr(self.out), raw_ptr(self.info), stream_ptr()), 'rc_lml_grad_multi')
        return self.out

    unpack = LmlGradPlan.unpack


def predict_reduce(A: torch.Tensor, a: torch.Tensor, L: int, nstar: int, kdiag: torch.Tensor, noise: Optional[torch.Tensor]):
    """A (batch, n_pad, c_pad) = L^-1 Kmn, a (batch, n_pad) = L^-1 y, kdiag/noise (batch, L)  ->  mean, var each (batch, nstar, L)."""
    b, n_pad, c_pad = A.shape
    parts = workspace(lib().rc_predict_bufsize(c_pad, b), A.device)
    mean = torch.empty((b, nstar, L), dtype=torch.float64, device=A.device)
    var = torch.empty_like(mean)
    check(lib().rc_predict_reduce(ptr(A), c_pad, n_pad * c_pad, ptr(a), n_pad, n_pad, c_pad, b, L, nstar, ptr(kdiag), ptr(noise), raw_ptr(parts),
                                  ptr(mean), ptr(var), stream_ptr()), 'rc_predict_reduce')
    return mean, var


def syrk_tn(A: torch.Tensor, alpha: float = 1.0, beta: float = 0.0, C: Optional[torch.Tensor] = None) -> torch.Tensor:
    """C = alpha * A^T A + beta * C for A (batch, n_pad, c_pad), both multiples of 128 -> (batch, c_pad, c_pad)."""
    b, n_pad, c_pad = A.shape
    if C is None:
        assert beta == 0.0
        C = torch.empty((b, c_pad, c_pad), dtype=torch.float64, device=A.device)
    check(lib().rc_syrk_tn(ptr(A), n_pad, c_pad, c_pad, n_pad * c_pad, b, float(alpha), float(beta), ptr(C), c_pad, c_pad * c_pad, stream_ptr()),
          'rc_syrk_tn')
    return C


def predict_gradient(X, xs, ls, variance, KinvY, fac: Factorization):
    """MOGP.predict_gradient of a variant GP (batch = L problems): X (N,M), xs (o,M), ls (L,M), variance (L,), KinvY (L,N) device tensors and the
    factorisation of the L noisy grams -> mean (o,L,M), var (o,o,L,M,M) on the device."""
    (N, M), o, L = X.shape, xs.shape[0], ls.shape[0]
    c_pad = padded(o * M)
    B = torch.zeros((L, fac.n_pad, c_pad), dtype=torch.float64, device=X.device)
    mean = torch.empty((o, L, M), dtype=torch.float64, device=X.device)
    check(lib().rc_predict_gradient_jacobian(ptr(X), N, M, ptr(xs), o, ptr(ls), ptr(variance), ptr(KinvY), L, ptr(B), c_pad, fac.n_pad * c_pad, ptr(mean),
                                             stream_ptr()), 'rc_predict_gradient_jacobian')
    fac.trsm_fwd_(B)
    C = syrk_tn(B, alpha=-1.0)
    var = torch.empty((o, o, L, M, M), dtype=torch.float64, device=X.device)
    check(lib().rc_predict_gradient_finish(ptr(C), c_pad, c_pad * c_pad, ptr(xs), o, M, ptr(ls), ptr(variance), L, ptr(var), stream_ptr()),
          'rc_predict_gradient_finish')
    return mean, var


def column_stats(data: torch.Tensor) -> torch.Tensor:
    """(5, C) rows mean, std (ddof = 1), rng, min, max of data (N, C): the rows of a fold's normalization.csv (data/storage.py:544-558)."""
    N, Cc = data.shape
    stats = torch.empty((5, Cc), dtype=torch.float64, device=data.device)
    check(lib().rc_column_stats(ptr(data.contiguous()), N, Cc, ptr(stats), stream_ptr()), 'rc_column_stats')
    return stats


def normalize(data: torch.Tensor, M: int, stats: torch.Tensor, margin: float = 1.0E-12, undo: bool = False) -> torch.Tensor:
    """Normalization.apply_to (or undo_from) of data (N, M + L) on the device with statistics in the column_stats layout (data/storage.py:469-503)."""
    data = data.contiguous()
    N, Cc = data.shape
    out = torch.empty_like(data)
    check(lib().rc_normalize(ptr(data), N, M, Cc, ptr(stats.contiguous()), margin, -1 if undo else 1, ptr(out), stream_ptr()), 'rc_normalize')
    return out


def test_metrics(truth: torch.Tensor, mean: torch.Tensor, sd: torch.Tensor):
    """GPR.test's numbers (gpr/models.py:235-272): -> reals (n, 2L) [Abs Error | Z Score], flags (n, L + 2), summary (3L + 2), all on the device."""
    n, L = truth.shape
    reals = torch.empty((n, 2 * L), dtype=torch.float64, device=truth.device)
    flags = torch.empty((n, L + 2), dtype=torch.float64, device=truth.device)
    summary = torch.empty(3 * L + 2, dtype=torch.float64, device=truth.device)
    check(lib().rc_test_metrics(ptr(truth.contiguous()), ptr(mean.contiguous()), ptr(sd.contiguous()), n, L, ptr(reals), ptr(flags), ptr(summary),
                                stream_ptr()), 'rc_test_metrics')
    return reals, flags, summary


def sobol_prepare(X, Lam, F, KinvY, is_F_diagonal: bool):
    N, M = X.shape
    L = Lam.shape[0]
    P = L if is_F_diagonal else L * L
    Phi = torch.empty((P, M), dtype=torch.float64, device=X.device)
    g0 = torch.empty((P, N), dtype=torch.float64, device=X.device)
    g0KY = torch.empty_like(g0)
    check(lib().rc_sobol_prepare(ptr(X), N, M, ptr(Lam), ptr(F), ptr(KinvY), L, int(is_F_diagonal), ptr(Phi), ptr(g0), ptr(g0KY), stream_ptr()),
          'rc_sobol_prepare')
    return Phi, g0, g0KY


def slice_mask(m0: int, m1: int) -> int:
    return ((1 << max(m1 - m0, 0)) - 1) << m0


def sobol_contract(X, Phi, c, L: int, is_F_diagonal: bool, masks: Sequence[int], parts: Optional[torch.Tensor] = None, part: int = 0,
                   nparts: int = 1) -> torch.Tensor:
    """-> V (len(masks), L, L) on the device; with nparts > 1 the partial sums over the row tiles ti % nparts == part."""
    N, M = X.shape
    P = Phi.shape[0]
    ns = len(masks)
    if parts is None:
        parts = workspace(lib().rc_sobol_bufsize(N, P, ns), X.device)
    V = torch.empty((ns, L, L), dtype=torch.float64, device=X.device)
    arr = (ctypes.c_ulonglong * ns)(*[int(m) for m in masks])
    check(lib().rc_sobol_contract_part(ptr(X), N, M, ptr(Phi), ptr(c), L, int(is_F_diagonal), ctypes.cast(arr, ctypes.c_void_p), ns, int(part),
                                       int(nparts), raw_ptr(parts), ptr(V), stream_ptr()), 'rc_sobol_contract_part')
    return V


def sobol_error(X, Lam, F, Phi, g0, g0KY, fac: Factorization, masks: Sequence[int], mixed: bool = False):
    """ClosedSobolWithError for a list of subsets (diagonal F): -> (V, W), each (len(masks), L, L) on the device; with ``mixed`` (is_T_partial=False)
    -> (V, W, WMm).  ``fac`` is the Cholesky factorisation of the GP's noisy gram: batch 1 (covariant, n = L*N) or L (variant, n = N)."""
    N, M = X.shape
    L = Lam.shape[0]
    ns = len(masks)
    bufsize = lib().rc_sobol_error_mixed_bufsize if mixed else lib().rc_sobol_error_bufsize
    nbytes = int(bufsize(N, M, L, ns, fac.n_pad, fac.batch))
    work = workspace(nbytes, X.device)
    V = torch.empty((ns, L, L), dtype=torch.float64, device=X.device)
    W = torch.empty_like(V)
    arr = (ctypes.c_ulonglong * ns)(*[int(m) for m in masks])
    common = (ptr(X), N, M, ptr(Lam), ptr(F), ptr(Phi), ptr(g0), ptr(g0KY), L, ptr(fac.A), fac.n_pad, fac.n_pad, fac.n_pad * fac.n_pad, fac.batch,
              raw_ptr(fac.work), ctypes.cast(arr, ctypes.c_void_p), ns, raw_ptr(work), nbytes, ptr(V), ptr(W))
    if mixed:
        WMm = torch.empty_like(V)
        check(lib().rc_sobol_error_mixed(*common, ptr(WMm), stream_ptr()), 'rc_sobol_error_mixed')
        return V, W, WMm
    check(lib().rc_sobol_error(*common, stream_ptr()), 'rc_sobol_error')
    return V, W
