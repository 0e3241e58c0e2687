"""Minimal driver for ncu / timing: the 25-slice Sobol sweep at cfg3 (prepare + contract)."""
import sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / 'rom-comma_b200'))
from romcomma import _capi as C, synthetic
w = synthetic.config(sys.argv[1] if len(sys.argv) > 1 else 'cfg3')
N, M = w.X.shape
L = w.Y.shape[1]
dX, dLam, dF = C.dev(w.X), C.dev(w.lengthscales), C.dev(np.diag(w.F).copy())
KiY = torch.randn(L, N, dtype=torch.float64, device='cuda', generator=torch.Generator('cuda').manual_seed(7)) * 0.1
slices = [(m, m + 1) for m in range(M)] + [(0, m + 1) for m in range(M)] + [(m + 1, M) for m in range(M)] + [(0, M)]
masks = [C.slice_mask(*s) for s in slices]
parts = C.workspace(C.lib().rc_sobol_bufsize(N, L, len(masks)))
def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
Phi, g0, g0KY = C.sobol_prepare(dX, dLam, dF, KiY, True)
print('prepare ms', timed(lambda: C.sobol_prepare(dX, dLam, dF, KiY, True)))
print('contract ms', timed(lambda: C.sobol_contract(dX, Phi, g0KY, L, True, masks, parts)))
print('launches', C.launch_count())
