"""The structured-slice sweep (3M + 1 slices) at N = 4096, L = 4 for a range of M: one rc_sobol_contract call each (register-form sweep kernel)."""
import sys, json
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, '/root/repo/rom-comma_b200')
from romcomma import _capi as C, synthetic
def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
rec = {}
w = synthetic.config('cfg3')
N, L = 4096, 4
rng = np.random.default_rng(1)
for M in (3, 4, 5, 6, 7, 8, 9, 10, 12, 16, 20):
    X = rng.standard_normal((N, M)); ls = rng.uniform(0.5, 3.0, (L, M))
    dX = C.dev(X)
    KiY = torch.randn(L, N, dtype=torch.float64, device='cuda', generator=torch.Generator('cuda').manual_seed(7)) * 0.1
    Phi, g0, g0KY = C.sobol_prepare(dX, C.dev(ls), C.dev(np.ones(L)), KiY, True)
    slices = [(m, m + 1) for m in range(M)] + [(0, m + 1) for m in range(M)] + [(m + 1, M) for m in range(M)] + [(0, M)]
    masks = [C.slice_mask(*s) for s in slices]
    rec[f'M{M}'] = round(timed(lambda: C.sobol_contract(dX, Phi, g0KY, L, True, masks)), 4)
print(json.dumps(rec))
