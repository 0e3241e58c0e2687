"""Sweep (cfg3, 25 slices) and lattice (cfg5, 8 blocks) kernel times + the sweep's result for the library named by ROMCOMMA_B200_LIB."""
import json, os, sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / 'rom-comma_b200'))
from romcomma import _capi as C, synthetic

def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
rec = {'lib': os.environ.get('ROMCOMMA_B200_LIB', 'default')}
for name in ('cfg3', 'cfg5'):
    w = synthetic.config(name)
    (N, M), L = w.X.shape, w.Y.shape[1]
    dX = C.dev(w.X)
    KiY = torch.randn(L, N, dtype=torch.float64, device='cuda', generator=torch.Generator('cuda').manual_seed(7)) * 0.1
    Phi, g0, g0KY = C.sobol_prepare(dX, C.dev(w.lengthscales), C.dev(np.diag(w.F).copy()), KiY, True)
    if name == 'cfg3':
        slices = [(m, m + 1) for m in range(M)] + [(0, m + 1) for m in range(M)] + [(m + 1, M) for m in range(M)] + [(0, M)]
        masks = [C.slice_mask(*s) for s in slices]
        rec['sweep_ms'] = timed(lambda: C.sobol_contract(dX, Phi, g0KY, L, True, masks))
        rec['sweep_V'] = C.sobol_contract(dX, Phi, g0KY, L, True, masks).cpu().numpy().reshape(-1)[:6].tolist()
        rec['gram_ms'] = timed(lambda: C.gram(dX, None, C.dev(w.lengthscales), C.dev(w.F[None]), C.dev(w.E[None]), lower_only=True, pad_to=L * N, pad_identity=True))
    else:
        masks = list(range(16 * 64, 24 * 64))
        rec['lattice8_ms'] = timed(lambda: C.sobol_contract(dX, Phi, g0KY, L, True, masks), reps=3)
        rec['lattice_V'] = C.sobol_contract(dX, Phi, g0KY, L, True, masks).cpu().numpy().reshape(-1)[:3].tolist()
print(json.dumps(rec))
