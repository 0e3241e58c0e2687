"""One rank's share of the cfg3 sweep on a multi-GPU run, timed on one GPU: rc_sobol_contract_part(part 0 of NPARTS) for the chunk length in RC_SOBOL_CHUNK."""
import json, os, sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / 'rom-comma_b200'))
from romcomma import _capi as C, synthetic
w = synthetic.config('cfg3')
(N, M), L = w.X.shape, w.Y.shape[1]
dX = C.dev(w.X)
KiY = torch.randn(L, N, dtype=torch.float64, device='cuda', generator=torch.Generator('cuda').manual_seed(7)) * 0.1
Phi, g0, g0KY = C.sobol_prepare(dX, C.dev(w.lengthscales), C.dev(np.diag(w.F).copy()), KiY, True)
slices = [(m, m + 1) for m in range(M)] + [(0, m + 1) for m in range(M)] + [(m + 1, M) for m in range(M)] + [(0, M)]
masks = [C.slice_mask(*s) for s in slices]
parts = C.workspace(C.lib().rc_sobol_bufsize(N, L, len(masks)))
rec = {'chunk': os.environ.get('RC_SOBOL_CHUNK', 'auto')}
for nparts in (1, 2, 4, 8):
    fn = lambda: C.sobol_contract(dX, Phi, g0KY, L, True, masks, parts, 0, nparts)
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): fn()
    b.record(); torch.cuda.synchronize()
    rec[f'ms_nparts{nparts}'] = a.elapsed_time(b) / 20
print(json.dumps(rec))
