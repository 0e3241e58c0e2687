"""Minimal driver for ncu: the lattice-form all-subsets kernel on cfg5 data, two blocks of the subset lattice (128 subsets) per call."""
import sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / 'rom-comma_b200'))
from romcomma import _capi as C, synthetic
w = synthetic.config(sys.argv[1] if len(sys.argv) > 1 else 'cfg5')
(N, M), L = w.X.shape, w.Y.shape[1]
dX = C.dev(w.X)
KiY = torch.randn(L, N, dtype=torch.float64, device='cuda', generator=torch.Generator('cuda').manual_seed(7)) * 0.1
Phi, g0, g0KY = C.sobol_prepare(dX, C.dev(w.lengthscales), C.dev(np.diag(w.F).copy()), KiY, True)
masks = list(range(21 * 64, 23 * 64))          # blocks 21 and 22: three of the six high inputs each
for _ in range(2):
    V = C.sobol_contract(dX, Phi, g0KY, L, True, masks)
    torch.cuda.synchronize()
print('checksum', float(V.sum()))
