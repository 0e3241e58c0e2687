"""Minimal driver for ncu / timing: the 25-slice Sobol sweep WITH errors at a configuration (rc_sobol_error on a resident factor)."""
import sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / 'rom-comma_b200'))
from romcomma import _capi as C, synthetic
w = synthetic.config(sys.argv[1] if len(sys.argv) > 1 else 'cfg3')
N, M = w.X.shape
L = w.Y.shape[1]
n = L * N
dX, dLam, dFdiag = C.dev(w.X), C.dev(w.lengthscales), C.dev(np.diag(w.F).copy())
K = C.gram(dX, None, dLam, C.dev(w.F[None]), C.dev(w.E[None]), lower_only=True, pad_to=n, pad_identity=True)
fac = C.Factorization(K)
fac.raise_if_failed()
KiY = torch.randn(L, N, dtype=torch.float64, device='cuda', generator=torch.Generator('cuda').manual_seed(7)) * 0.1
Phi, g0, g0KY = C.sobol_prepare(dX, dLam, dFdiag, KiY, True)
slices = [(m, m + 1) for m in range(M)] + [(0, m + 1) for m in range(M)] + [(m + 1, M) for m in range(M)] + [(0, M)]
masks = [C.slice_mask(*s) for s in slices]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
C.sobol_error(dX, dLam, dFdiag, Phi, g0, g0KY, fac, masks)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(reps):
    V, W = C.sobol_error(dX, dLam, dFdiag, Phi, g0, g0KY, fac, masks)
b.record()
torch.cuda.synchronize()
print('sobol_error ms', a.elapsed_time(b) / reps, 'checksum', float(V.sum()), float(W.sum()), 'launches', C.launch_count())
