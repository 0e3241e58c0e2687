"""TRSM (B <- L^-1 B) and prediction at a BASELINE configuration: ms and TFLOP/s on the n^2 * nrhs flops of the solve."""
import json, sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / 'rom-comma_b200'))
from romcomma import _capi as C, synthetic, gf_compat as gf
cfg = sys.argv[1] if len(sys.argv) > 1 else 'cfg3'
w = synthetic.config(cfg)
N, M = w.X.shape
L = w.Y.shape[1]
n = L * N
dX, dY, dls = C.dev(w.X), C.dev(w.Y), C.dev(w.lengthscales)
K = C.gram(dX, None, dls, C.dev(w.F[None]), C.dev(w.E[None]), lower_only=True, pad_to=n, pad_identity=True)
fac = C.Factorization(K)
fac.raise_if_failed()
n_pad = fac.n_pad

def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

g = torch.Generator('cuda').manual_seed(3)
import os
for sb, nrhs in [(sb, nrhs) for nrhs in (512, 2048, 8192) for sb in (1, 2, 4, 8)]:
    os.environ['RC_TRSM_SB'] = str(sb)
    B0 = torch.randn(1, n_pad, nrhs, dtype=torch.float64, device='cuda', generator=g)
    B = B0.clone()
    ms = timed(lambda: fac.trsm_fwd_(B.copy_(B0)))
    ms_copy = timed(lambda: B.copy_(B0))
    # residual check against the factor: L (L^-1 B) = B on a few columns
    Lm = C.extract_lower(fac.A, n_pad)[0]
    X = fac.trsm_fwd_(B.copy_(B0))[0, :, :8]
    res = float((Lm @ X - B0[0, :, :8]).abs().max() / B0[0, :, :8].abs().max())
    del Lm
    print(json.dumps({'cfg': cfg, 'op': 'trsm_fwd', 'sb': sb, 'n': n_pad, 'nrhs': nrhs, 'ms': ms - ms_copy, 'tflops': float(n_pad) ** 2 * nrhs / ((ms - ms_copy) * 1e-3) * 1e-12,
                      'rel_residual': res}), flush=True)
    del B, B0
    torch.cuda.empty_cache()
os.environ['RC_TRSM_SB'] = '4'
rng = np.random.default_rng(5)
for nstar in (256, 2048):
    Xn = C.dev(rng.standard_normal((nstar, M)))
    ms = timed(lambda: gf.predict_core(dX, dY, dls, w.F[None], w.E[None], Xn, L, 1, True), reps=2)
    print(json.dumps({'cfg': cfg, 'op': 'predict (gram + potrf + trsv + cross gram + trsm + reduce)', 'n': n_pad, 'nstar': nstar, 'nrhs': L * nstar, 'ms': ms}), flush=True)
