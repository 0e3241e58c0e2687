"""rc_trsm_fwd (block substitution) against rc_trsm_fwd_sbinv (inverted diagonal super-blocks) at a BASELINE configuration, and MOGP-style prediction."""
import json, os, sys
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
for nrhs in [int(v) for v in os.environ.get('NRHS', '128,512,1024,2048').split(',')]:
    B0 = torch.randn(1, n_pad, nrhs, dtype=torch.float64, device='cuda', generator=g)
    B = B0.clone()
    ms_copy = timed(lambda: B.copy_(B0))
    rec = {'cfg': cfg, 'n': n_pad, 'nrhs': nrhs}
    for mode in ('1', '0'):
        os.environ['RC_TRSM_SBINV'] = mode
        ms = timed(lambda: fac.trsm_fwd_(B.copy_(B0))) - ms_copy
        rec['sbinv_ms' if mode == '1' else 'block_ms'] = ms
        rec['sbinv_tflops' if mode == '1' else 'block_tflops'] = float(n_pad) ** 2 * nrhs / (ms * 1e-3) * 1e-12
    print(json.dumps(rec), flush=True)
    del B, B0
rng = np.random.default_rng(5)
Xn = C.dev(rng.standard_normal((256, M)))
for mode in ('1', '0'):
    os.environ['RC_TRSM_SBINV'] = mode
    ms = timed(lambda: gf.predict_core(dX, dY, dls, w.F[None], w.E[None], Xn, L, 1, True, fac=fac), reps=3)
    print(json.dumps({'cfg': cfg, 'op': 'predict 256 points on the resident factor', 'sbinv': mode, 'ms': ms}), flush=True)
