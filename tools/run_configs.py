"""Time one LML(+gradient) evaluation and one Sobol sweep at each BASELINE.json configuration (device-resident, after one warm-up)."""
import json, sys, time
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / 'rom-comma_b200'))
from romcomma import _capi as C, synthetic

def timed(fn, reps=2):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

out = []
for name in sys.argv[1:] or ['cfg1', 'cfg2', 'cfg3', 'cfg4', 'cfg5']:
    w = synthetic.config(name)
    N, M = w.X.shape
    L = w.Y.shape[1]
    dX, dY = C.dev(w.X), C.dev(w.Y)
    rec = {'config': name, 'N': N, 'M': M, 'L': L}
    if L == 1 or name == 'cfg2':                       # L = 1 can never be covariant (quirk Q4): gpflow GPR per output
        plan = C.LmlGradPlan(dX, dY, 1, L, C.RC_GRAD_VARIANCE | C.RC_GRAD_LENGTHSCALES)
        args = (C.dev(w.lengthscales), C.dev(np.diag(w.F).reshape(L, 1, 1).copy()), C.dev(np.diag(w.E).reshape(L, 1, 1).copy()))
        rec['path'] = 'variant (L independent N x N GPs, gradients incl. lengthscales)'
    else:
        plan = C.LmlGradPlan(dX, dY, L, 1, C.RC_GRAD_VARIANCE | C.RC_GRAD_F_DIAGONAL)
        args = (C.dev(w.lengthscales), C.dev(w.F[None]), C.dev(w.E[None]))
        rec['path'] = 'covariant (default trainables)'
    rec['lml_grad_ms'] = timed(lambda: plan(*args))
    rec['lml'] = plan.out[:, 0].cpu().numpy().tolist()
    rec['info'] = plan.info.cpu().numpy().tolist()
    n = (L if rec['path'].startswith('cov') else 1) * N
    rec['n'] = n
    rec['tflops_on_n3'] = (L if not rec['path'].startswith('cov') else 1) * float(n) ** 3 / (rec['lml_grad_ms'] * 1e-3) * 1e-12
    del plan
    torch.cuda.empty_cache()
    KiY = torch.randn(L, N, dtype=torch.float64, device='cuda', generator=torch.Generator('cuda').manual_seed(7)) * 0.1
    Phi, g0, g0KY = C.sobol_prepare(dX, C.dev(w.lengthscales), C.dev(np.diag(w.F).copy()), KiY, True)
    slices = [(m, m + 1) for m in range(M)] + [(0, m + 1) for m in range(M)] + [(m + 1, M) for m in range(M)] + [(0, M)]
    masks = [C.slice_mask(*s) for s in slices]
    parts = C.workspace(C.lib().rc_sobol_bufsize(N, L, len(masks)))
    rec['sobol_sweep_ms'] = timed(lambda: C.sobol_contract(dX, Phi, g0KY, L, True, masks, parts), reps=3)
    rec['sobol_slices'] = len(masks)
    print(json.dumps(rec), flush=True)
