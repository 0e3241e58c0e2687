"""A matrix of operation x shape timings at cfg3 sizes, to catch routes that fall off the fast paths: rc_lml_grad per flag combination (covariant and
variant), predictions per number of test points, the error sweep per M."""
import json, sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / 'rom-comma_b200'))
from romcomma import _capi as C, synthetic, gf_compat as gf

def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return round(a.elapsed_time(b) / reps, 3)

w = synthetic.config('cfg3')
(N, M), L = w.X.shape, w.Y.shape[1]
dX, dY = C.dev(w.X), C.dev(w.Y)
rec = {}
cov = (C.dev(w.lengthscales), C.dev(w.F[None]), C.dev(w.E[None]))
for name, flags in (('lml_only', C.RC_GRAD_NONE), ('default_trainables', C.RC_GRAD_VARIANCE | C.RC_GRAD_F_DIAGONAL), ('variance_full_F', C.RC_GRAD_VARIANCE),
                    ('all_trainables', C.RC_GRAD_VARIANCE | C.RC_GRAD_LENGTHSCALES)):
    plan = C.LmlGradPlan(dX, dY, L, 1, flags)
    rec[f'covariant_{name}_ms'] = timed(lambda: plan(*cov))
    del plan
    torch.cuda.empty_cache()
var = (C.dev(w.lengthscales), C.dev(np.diag(w.F).reshape(L, 1, 1).copy()), C.dev(np.diag(w.E).reshape(L, 1, 1).copy()))
plan = C.LmlGradPlan(dX, dY, 1, L, C.RC_GRAD_VARIANCE | C.RC_GRAD_LENGTHSCALES)
rec['variant_L_gps_all_trainables_ms'] = timed(lambda: plan(*var))
del plan
torch.cuda.empty_cache()
K = C.gram(dX, None, cov[0], cov[1], cov[2], lower_only=True, pad_to=L * N, pad_identity=True)
fac = C.Factorization(K)
rng = np.random.default_rng(5)
for nstar in (16, 256, 1024, 2048):
    Xn = C.dev(rng.standard_normal((nstar, M)))
    rec[f'predict_{nstar}_ms'] = timed(lambda: gf.predict_core(dX, dY, cov[0], w.F[None], w.E[None], Xn, L, 1, True, fac=fac))
del fac, K
torch.cuda.empty_cache()
for Mx in (3, 5, 8, 12, 16):
    X = rng.standard_normal((N, Mx)); ls = rng.uniform(0.5, 3.0, (L, Mx))
    dXm, dls, dF = C.dev(X), C.dev(ls), C.dev(np.ones(L))
    KiY = torch.randn(L, N, dtype=torch.float64, device='cuda', generator=torch.Generator('cuda').manual_seed(7)) * 0.1
    Km = C.gram(dXm, None, dls, C.dev(np.eye(L)[None].copy()), C.dev(0.01 * np.eye(L)[None]), lower_only=True, pad_to=L * N, pad_identity=True)
    f = C.Factorization(Km)
    Phi, g0, g0KY = C.sobol_prepare(dXm, dls, dF, KiY, True)
    slices = [(m, m + 1) for m in range(Mx)] + [(0, m + 1) for m in range(Mx)] + [(m + 1, Mx) for m in range(Mx)] + [(0, Mx)]
    masks = [C.slice_mask(*s) for s in slices]
    rec[f'sobol_error_M{Mx}_ms'] = timed(lambda: C.sobol_error(dXm, dls, dF, Phi, g0, g0KY, f, masks), reps=2)
    rec[f'sobol_error_mixed_M{Mx}_ms'] = timed(lambda: C.sobol_error(dXm, dls, dF, Phi, g0, g0KY, f, masks, mixed=True), reps=2)
    del f, Km
    torch.cuda.empty_cache()
print(json.dumps(rec))
