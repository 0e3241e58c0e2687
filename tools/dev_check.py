"""Developer check: raw C-ABI vs the numpy oracle on a GPU box (not part of the test-suite)."""
import sys, time
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / 'rom-comma_b200'))
from romcomma import _capi as C
from oracle import gp, sobol

def close(a, b, name, rtol=1e-8, atol=1e-10):
    a, b = np.asarray(a), np.asarray(b)
    err = np.max(np.abs(a - b) / (atol + rtol * np.abs(b)))
    print(f'{name:28s} max|err|/(atol+rtol|ref|) = {err:.3e}  {"OK" if err <= 1 else "FAIL"}', flush=True)
    return err <= 1

def problem(N, M, L, seed=0, fullF=False):
    rng = np.random.default_rng(seed)
    X = rng.normal(size=(N, M)); ls = rng.uniform(0.5, 3.0, (L, M))
    W = rng.normal(size=(M, L)); Y = np.sin(X @ W) + 0.1 * rng.normal(size=(N, L)); Y = (Y - Y.mean(0)) / Y.std(0)
    F = np.diag(rng.uniform(0.5, 2.0, L))
    if fullF:
        A = rng.normal(size=(L, L)); F = A @ A.T / L + np.eye(L)
    B = rng.normal(size=(L, L)); E = 0.01 * (np.eye(L) + B @ B.T / L)
    return X, Y, ls, F, E

ok = True
for (N, M, L, fullF) in [(50, 3, 2, True), (200, 5, 3, False), (300, 4, 1, False)]:
    X, Y, ls, F, E = problem(N, M, L, fullF=fullF)
    dX, dY, dls, dF, dE = (C.dev(a) for a in (X, Y, ls, F[None], E[None]))
    n = L * N
    K = C.gram(dX, None, dls, dF, dE)[0, :n, :n].cpu().numpy()
    ok &= close(K, gp.add_noise_mo(gp.gram_mo(X, None, ls, F), E), f'gram N={N} L={L}')
    Kp = C.gram(dX, None, dls, dF, dE, pad_to=n, pad_identity=True, lower_only=True)
    fac = C.Factorization(Kp); fac.raise_if_failed()
    Lc = fac.lower(n)[0].cpu().numpy()
    ok &= close(Lc, np.linalg.cholesky(gp.add_noise_mo(gp.gram_mo(X, None, ls, F), E)), 'potrf')
    Kinv = C.extract_lower(fac.inverse_(), n, symmetrize=True)[0].cpu().numpy()
    ok &= close(Kinv, np.linalg.inv(gp.add_noise_mo(gp.gram_mo(X, None, ls, F), E)), 'potri', rtol=1e-7, atol=1e-8)
    plan = C.LmlGradPlan(dX, dY, L, 1, C.RC_GRAD_VARIANCE | C.RC_GRAD_LENGTHSCALES)
    res = plan.unpack(plan(dls, dF, dE).cpu().numpy())[0]
    ref = gp.lml_grad_mo(X, Y, ls, F, E)
    for k in ('lml', 'dF', 'dE', 'dls'):
        ok &= close(res[k], ref[k], f'lml_grad {k}')
    # variant: batch over outputs
    var = np.diag(F).copy(); noise = np.diag(E).copy()
    planv = C.LmlGradPlan(dX, dY, 1, L, C.RC_GRAD_VARIANCE | C.RC_GRAD_LENGTHSCALES)
    resv = planv.unpack(planv(dls, C.dev(var.reshape(L, 1, 1)), C.dev(noise.reshape(L, 1, 1))).cpu().numpy())
    for l in range(L):
        r = gp.lml_grad_rbf(X, Y[:, l], ls[l], var[l], noise[l])
        ok &= close(resv[l]['lml'], r['lml'], f'variant lml l={l}')
        ok &= close(resv[l]['dF'][0, 0], r['dvariance'], f'variant dvar l={l}')
        ok &= close(resv[l]['dE'][0, 0], r['dnoise'], f'variant dnoise l={l}')
        ok &= close(resv[l]['dls'][0], r['dls'], f'variant dls l={l}')
    # sobol
    Ed = np.diag(np.diag(E))
    KiY = gp.k_inv_y_mo(X, Y, ls, F, Ed)
    for diag in (True, False):
        cal = sobol.ClosedSobol(X, ls, F, KiY, diag)
        Fin = np.diag(F).copy() if diag else F
        Phi, g0, g0KY = C.sobol_prepare(dX, dls, C.dev(Fin), C.dev(KiY.reshape(L, N)), diag)
        ok &= close(g0KY.cpu().numpy().reshape(cal.g0KY.shape), cal.g0KY, f'sobol g0KY diag={diag}')
        slices = [(0, M), (0, 1), (1, 2), (M - 1, M), (0, 2), (1, M), (M, M)]
        V = C.sobol_contract(dX, Phi, g0KY, L, diag, [C.slice_mask(*s) for s in slices]).cpu().numpy()
        for i, s in enumerate(slices):
            ok &= close(V[i], cal._V(*s), f'sobol V{s} diag={diag}')
print('ALL OK' if ok else 'SOME FAILED')

if '--big' in sys.argv:
    N, M, L = 4096, 8, 4
    X, Y, ls, F, E = problem(N, M, L, seed=3)
    dX, dY, dls, dF, dE = (C.dev(a) for a in (X, Y, ls, F[None], E[None]))
    plan = C.LmlGradPlan(dX, dY, L, 1, C.RC_GRAD_VARIANCE)
    for it in range(3):
        torch.cuda.synchronize(); t0 = time.time()
        out = plan(dls, dF, dE); torch.cuda.synchronize(); t1 = time.time()
        print(f'lml+grad n={L*N}: {1e3*(t1-t0):.1f} ms  lml={out[0,0].item():.6f} info={plan.info.tolist()}', flush=True)
    # stage timings
    n = L * N
    ev = lambda: torch.cuda.Event(enable_timing=True)
    def timed(f, name, flops=None):
        torch.cuda.synchronize(); a, b = ev(), ev(); a.record(); r = f(); b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b); print(f'  {name:12s} {ms:8.2f} ms' + (f'  {flops/ms*1e-9:6.2f} TF' if flops else ''), flush=True); return r
    for it in range(2):
        Kp = timed(lambda: C.gram(dX, None, dls, dF, dE, pad_to=n, pad_identity=True, lower_only=True), 'gram')
        fac = timed(lambda: C.Factorization(Kp), 'potrf', n**3 / 3)
        Kinv = timed(lambda: fac.inverse_(), 'potri', 2 * n**3 / 3)
    Ed = np.diag(np.diag(E)); 
    KiY = torch.randn(L, N, dtype=torch.float64, device='cuda')
    Phi, g0, g0KY = C.sobol_prepare(dX, dls, C.dev(np.diag(F).copy()), KiY, True)
    masks = [C.slice_mask(m, m + 1) for m in range(M)] + [C.slice_mask(0, m + 1) for m in range(M)] + [C.slice_mask(m + 1, M) for m in range(M)]
    for it in range(2):
        timed(lambda: C.sobol_contract(dX, Phi, g0KY, L, True, masks), 'sobol sweep')
