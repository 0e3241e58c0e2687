"""Where the time of the cfg2 fold loop goes: one batched evaluation of 10 folds (rc_lml_grad_multi) against 10 single ones, then
user.run.gpr with and without testing, lock-step on and off (ROMCOMMA_B200_LOCKSTEP)."""
import json, os, random, sys, tempfile, time
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / 'rom-comma_b200'))
from romcomma import _capi as C
from romcomma.user import functions, run, sample

def timed(fn, reps=10):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3

rng = np.random.default_rng(0)
M, flags = 10, C.RC_GRAD_VARIANCE | C.RC_GRAD_LENGTHSCALES
Ns = [1843, 1844] * 5
Xs = [C.dev(rng.normal(size=(n, M))) for n in Ns]
Ys = [C.dev(rng.normal(size=(n, 1))) for n in Ns]
ls, F, E = C.dev(np.full((10, M), 2.0)), C.dev(np.full((10, 1, 1), 1.5)), C.dev(np.full((10, 1, 1), 0.05))
multi = C.LmlGradMultiPlan(Xs, Ys, 1, flags)
rec = {'multi10_ms': timed(lambda: multi(ls, F, E).cpu())}
single = C.LmlGradPlan(Xs[0], Ys[0], 1, 1, flags)
rec['single_ms'] = timed(lambda: single(ls[:1], F[:1], E[:1]).cpu())
print(json.dumps(rec), flush=True)
del multi, single
for lock in ('1', '0'):
    os.environ['ROMCOMMA_B200_LOCKSTEP'] = lock
    for tested in (False, True):
        with tempfile.TemporaryDirectory() as tmp:
            np.random.seed(2); random.seed(2)
            fn = sample.Function(tmp, lambda N, M: sample.DOE.latin_hypercube(N, M, seed=2), functions.SOBOL_G.subVector('sobol_g', ['weak5_2']), N=2048, M=10,
                                 noise_variance=sample.GaussianNoise.Variance(1, 0.04, False, False), overwrite_existing=True)
            repo = fn.repo.into_K_folds(10)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            run.gpr('gpr', repo, is_read=False, is_covariant=False, is_isotropic=False, maxiter=50, is_tested=tested)
            torch.cuda.synchronize(); t1 = time.perf_counter()
            run.gsa('gpr', repo, is_covariant=False, is_isotropic=False)
            torch.cuda.synchronize(); t2 = time.perf_counter()
            print(json.dumps({'lockstep': lock, 'tested': tested, 'gpr_s': t1 - t0, 'gsa_s': t2 - t1}), flush=True)
