"""Run-to-run bitwise reproducibility and independence from uninitialised workspace contents of the LML+gradient path."""
import sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / 'rom-comma_b200'), str(ROOT / 'tests')]
from romcomma import _capi as C
from conftest import random_problem
C.lib()
bad = 0
for (N, M, L, batch) in [(60, 3, 1, 2), (120, 3, 2, 1), (300, 4, 3, 1), (700, 5, 1, 3), (1100, 6, 2, 1), (2048, 8, 4, 1)]:
    X, Y, ls, F, E = random_problem(N, M, L * batch, seed=N, full_E=False)
    flags = C.RC_GRAD_VARIANCE | C.RC_GRAD_LENGTHSCALES
    outs = []
    for rep in range(4):
        plan = C.LmlGradPlan(C.dev(X), C.dev(Y), L, batch, flags)
        fill = [float('nan'), 1e300, -1e300, 0.0][rep]
        plan.work.view(torch.float64)[: plan.work.numel() // 8].fill_(fill)       # poison the whole workspace
        if batch == 1:
            out = plan(C.dev(ls), C.dev(F[None]), C.dev(E[None]))
        else:
            out = plan(C.dev(ls), C.dev(np.diag(F).reshape(batch, 1, 1).copy()), C.dev(np.diag(E).reshape(batch, 1, 1).copy()))
        outs.append(out.cpu().numpy().copy())
        del plan
    same = all(np.array_equal(outs[0], o, equal_nan=True) for o in outs[1:])
    finite = np.isfinite(outs[0]).all()
    print(f'N={N} M={M} L={L} batch={batch}: bitwise identical across poisoned workspaces: {same}; finite: {finite}; lml {outs[0][:, 0]}')
    if not (same and finite):
        bad += 1
        for o in outs:
            print('   ', o[:, :6])
print('FAILED' if bad else 'OK')
