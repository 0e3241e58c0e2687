"""Minimal driver for ncu: a few full LML+grad evaluations at cfg3 (no e2e / sobol / cpu legs)."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / 'rom-comma_b200'))
from romcomma import _capi as C, synthetic
w = synthetic.config('cfg3')
L = 4
flags = C.RC_GRAD_VARIANCE | (C.RC_GRAD_F_DIAGONAL if 'sel' in sys.argv[2:] else 0)
plan = C.LmlGradPlan(C.dev(w.X), C.dev(w.Y), L, 1, flags)
args = (C.dev(w.lengthscales), C.dev(w.F[None]), C.dev(w.E[None]))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
for _ in range(n):
    plan(*args)
torch.cuda.synchronize()
print('lml', plan.out[0, 0].item(), 'launches', C.launch_count())
