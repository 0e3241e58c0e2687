"""Two LML+gradient evaluations of cfg3 (default trainables) with every kernel on ONE stream (RC_NO_OVERLAP) - the command behind the ncu launch
list profiles/rNN_launches_cfg3_eval.md.  `python tools/eval_cfg3.py [config] [lookahead]`: with a second argument the look-ahead path runs."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / 'rom-comma_b200'))
from romcomma import _capi as C, synthetic
w = synthetic.config(sys.argv[1] if len(sys.argv) > 1 else 'cfg3')
L = w.Y.shape[1]
flags = C.RC_GRAD_VARIANCE | C.RC_GRAD_F_DIAGONAL | (0 if len(sys.argv) > 2 else C.RC_NO_OVERLAP)
plan = C.LmlGradPlan(C.dev(w.X), C.dev(w.Y), L, 1, flags)
args = (C.dev(w.lengthscales), C.dev(w.F[None]), C.dev(w.E[None]))
for _ in range(2):
    plan(*args)
    torch.cuda.synchronize()
print('lml', float(plan.out[0, 0]), 'info', int(plan.info[0]), 'launches', C.launch_count())
