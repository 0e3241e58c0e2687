"""Step time of the cfg LML+grad evaluation for several panel counts of the overlapped potrf+trtri (RC_OVERLAP_PANELS; 0 = one stream)."""
import os, sys, json
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / 'rom-comma_b200'))
from romcomma import _capi as C, synthetic
cfg = sys.argv[1] if len(sys.argv) > 1 else 'cfg3'
w = synthetic.config(cfg)
N, M = w.X.shape
L = w.Y.shape[1]
dX, dY, dls, dF, dE = C.dev(w.X), C.dev(w.Y), C.dev(w.lengthscales), C.dev(w.F[None]), C.dev(w.E[None])
plan = C.LmlGradPlan(dX, dY, L, 1, C.RC_GRAD_VARIANCE | C.RC_GRAD_F_DIAGONAL)
ref = None
for panels in [int(a) for a in (sys.argv[2:] or ['0', '2', '4', '8', '12', '16'])]:
    os.environ['RC_OVERLAP_PANELS'] = str(panels)
    for _ in range(2):
        plan(dls, dF, dE)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    reps = 5
    for _ in range(reps):
        plan(dls, dF, dE)
    b.record()
    torch.cuda.synchronize()
    out = plan.out.cpu().numpy()[0]
    ref = out if ref is None else ref
    print(json.dumps({'cfg': cfg, 'panels': panels, 'ms': a.elapsed_time(b) / reps, 'lml': float(out[0]), 'info': int(plan.info.cpu()[0]),
                      'max_rel_diff_vs_first': float(np.max(np.abs(out - ref) / (np.abs(ref) + 1e-300)))}), flush=True)
