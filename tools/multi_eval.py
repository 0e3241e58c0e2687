"""One batched LML+gradient evaluation of 10 cfg2-sized folds (rc_lml_grad_multi) - for `ncu --metrics gpu__time_duration.sum` launch lists."""
import sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / 'rom-comma_b200'))
from romcomma import _capi as C
rng = np.random.default_rng(0)
M, flags = 10, C.RC_GRAD_VARIANCE | C.RC_GRAD_LENGTHSCALES
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 10
Ns = ([1843, 1844] * 8)[:batch]
Xs = [C.dev(rng.normal(size=(n, M))) for n in Ns]
Ys = [C.dev(rng.normal(size=(n, 1))) for n in Ns]
ls, F, E = C.dev(np.full((batch, M), 2.0)), C.dev(np.full((batch, 1, 1), 1.5)), C.dev(np.full((batch, 1, 1), 0.05))
multi = C.LmlGradMultiPlan(Xs, Ys, 1, flags)
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 2):
    multi(ls, F, E)
    torch.cuda.synchronize()
print(multi.out[:, 0].cpu().numpy())
