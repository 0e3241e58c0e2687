"""Where does the end-to-end (host-buffer) LML+grad step spend its time?  python tools/e2e_profile.py"""
import cProfile, pstats, sys, time
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / 'rom-comma_b200')]
from romcomma import synthetic, _capi as C
from romcomma.gpf import kernels, models
w = synthetic.config('cfg3')
Xh, Yh = torch.as_tensor(w.X).pin_memory(), torch.as_tensor(w.Y).pin_memory()
def step():
    model = models.MOGPR((Xh, Yh), kernels.RBF(w.F, w.lengthscales), noise_variance=w.E)
    return model._loss_and_grad(model.trainable_variables)
for _ in range(3): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5): step()
torch.cuda.synchronize()
print('e2e ms/step', (time.perf_counter() - t0) / 5 * 1e3)
pr = cProfile.Profile(); pr.enable()
for _ in range(5): step()
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(25)
