import sys, time
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / 'rom-comma_b200'))
from romcomma import _capi as C
rng = np.random.default_rng(0)
for n in (128, 256, 1024):
    A = rng.normal(size=(n, n)); K = A @ A.T / n + np.eye(n)
    Kp = C.pad_identity(C.dev(K))
    work = C.workspace(C.lib().rc_potrf_bufsize(n, 1)); info = torch.zeros(1, dtype=torch.int32, device='cuda')
    def run():
        Kc = Kp.clone()
        C.check(C.lib().rc_potrf(C.ptr(Kc), n, n, n * n, 1, C.raw_ptr(work), C.raw_ptr(info), C.stream_ptr()), 'potrf')
    for _ in range(3): run()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): run()
    e1.record(); torch.cuda.synchronize()
    print(f'n={n}: potrf {e0.elapsed_time(e1)/50*1e3:.1f} us per call (incl clone)')
