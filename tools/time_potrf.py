"""rc_potrf alone at the benchmark sizes: `python tools/time_potrf.py [cfg3 cfg5 cfg4]`.  One JSON line per configuration: milliseconds and
TFLOP/s (n^3/3) of the factorisation with the environment's settings (RC_POTRF_LOOKAHEAD=0|1, RC_POTRF_YIELD, RC_POTRF_LA_TAIL, RC_POTRF_T4/T8),
a checksum of the factor (identical bits with and without look-ahead) and the LML+gradient evaluation time with the same settings."""
import hashlib, json, os, sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / 'rom-comma_b200'))
from romcomma import _capi as C, synthetic

lib = C.lib()
for name in sys.argv[1:] or ['cfg3']:
    w = synthetic.config(name)
    (N, M), L = w.X.shape, w.Y.shape[1]
    n = C.padded(L * N)
    dX, dY, dls, dF, dE = C.dev(w.X), C.dev(w.Y), C.dev(w.lengthscales), C.dev(w.F[None]), C.dev(w.E[None])
    K = torch.empty((1, n, n), dtype=torch.float64, device='cuda')
    work = C.workspace(lib.rc_potrf_bufsize(n, 1))
    info = torch.zeros(1, dtype=torch.int32, device='cuda')
    times = []
    for rep in range(4):
        C.gram(dX, None, dls, dF, dE, pad_to=L * N, pad_identity=True, lower_only=True, out=K)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        C.check(lib.rc_potrf(C.ptr(K), n, n, n * n, 1, C.raw_ptr(work), C.raw_ptr(info), C.stream_ptr()), 'rc_potrf')
        b.record()
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b))
    assert int(info.cpu()[0]) == 0
    digest = hashlib.sha256(torch.tril(K[0]).cpu().numpy().tobytes()).hexdigest()[:16]
    del K
    torch.cuda.empty_cache()
    plan = C.LmlGradPlan(dX, dY, L, 1, C.RC_GRAD_VARIANCE | C.RC_GRAD_F_DIAGONAL)
    ev = []
    for rep in range(4):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        plan(dls, dF, dE)
        b.record()
        torch.cuda.synchronize()
        ev.append(a.elapsed_time(b))
    out = plan.out.cpu().numpy()
    env = {k: v for k, v in os.environ.items() if k.startswith('RC_')}
    print(json.dumps({'config': name, 'n': n, 'env': env, 'potrf_ms': min(times[1:]), 'potrf_ms_all': times, 'potrf_tflops': n ** 3 / 3 / min(times[1:]) * 1e-9,
                      'factor_sha256_16': digest, 'lml_grad_ms': min(ev[1:]), 'lml_grad_ms_all': ev,
                      'result_sha256_16': hashlib.sha256(out.tobytes()).hexdigest()[:16], 'lml': float(out[0, 0])}), flush=True)
    del plan
    torch.cuda.empty_cache()
