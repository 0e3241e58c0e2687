"""The workloads of BASELINE.json that SHARD (north_star: folds, per-input-subset Sobol sweep), timed at the current world size.

    python tools/scaling.py                                       # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/scaling.py

One JSON line per leg on rank 0 (stdout).  Legs:
  cfg5_all_subsets  closed Sobol V for every non-empty input subset (2^12 - 1 = 4095 masks) of cfg5 (N=8192, M=12, L=3): masks round-robin
                    over the ranks (romcomma.distributed.shard), one all_gather of the (L,L) blocks.  Strong scaling.
  cfg3_sweep        the 25-slice first-order/closed/total sweep of cfg3, sample-pair space split by row tile, one NCCL all-reduce.  Strong.
  cfg2_folds        romcomma.user.run.gpr + run.gsa over the 10 folds (+ improper fold) of the cfg2 repository (Sobol-G, N=2048, M=10, L=1):
                    fold k on rank k % world, csvs collected by rank 0.  Strong scaling, end to end through the public API (files included).
Timing: CUDA events on the launching stream for the device legs, wall clock for the API leg (it is host + device + files); barrier on both
sides, MAX over ranks.
"""
import json, os, random, shutil, sys, tempfile, time
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / 'rom-comma_b200'))
from romcomma import _capi as C, distributed, synthetic

legs = sys.argv[1:] or ['cfg5_all_subsets', 'cfg3_sweep', 'cfg2_folds']
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)
distributed.init_from_env()
rank, world = distributed.rank(), distributed.world_size()
if torch.cuda.is_available() and not distributed.is_initialized():
    torch.cuda.set_device(0)


def emit(rec):
    if rank == 0:
        os.write(_REAL_STDOUT, (json.dumps(rec) + '\n').encode())


def sync():
    distributed.barrier()
    torch.cuda.synchronize()


def device_timed(fn, reps):
    fn()
    sync()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    sync()
    return distributed.all_reduce_max(a.elapsed_time(b) / reps)


def sobol_handle(name):
    w = synthetic.config(name)
    N, M = w.X.shape
    L = w.Y.shape[1]
    dX = C.dev(w.X)
    KiY = torch.randn(L, N, dtype=torch.float64, device='cuda', generator=torch.Generator('cuda').manual_seed(7)) * 0.1
    Phi, g0, g0KY = C.sobol_prepare(dX, C.dev(w.lengthscales), C.dev(np.diag(w.F).copy()), KiY, True)
    return w, N, M, L, dX, Phi, g0KY


if 'cfg5_all_subsets' in legs:
    w, N, M, L, dX, Phi, g0KY = sobol_handle('cfg5')
    masks = list(range(1, 2 ** M))
    mine = distributed.shard(masks)
    parts = C.workspace(C.lib().rc_sobol_bufsize(N, L, len(mine)))
    result = {}

    def run():
        V = C.sobol_contract(dX, Phi, g0KY, L, True, mine, parts)                      # (len(mine), L, L) on the device
        result['V'] = distributed.all_gather_rows(V.cpu().numpy(), len(masks))         # every rank ends with all 4095 blocks
    ms = device_timed(run, reps=1)
    V = result['V']
    exps = float(sum(bin(m).count('1') > 0 for m in masks)) * (L * (L + 1) / 2) * N * N   # one exp per (pair, subset), output pairs a >= b
    emit({'leg': 'cfg5_all_subsets', 'n_gpus': world, 'scaling': 'strong', 'subsets': len(masks), 'ms': ms, 'sweeps_per_s': 1e3 / ms,
          'gexp_per_s_per_gpu': exps / (ms * 1e-3) / world * 1e-9, 'checksum': float(np.sum(V)), 'V_full': V[-1].tolist()})
    del parts
    torch.cuda.empty_cache()

if 'cfg3_sweep' in legs:
    w, N, M, L, dX, Phi, g0KY = sobol_handle('cfg3')
    slices = [(m, m + 1) for m in range(M)] + [(0, m + 1) for m in range(M)] + [(m + 1, M) for m in range(M)] + [(0, M)]
    masks = [C.slice_mask(*s) for s in slices]
    parts = C.workspace(C.lib().rc_sobol_bufsize(N, L, len(masks)))
    result = {}

    def run():
        V = C.sobol_contract(dX, Phi, g0KY, L, True, masks, parts, part=rank, nparts=world)
        result['V'] = distributed.all_reduce_sum_tensor(V)
    ms = device_timed(run, reps=20)
    emit({'leg': 'cfg3_sweep', 'n_gpus': world, 'scaling': 'strong', 'slices': len(masks), 'ms': ms, 'sweeps_per_s': 1e3 / ms,
          'checksum': float(result['V'].sum().item())})
    del parts
    torch.cuda.empty_cache()

if 'cfg2_folds' in legs:
    from romcomma.user import functions, run, sample
    root = Path(os.environ.get('RC_SCALING_ROOT', tempfile.gettempdir())) / f'rc_scaling_{os.environ.get("MASTER_PORT", "0")}'
    if rank == 0:
        shutil.rmtree(root, ignore_errors=True)
        root.mkdir(parents=True)
        np.random.seed(2)
        random.seed(2)
        fn = sample.Function(root, lambda N, M: sample.DOE.latin_hypercube(N, M, seed=2), functions.SOBOL_G.subVector('sobol_g', ['weak5_2']), N=2048, M=10,
                             noise_variance=sample.GaussianNoise.Variance(1, 0.04, False, False), overwrite_existing=True)
        fn.repo.into_K_folds(10)
        folder = fn.repo.folder
    distributed.barrier()
    from romcomma.data.storage import Repository
    folder = next(p for p in root.iterdir() if p.is_dir())
    repo = Repository(folder)
    sync()
    t0 = time.perf_counter()
    names = run.gpr('gpr', repo, is_read=None, is_covariant=False, is_isotropic=False, maxiter=50)
    sync()
    t1 = time.perf_counter()
    run.gsa('gpr', repo, is_covariant=False, is_isotropic=False)
    sync()
    t2 = time.perf_counter()
    fit_s, gsa_s = distributed.all_reduce_max(t1 - t0), distributed.all_reduce_max(t2 - t1)
    if rank == 0:
        import pandas as pd
        S = pd.read_csv(folder / 'gpr.v.a' / 'gsa' / 'closed' / 'S.csv', index_col=[0, 1])
        lm = pd.read_csv(folder / 'gpr.v.a' / 'likelihood' / 'log_marginal.csv', index_col=0)
        emit({'leg': 'cfg2_folds', 'n_gpus': world, 'scaling': 'strong', 'folds': len(repo.folds), 'fit_test_s': fit_s, 'gsa_s': gsa_s,
              'folds_per_s': len(repo.folds) / (fit_s + gsa_s), 'models': names, 'mean_log_marginal': float(np.mean(lm.values)),
              'closed_S_last_column_mean': float(S.values[:, -1].mean())})
        shutil.rmtree(root, ignore_errors=True)
    distributed.barrier()
