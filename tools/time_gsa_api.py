"""End-to-end through the public API at a BASELINE configuration: romcomma.gpr.models.MOGP (default hyper-parameters, not calibrated) ->
predict, and romcomma.gsa.models.Sobol(kind).calibrate() for the three kinds, without and with errors.  Wall clock around each call
(device work + host + csv), torch.cuda.synchronize on both sides.  Shows what the shared factorisation of a fitted GP buys: the first
calibrator pays for gram + Cholesky, the following ones (and predict) do not."""
import json, sys, tempfile, time, shutil
from pathlib import Path
import numpy as np, pandas as pd, scipy.stats, torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / 'rom-comma_b200'))
from romcomma import synthetic
from romcomma.data.storage import Repository, Fold
from romcomma.gpr.models import MOGP
from romcomma.gsa.models import GSA, Sobol

cfg = sys.argv[1] if len(sys.argv) > 1 else 'cfg3'
c = synthetic.CONFIGS[cfg]
N, M, L = c['N'], c['M'], c['L']
U = scipy.stats.qmc.LatinHypercube(d=M, scramble=False, seed=c['seed']).random(N)
Y = np.concatenate([synthetic._VECTORS[s.split('.')[0]][s.split('.')[1]](U) for s in c['outputs']], axis=1)
Y = Y + 0.04 * Y.std(axis=0, keepdims=True) * np.random.default_rng(c['seed'] + 1).standard_normal((N, L))
cols = [('X', f'X.{i:d}') for i in range(M)] + [('Y', f'Y.{i:d}') for i in range(L)]
root = Path(tempfile.mkdtemp(prefix='rc_gsa_api_'))
repo = Repository.from_df(root / cfg, pd.DataFrame(np.concatenate((U, Y), axis=1), columns=pd.MultiIndex.from_tuples(cols), dtype=float))
repo.into_K_folds(1)
fold = Fold(repo, 0)
covariant = L > 1
gp = MOGP('gp', fold, is_read=False, is_covariant=covariant, is_isotropic=False)


def wall(fn):
    torch.cuda.synchronize()
    t = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) * 1e3, out


x = fold.test_x.values[:256]
rec = {'cfg': cfg, 'N': N, 'M': M, 'L': L, 'covariant': covariant}
rec['predict_256_first_ms'], _ = wall(lambda: gp.predict(x))
rec['predict_256_again_ms'], _ = wall(lambda: gp.predict(x))
for err in (False, True):
    for kind in GSA.ALL_KINDS:
        ms, res = wall(lambda: Sobol(gp, kind, -1, err).calibrate())
        rec[f'sobol_{kind.name.lower()}{"_with_error" if err else ""}_ms'] = ms
S = pd.read_csv(fold.folder / 'gp' / 'gsa' / 'closed' / 'S.csv', index_col=[0, 1]).values
rec['closed_S_last_column'] = S[:, -1].tolist()[:4]
print(json.dumps(rec), flush=True)
shutil.rmtree(root, ignore_errors=True)
