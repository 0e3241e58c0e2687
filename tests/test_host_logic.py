"""CPU: host-side logic of the drop-in API - parameter transforms, Variance chain rule, optimizer packing order, storage layout,
test functions, synthetic workloads, slice lists, and the world_size-2 sharding/gather helpers on gloo."""
import json
import os
import random
import subprocess
import sys
import textwrap
from pathlib import Path

import numpy as np
import pandas as pd
import pytest

from conftest import assert_close, ROOT
from oracle import gp as ogp, sobol as osobol


def test_transforms_match_oracle():
    from romcomma import gf_compat as gf
    u = np.linspace(-30, 30, 13)
    t = gf.positive(lower=1e-3)
    assert_close(t.forward(u), ogp.softplus(u) + 1e-3)
    v = u[u > -6]                                 # below that softplus(u) << the 1e-3 shift and the round trip is ill-conditioned
    assert_close(t.inverse(t.forward(v)), v, rtol=1e-9, atol=1e-9)
    assert_close(t.dforward(u), ogp.sigmoid(u))
    p = gf.Parameter([0.5, 2.0], transform=gf.positive())
    assert_close(p.numpy(), [0.5, 2.0])
    with pytest.raises(ValueError):
        gf.Parameter([0.0], transform=gf.positive())


def test_variance_parametrisation_and_chain_rule():
    from romcomma.gpf.base import Variance
    rng = np.random.default_rng(0)
    A = rng.normal(size=(3, 3))
    V = A @ A.T + np.eye(3)
    var = Variance(V)
    assert var.shape == (3, 3)
    assert_close(var.value.numpy(), V, rtol=1e-12)
    assert_close(var.cholesky.numpy(), np.linalg.cholesky(V), rtol=1e-12)
    assert var.value_to_broadcast.shape == (3, 1, 3, 1)
    u, low = ogp.variance_pack(V)
    assert_close(var._cholesky_diagonal.unconstrained_variable, u, rtol=1e-12)
    assert_close(var._cholesky_lower_triangle.numpy(), low, rtol=1e-12)       # row-major strict lower triangle (base.py:93 mask)
    dV = rng.normal(size=(3, 3))
    gd, gl = var._chain(dV)
    rd, rl = ogp.chain_variance(dV, u, low)
    assert_close(gd, rd)
    assert_close(gl, rl)
    with pytest.raises(ValueError):
        Variance(np.diag([1e-7, 1.0]))          # Cholesky diagonal below the 1e-3 bound (base.py:88-89)
    with pytest.raises(ValueError):
        Variance(np.ones((2, 3)))


def test_trainable_variable_order_follows_tf_module():
    """tf.Module flattens attributes in sorted-name order: kernel.lengthscales, kernel.variance.{_cholesky_diagonal,_cholesky_lower_triangle},
    likelihood.variance.{...}; the variant model gives kernel.lengthscales, kernel.variance, likelihood.variance."""
    from romcomma import gf_compat as gf
    from romcomma.gpf.kernels import RBF

    class Holder(gf.Module):
        def __init__(self):
            super().__init__()
            self.kernel = RBF(np.eye(2) * 0.5, np.ones((2, 3)))
            self.likelihood = gf.Module('lik')
            self.likelihood.variance = __import__('romcomma.gpf.base', fromlist=['Variance']).Variance(np.eye(2) * 0.1)
    h = Holder()
    assert [p.name for p in h.trainable_variables] == ['KernelVariance.cholesky_diagonal', 'KernelVariance.cholesky_lower_triangle',
                                                       'Variance.cholesky_diagonal', 'Variance.cholesky_lower_triangle']
    gf.set_trainable(h.kernel.lengthscales, True)
    assert h.trainable_variables[0].name == 'KernelLengthscales'
    k = gf.kernels.RBF(variance=2.0, lengthscales=[1.0, 2.0])
    assert [p.name for p in k.trainable_variables] == ['lengthscales', 'variance']


def test_kernel_lengthscale_broadcast_shapes():
    from romcomma.gpf.kernels import RBF
    k = RBF(np.eye(2), 0.7)
    assert (k.L, k.M) == (2, 1) and k.lengthscales.shape == (2, 1, 1) and not k.lengthscales.trainable
    k = RBF(np.eye(2), [[1.0, 2.0, 3.0], [4.0, 5.0, 6.0]])
    assert (k.L, k.M) == (2, 3)
    assert_close(k.lengthscales_neat.numpy(), [[1, 2, 3], [4, 5, 6]])


def test_test_functions_closed_forms():
    from romcomma.user import functions
    U = np.random.default_rng(0).random((11, 7))
    x = -np.pi + 2 * np.pi * U[:, :3]
    assert_close(functions.ISHIGAMI['standard'](U)[:, 0], np.sin(x[:, 0]) + 7 * np.sin(x[:, 1]) ** 2 + 0.1 * x[:, 2] ** 4 * np.sin(x[:, 0]))
    a = np.array([3, 6, 9, 18, 27.0])
    assert_close(functions.SOBOL_G['weak5_2'](U)[:, 0], np.prod((3 * np.abs(2 * U[:, :5] - 1) ** 2 + a) / (1 + a), axis=1))
    z = -1 + 2 * U
    assert_close(functions.OAKLEY2004['lin7'](U)[:, 0], z @ np.linspace(7, 3.5, 7))
    assert functions.ALL(U).shape == (11, 9)


def test_slice_lists_and_masks():
    from romcomma import _capi
    assert _capi.slice_mask(0, 3) == 0b111 and _capi.slice_mask(2, 3) == 0b100 and _capi.slice_mask(3, 3) == 0 and _capi.slice_mask(1, 4) == 0b1110
    assert osobol.m_slices(osobol.FIRST_ORDER, 3) == [(0, 1), (1, 2), (2, 3)]
    assert osobol.m_slices(osobol.CLOSED, 3) == [(0, 1), (0, 2), (0, 3)]
    assert osobol.m_slices(osobol.TOTAL, 3) == [(1, 3), (2, 3), (3, 3)]


def test_synthetic_workloads_are_deterministic_and_normalised():
    from romcomma import synthetic
    a, b = synthetic.config('cfg3', N=200), synthetic.config('cfg3', N=200)
    assert np.array_equal(a.X, b.X) and np.array_equal(a.Y, b.Y) and np.array_equal(a.lengthscales, b.lengthscales)
    assert a.X.shape == (200, 8) and a.Y.shape == (200, 4) and a.F.shape == (4, 4)
    assert_close(a.Y.mean(0), np.zeros(4), atol=1e-12)
    assert_close(a.Y.std(0, ddof=1), np.ones(4), rtol=1e-12)
    assert np.abs(a.X).max() < 7.1 and abs(a.X.mean()) < 1e-8
    assert ogp.lml_mo(a.X, a.Y, a.lengthscales, a.F, a.E) < 0


def test_repository_fold_layout_and_normalisation(tmp_path):
    from romcomma.data.storage import Repository, Fold, Normalization
    from romcomma.user import functions, sample
    np.random.seed(0)
    random.seed(3)
    fn = sample.Function(tmp_path, sample.DOE.latin_hypercube, functions.ISHIGAMI.subVector('ish', ['standard', 'sin']), N=40, M=3,
                         noise_variance=sample.GaussianNoise.Variance(2, 0.04, False, False), overwrite_existing=True)
    repo = fn.repo
    assert repo.folder.name == 'ish.M.3.u.v.4.00.N.40' and (repo.N, repo.M, repo.L) == (40, 3, 2)
    repo.into_K_folds(4)
    assert list(repo.folds) == [0, 1, 2, 3, 4] and repo.meta['has_improper_fold'] and repo.K == 4
    for k in repo.folds:
        folder = repo.fold_folder(k)
        for f in ('data.csv', 'test.csv', 'meta.json', 'normalization.csv'):
            assert (folder / f).exists(), (k, f)
    fold = Fold(repo, 1)
    assert (fold.N, fold.M, fold.L) == (30, 3, 2) and fold.meta['k'] == 1 and fold.test_x.shape == (10, 3)
    improper = Fold(repo, 4)
    assert improper.N == 40 and improper.test_data.df.shape[0] == 40
    # normalization.csv: rows mean,std,rng,min,max computed on ALL data and copied into every fold (storage.py:187,432-433)
    norm = pd.read_csv(repo.folder / 'normalization.csv', header=[0, 1], index_col=0)
    assert list(norm.index) == ['mean', 'std', 'rng', 'min', 'max']
    assert_close(norm.loc['rng'].values, 2 * np.sqrt(3) * norm.loc['std'].values, rtol=1e-12)
    assert (repo.fold_folder(0) / 'normalization.csv').read_text() == (repo.folder / 'normalization.csv').read_text()
    # the improper fold is normalised with statistics of all rows: X ~ N(0,1)-ish, Y standardised
    assert_close(improper.Y.values.mean(0), np.zeros(2), atol=1e-12)
    assert_close(improper.Y.values.std(0, ddof=1), np.ones(2), rtol=1e-10)
    raw = repo.data.df
    back = improper.normalization.undo_from(improper.data.df)
    assert_close(back.values, raw.values, rtol=1e-8, atol=1e-8)
    meta = json.loads((repo.folder / 'meta.json').read_text())
    assert meta['K'] == 4 and meta['data'] == {'X_heading': 'X', 'Y_heading': 'Y', 'N': 40, 'M': 3, 'L': 2}


def test_model_data_frames_roundtrip(tmp_path):
    from romcomma.gpr.kernels import Kernel, RBF
    data = RBF.Data(tmp_path / 'kernel', variance=np.array([[1.0, 2.0]]), lengthscales=np.array([[0.5]]))
    assert (tmp_path / 'kernel' / 'variance.csv').exists() and (tmp_path / 'kernel' / 'lengthscales.csv').exists()
    again = RBF.Data.read(tmp_path / 'kernel')
    assert_close(again.frames.variance.np, [[1.0, 2.0]])
    again.frames.lengthscales.broadcast_value((2, 3), is_diagonal=False)
    assert again.frames.lengthscales.np.shape == (2, 3)
    again.frames.variance.broadcast_value((2, 2), is_diagonal=True)
    assert_close(again.frames.variance.np, [[1.0, 0.0], [0.0, 2.0]])
    with pytest.raises(IndexError):
        again.frames.variance.broadcast_value((1, 3))
    assert RBF.TYPE_IDENTIFIER == 'kernels.RBF' and Kernel.TypeFromIdentifier('kernels.RBF') is RBF
    assert Kernel.META == {'variance': True, 'covariance': False, 'lengthscales': {'variant': True, 'covariant': False}}


def test_gsa_columns_index_and_post_processing():
    from romcomma.gsa.models import GSA, Sobol
    assert list(GSA._columns(3, 4, [0, 1, 2])) == [0, 1, 2, 3]
    assert list(GSA._index([2, 2, 4]).names) == ['l.0', 'l.1'] and len(GSA._index([2, 2, 4])) == 4
    assert [k.name for k in GSA.ALL_KINDS] == ['FIRST_ORDER', 'CLOSED', 'TOTAL'] and int(GSA.Kind.FIRST_ORDER) == 1
    assert Sobol.META == {'is_T_partial': True}

    class Cal:
        V = {0: np.full((2, 2), 4.0)}
        S = np.full((2, 2), 0.9)
    fake = Sobol.__new__(Sobol)
    fake.kind = GSA.Kind.TOTAL
    out = fake._post_calibrate(Cal, {'V': np.ones((2, 2, 3)), 'S': np.full((2, 2, 3), 0.25)})
    assert out['V'].shape == (2, 2, 4) and np.all(out['V'][..., -1] == 4.0)
    assert_close(out['S'][..., :3], np.full((2, 2, 3), 0.65))
    assert_close(out['S'][..., 3], np.full((2, 2), 0.9))


def test_shard_round_robin():
    from romcomma import distributed
    items = list(range(11))
    shards = [distributed.shard(items, r, 4) for r in range(4)]
    assert sorted(sum(shards, [])) == items and shards[1] == [1, 5, 9]
    assert distributed.shard(items) == items and distributed.world_size() == 1


WORKER = textwrap.dedent('''
    import sys, numpy as np
    sys.path.insert(0, sys.argv[1])
    from romcomma import distributed
    distributed.init_from_env('gloo')
    r, w = distributed.rank(), distributed.world_size()
    assert w == 2
    total = 7
    rows = distributed.shard(list(range(total)))
    local = np.array([[10.0 * i, i + 0.5] for i in rows])
    full = distributed.all_gather_rows(local, total)
    assert full.shape == (7, 2) and np.array_equal(full[:, 0], 10.0 * np.arange(7)), full
    assert distributed.all_reduce_max(float(r)) == 1.0 and distributed.all_reduce_sum(1.0) == 2.0
    # the Sobol sweep's partial V (slices, L, L): each rank holds the sum over its row tiles, one all-reduce adds them
    import torch
    tiles = np.arange(12.0).reshape(6, 2, 1) * np.ones((6, 2, 2))          # 6 row tiles, contribution of tile t is t everywhere
    part = torch.from_numpy(tiles[[t for t in range(6) if t % w == r]].sum(axis=0))
    distributed.all_reduce_sum_tensor(part)
    assert np.array_equal(part.numpy(), tiles.sum(axis=0)), part
    # the all-subsets sweep: blocks of consecutive masks round-robin over the ranks, one all_gather on the tensor's own device
    masks = list(range(20))
    mine = distributed.shard_blocks(masks, 4)
    assert mine == [m for m in masks if (m // 4) % w == r]
    local_t = torch.tensor([[float(m), -float(m)] for m in mine], dtype=torch.float64)
    full_t = distributed.all_gather_rows_tensor(local_t, len(masks), 4)
    assert full_t.shape == (20, 2) and np.array_equal(full_t[:, 0].numpy(), np.arange(20.0)), full_t
    ragged = distributed.shard_blocks(list(range(11)), 4)                     # a short last block
    rag_t = distributed.all_gather_rows_tensor(torch.tensor([[float(m)] for m in ragged], dtype=torch.float64), 11, 4)
    assert np.array_equal(rag_t[:, 0].numpy(), np.arange(11.0)), rag_t
    one = distributed.all_gather_rows_tensor(torch.tensor([[float(m)] for m in distributed.shard(list(range(7)))], dtype=torch.float64), 7)
    assert np.array_equal(one[:, 0].numpy(), np.arange(7.0)), one
    distributed.barrier()
    print('rank', r, 'ok')
''')


FOLD_WORKER = textwrap.dedent('''
    import sys
    from types import SimpleNamespace
    sys.path.insert(0, sys.argv[1])
    from romcomma import distributed
    distributed.init_from_env('gloo')
    from romcomma.user import run
    r = distributed.rank()
    run.Fold = lambda repo, k: k                                   # the job only needs the fold number here
    repo = SimpleNamespace(folds=range(5), L=1, N=10)
    done = run._each_owned_fold(repo, lambda k: 10 * k, side_by_side=False)
    assert done == [10 * k for k in range(5) if k % 2 == r], done
    def job(k):
        if k == 3:
            raise ValueError('fold 3 is broken')                    # fold 3 belongs to rank 1
        return k
    try:
        run._each_owned_fold(repo, job, side_by_side=False)
        print('rank', r, 'no error')
    except ValueError as e:
        print('rank', r, 'own error:', e)
    except RuntimeError as e:
        print('rank', r, 'peer error:', e)
    distributed.barrier()                                          # nobody is left behind: both ranks reach the barrier
    print('rank', r, 'ok')
''')


def test_world_size_2_fold_failure_is_agreed_on(tmp_path):
    """user.run shards the folds over the ranks; when one rank's fold raises, the other ranks learn about it (one all-reduce of an error flag)
    and raise too instead of waiting at the collecting barrier until the backend times out (round-1 advice)."""
    script = tmp_path / 'fold_worker.py'
    script.write_text(FOLD_WORKER)
    env = dict(os.environ, MASTER_ADDR='127.0.0.1', MASTER_PORT='29537', WORLD_SIZE='2', OMP_NUM_THREADS='1')
    procs = [subprocess.Popen([sys.executable, str(script), str(ROOT / 'rom-comma_b200')], env=dict(env, RANK=str(r), LOCAL_RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert 'rank 0 peer error' in outs[0] and 'rank 1 own error: fold 3 is broken' in outs[1], outs
    assert 'rank 0 ok' in outs[0] and 'rank 1 ok' in outs[1]


def test_world_size_2_gloo_gather(tmp_path):
    script = tmp_path / 'worker.py'
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR='127.0.0.1', MASTER_PORT='29531', WORLD_SIZE='2', OMP_NUM_THREADS='1')
    procs = [subprocess.Popen([sys.executable, str(script), str(ROOT / 'rom-comma_b200')], env=dict(env, RANK=str(r), LOCAL_RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert 'rank 0 ok' in outs[0] and 'rank 1 ok' in outs[1]


def test_exp_pairwise_polynomial_constants():
    """The device exp of the pairwise kernels (csrc/common.cuh: exp_pairwise) evaluates a degree-10 polynomial on |r| <= ln2/2 whose
    coefficients come from tools/exp_poly.py.  Re-derive them and check (a) the literals in the CUDA source are those numbers, (b) the
    polynomial with double-rounded coefficients is within 1e-15 of exp on the reduced range, (c) the Cody-Waite split of ln 2 is exact
    enough: hi has 21 trailing zero bits, i.e. 32 significant ones (k * hi is exact for |k| < 2^21), and hi + lo = ln 2 to 1e-18."""
    import re
    import subprocess
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    src = (root / 'rom-comma_b200' / 'csrc' / 'common.cuh').read_text()
    body = src[src.index('__device__ __forceinline__ double exp_pairwise'):]
    body = body[:body.index('inline int round_up')]
    lits = [float(v) for v in re.findall(r'fma\(p, r, ([0-9.e+-]+)\)', body)]
    lead = float(re.search(r'double p = ([0-9.e+-]+);', body).group(1))
    coef_src = [lead] + lits                                  # c10 ... c0
    assert len(coef_src) == 11
    out = subprocess.run([sys.executable, str(root / 'tools' / 'exp_poly.py'), '10'], capture_output=True, text=True, check=True).stdout
    line = [ln for ln in out.splitlines() if ln.startswith('   ')][0]
    coef_fit = [float(v) for v in line.split(',')]            # c0 ... c10
    assert coef_src[::-1] == coef_fit, 'common.cuh does not hold the coefficients tools/exp_poly.py derives'
    r = np.linspace(-0.34658, 0.34658, 20001).astype(np.longdouble)
    p = np.full_like(r, np.longdouble(coef_src[0]))
    for c in coef_src[1:]:
        p = p * r + np.longdouble(c)
    assert float(np.max(np.abs(p / np.exp(r) - 1))) < 1e-15
    hi = -float(re.search(r'fma\(k, (-[0-9.e+-]+), x\)', body).group(1))
    lo = -float(re.search(r'fma\(k, (-[0-9.e+-]+), r\)', body).group(1))
    assert (int(np.float64(hi).view(np.uint64)) & ((1 << 21) - 1)) == 0
    assert abs((np.longdouble(hi) + np.longdouble(lo)) - np.log(np.longdouble(2))) < 1e-18


def test_run_stages_follow_the_reference_hierarchy():
    """user.run flattens the (variant|covariant) x (isotropic|anisotropic) recursion of the reference (user/run.py:66-75) into an ordered list."""
    from romcomma.user.run import Stage, stages
    name = lambda plan: [s.model_name('gpr') for s in plan]
    assert name(stages(False, None, None)) == ['gpr.v.i', 'gpr.v.a', 'gpr.c.a']
    assert [s.is_read for s in stages(False, None, None)] == [False, None, None]
    assert name(stages(True, None, False)) == ['gpr.v.a', 'gpr.c.a'] and [s.is_read for s in stages(True, None, False)] == [True, None]
    assert name(stages(None, None, True)) == ['gpr.v.i', 'gpr.c.i']
    assert name(stages(False, True, None)) == ['gpr.c.i', 'gpr.c.a'] and name(stages(False, False, False)) == ['gpr.v.a']
    assert stages(None, True, False) == [Stage(True, False, None)]


def test_run_seeds_a_model_from_its_nearest_ancestor(tmp_path):
    """is_read=None (user/run.py:76-86): an existing folder is kept; a covariant model copies the variant one of the same isotropy, else the
    isotropic one of the same covariance; without an ancestor the model starts from the defaults."""
    from types import SimpleNamespace
    from romcomma.user.run import Stage, _seed_from_ancestor
    fold = SimpleNamespace(folder=tmp_path)
    assert _seed_from_ancestor(fold, 'gpr', Stage(False, True, None)) is False            # nothing there: defaults
    (tmp_path / 'gpr.v.i').mkdir()
    (tmp_path / 'gpr.v.i' / 'kernel').mkdir()
    (tmp_path / 'gpr.v.i' / 'kernel' / 'variance.csv').write_text('v.i')
    assert _seed_from_ancestor(fold, 'gpr', Stage(False, False, None)) is True            # v.a <- v.i
    assert (tmp_path / 'gpr.v.a' / 'kernel' / 'variance.csv').read_text() == 'v.i'
    (tmp_path / 'gpr.v.a' / 'kernel' / 'variance.csv').write_text('v.a')
    assert _seed_from_ancestor(fold, 'gpr', Stage(True, False, None)) is True             # c.a <- v.a (same isotropy) before c.i
    assert (tmp_path / 'gpr.c.a' / 'kernel' / 'variance.csv').read_text() == 'v.a'
    (tmp_path / 'gpr.c.a' / 'kernel' / 'variance.csv').write_text('kept')
    assert _seed_from_ancestor(fold, 'gpr', Stage(True, False, None)) is True and (tmp_path / 'gpr.c.a' / 'kernel' / 'variance.csv').read_text() == 'kept'


def test_lockstep_broker_batches_concurrent_optimisers(monkeypatch):
    """romcomma.lockstep: optimiser threads that run side by side get their evaluations in ONE batched call per lock step, each follows the
    trajectory it has alone, a thread that finishes early does not stall the others, and an error reaches the thread it belongs to."""
    import scipy.optimize
    import torch
    from romcomma import _capi, lockstep
    calls = []

    class FakePlan:                                   # stands in for rc_lml_grad_multi: "lml" = -|ls - Y[0]|^2, "dls" its gradient
        def __init__(self, Xs, Ys, L, flags):
            self.targets = [x[0].numpy().copy() for x in Xs]
            self.info = torch.zeros(len(Xs), dtype=torch.int32)

        def __call__(self, ls, F, E):
            calls.append(ls.shape[0])
            self.ls = ls.numpy().copy()
            self.info.zero_()
            for z in range(len(self.targets)):
                if self.ls[z, 0] > 1e6:
                    self.info[z] = 3
            self.out = torch.zeros(len(self.targets), 1)
            return self.out

        def unpack(self, out):
            return [{'lml': -float(np.sum((x - t) ** 2)), 'dls': -2.0 * (x - t)[None]} for x, t in zip(self.ls, self.targets)]

    monkeypatch.setattr(_capi, 'LmlGradMultiPlan', FakePlan)

    def fit(target, x0, poison=False):
        X, Y = torch.as_tensor(target)[None].repeat(10, 1), torch.zeros(10, 1)
        model = object()

        def fun(x):
            broker = lockstep.current()
            if broker is None:                        # no session (a single job): the model's own, unbatched evaluation
                calls.append(1)
                return float(np.sum((x - target) ** 2)), 2.0 * (x - target)
            res = broker.evaluate(model, X, Y, 1, 0, (x * (1e9 if poison else 1.0))[None], np.eye(1), np.eye(1))
            return -res['lml'], -res['dls'].reshape(-1)
        with lockstep.participating():
            return scipy.optimize.minimize(fun, x0, jac=True, method='L-BFGS-B', options={'maxiter': 50})

    rng = np.random.default_rng(0)
    targets = [rng.normal(size=3) * (k + 1) for k in range(5)]
    starts = [rng.normal(size=3) * 10 ** k for k in range(5)]              # different scales: different iteration counts
    alone = [lockstep.run_together([lambda t=t, s=s: fit(t, s)])[0] for t, s in zip(targets, starts)]       # one job each: no session, no threads
    assert all(c == 1 for c in calls)
    n_alone = len(calls)
    together = lockstep.run_together([(lambda t=t, s=s: fit(t, s)) for t, s in zip(targets, starts)])
    for a, b, t in zip(alone, together, targets):
        assert np.array_equal(a.x, b.x) and a.nfev == b.nfev and np.allclose(b.x, t, atol=1e-6)
    batched = calls[n_alone:]
    assert max(batched) >= 2 and len(batched) < n_alone, 'optimisers that run side by side must share launches (how many at a time depends on thread timing)'
    with pytest.raises(_capi.RomcommaB200Error):                         # a failed Cholesky in one problem raises in that thread, the others finish
        lockstep.run_together([lambda: fit(targets[0], starts[0]), lambda: fit(targets[1], starts[1], poison=True)])


def test_exp_tab_constants():
    """The table form of the device exp (csrc/common.cuh: exp_tab, the exp of the Sobol sweep / lattice / error, gram and gradient kernels):
    (a) the polynomial literals are the coefficients tools/exp_poly.py derives for g(r) = (e^r - 1 - r)/r^2 on |r| <= ln2/64 and
    1 + r(1 + r g(r)) is within 3e-16 of exp there; (b) the 32 table entries are 2^(j/32) correctly rounded; (c) the scale 32/ln2 and the
    one-step reduction constant ln2/32 are the doubles nearest to those numbers; (d) a numpy emulation of the whole function (FMAs in long
    double, which over-states nothing that matters here) stays within 6e-16 of exp on [-40, 2] and within 6e-16 + 3.4e-17 |x| beyond."""
    import re
    import subprocess
    import sys
    from decimal import Decimal, getcontext
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    src = (root / 'rom-comma_b200' / 'csrc' / 'common.cuh').read_text()
    body = src[src.index('__device__ __forceinline__ double exp_tab('):]
    body = body[:body.index('// Programmatic dependent launch')]
    lead = float(re.search(r'double s = ([0-9.e+-]+);', body).group(1))
    lits = [float(v) for v in re.findall(r's = fma\(s, r, ([0-9.e+-]+)\);', body)]
    assert lits[-1] == 1.0 and len(lits) == 4
    coef_src = ([lead] + lits[:-1])[::-1]                      # c0 .. c3 of g
    out = subprocess.run([sys.executable, str(root / 'tools' / 'exp_poly.py'), 'table'], capture_output=True, text=True, check=True).stdout
    nums = [float(v) for v in re.findall(r'-?\d\.\d+e[+-]\d+', out)]
    assert coef_src == nums[:4], 'common.cuh does not hold the coefficients tools/exp_poly.py derives for the table form'
    assert nums[4] < 3e-16
    getcontext().prec = 60
    tab_src = [float(v) for v in re.findall(r'(\d\.\d{17}e[+-]\d\d),', src[src.index('rc_exp2_table[32] = {'):src.index('};', src.index('rc_exp2_table[32] = {'))])]
    assert len(tab_src) == 32
    assert tab_src == [float(Decimal(2) ** (Decimal(j) / Decimal(32))) for j in range(32)]
    ln2 = Decimal(2).ln()
    scale = float(re.search(r'fma\(x, ([0-9.e+-]+), magic\)', body).group(1))
    step = -float(re.search(r'fma\(k, (-[0-9.e+-]+), x\)', body).group(1))
    assert scale == float(Decimal(32) / ln2) and step == float(ln2 / Decimal(32))
    ld = np.longdouble
    x = np.concatenate([np.random.default_rng(3).uniform(-40.0, 2.0, 200000), np.random.default_rng(4).uniform(-700.0, 700.0, 50000)])
    k = np.rint(x * scale)
    r = (x.astype(ld) - k.astype(ld) * ld(step)).astype(np.float64)           # one FMA: exact product, one rounding
    ki = k.astype(np.int64)
    T = np.array(tab_src)[ki & 31]
    s = np.full_like(r, lead)
    for c in lits:
        s = (s.astype(ld) * r + ld(c)).astype(np.float64)
    p = (T.astype(ld) + (T * r).astype(ld) * s).astype(np.float64)
    y = np.ldexp(p, ki >> 5)
    rel = np.abs(y.astype(ld) / np.exp(x.astype(ld)) - 1).astype(np.float64)
    assert rel[:200000].max() < 6e-16 + 3.4e-17 * 40
    assert np.all(rel <= 6e-16 + 3.4e-17 * np.abs(x))
