import cProfile, pstats, os, random, sys, tempfile, io
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, '/root/repo/rom-comma_b200')
os.environ['ROMCOMMA_B200_LOCKSTEP'] = '0'
from romcomma.user import functions, run, sample
with tempfile.TemporaryDirectory() as tmp:
    np.random.seed(2); random.seed(2)
    fn = sample.Function(tmp, lambda N, M: sample.DOE.latin_hypercube(N, M, seed=2), functions.SOBOL_G.subVector('sobol_g', ['weak5_2']), N=2048, M=10,
                         noise_variance=sample.GaussianNoise.Variance(1, 0.04, False, False), overwrite_existing=True)
    repo = fn.repo.into_K_folds(10)
    run.gpr('warm', repo, is_read=False, is_covariant=False, is_isotropic=False, maxiter=3)
    pr = cProfile.Profile(); pr.enable()
    run.gpr('gpr', repo, is_read=False, is_covariant=False, is_isotropic=False, maxiter=50)
    pr.disable()
    s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(45); print(s.getvalue()[:9000])
    pr = cProfile.Profile(); pr.enable()
    run.gsa('gpr', repo, is_covariant=False, is_isotropic=False)
    pr.disable()
    s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(35); print(s.getvalue()[:7000])
