import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / 'rom-comma_b200'):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))

GOLDEN = ROOT / 'tests' / 'golden'

# Parity tolerance stated by BASELINE.json:north_star for all floating-point results (LML, gradients, predictions, Sobol V/S).
RTOL, ATOL = 1e-8, 1e-10


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def assert_close(actual, desired, rtol=RTOL, atol=ATOL, what=''):
    actual, desired = np.asarray(actual, dtype=float), np.asarray(desired, dtype=float)
    assert actual.shape == desired.shape, f'{what}: shape {actual.shape} != {desired.shape}'
    err = np.abs(actual - desired)
    bound = atol + rtol * np.abs(desired)
    worst = np.max(err / bound) if err.size else 0.0
    assert worst <= 1.0, f'{what}: max |err|/(atol+rtol|ref|) = {worst:.3e} (atol={atol}, rtol={rtol}); max abs err {err.max():.3e}'


@pytest.fixture(scope='session')
def golden():
    return {p.stem: dict(np.load(p)) for p in sorted(GOLDEN.glob('*.npz')) if not p.stem.startswith('ref_')}


def random_problem(N, M, L, seed=0, full_F=False, full_E=True):
    rng = np.random.default_rng(seed)
    X = rng.normal(size=(N, M))
    W = rng.normal(size=(M, L))
    Y = np.sin(X @ W) + 0.1 * rng.normal(size=(N, L))
    Y = (Y - Y.mean(0)) / Y.std(0)
    ls = rng.uniform(0.5, 3.0, (L, M))
    F = np.diag(rng.uniform(0.5, 2.0, L))
    if full_F:
        A = rng.normal(size=(L, L))
        F = A @ A.T / L + np.eye(L)
    E = 0.01 * np.eye(L)
    if full_E:
        B = rng.normal(size=(L, L))
        E = 0.01 * (np.eye(L) + B @ B.T / L)
    return X, Y, ls, F, E
