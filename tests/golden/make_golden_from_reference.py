"""Generates tests/golden/ref_*.npz by EXECUTING THE REFERENCE'S OWN SOURCE FILES.   Run from the repo root, in the build
container (needs /root/reference; the GPU box never runs this):

    python tests/golden/make_golden_from_reference.py

How: rom-comma is pure Python on top of TensorFlow + GPflow, neither of which is installable here.  ``tests/golden/_refshim``
provides torch-CPU float64 stand-ins for the ~60 ``tf.*`` symbols and the sliver of gpflow 2.5.2 the path uses (each restating the
library's documented semantics); with that directory first on sys.path the unmodified files
/root/reference/romcomma/{base,data,gpf,gpr,gsa}/*.py import and run.  ``romcomma/__init__.py`` itself is bypassed (it also imports
``rom`` - dead code - and ``user``, which needs SALib), by registering an empty package whose ``__path__`` points at the reference.
Nothing is copied from the reference and nothing of it is modified; two pieces of library rot are patched in the *environment*
(numpy 2 dropped ``np.NaN``; the image's Python is 3.12 where stacked classmethod/property is deprecated but functional).

What this pins: every line of rom-comma's own code on the hot path - Repository.into_K_folds and Normalization, Variance, MOStationary/RBF,
MOGaussian.add_to, MOGPR LML / predict_f, gpr.MOGP (both the covariant and the variant branch, predict, K_cho, K_inv_Y,
check_K_inv_Y, calibrate), gsa.base.Gaussian, ClosedSobol, ClosedSobolWithError, gsa.models.Sobol post-processing.
What it cannot pin: the third-party numerics themselves (TF's cholesky / triangular_solve / einsum kernels and gpflow's bijectors are
replaced by torch's float64 LAPACK-backed equivalents; tf.GradientTape by torch.autograd).
"""
from __future__ import annotations

import importlib
import sys
import tempfile
import types
import warnings
from pathlib import Path
from types import SimpleNamespace

import numpy as np

HERE = Path(__file__).resolve().parent
REFERENCE = Path('/root/reference/romcomma')


def load_reference() -> SimpleNamespace:
    warnings.filterwarnings('ignore')
    if not hasattr(np, 'NaN'):
        np.NaN = np.nan                                             # numpy 2 (reference base/definitions.py:93 uses np.NaN)
    sys.path.insert(0, str(HERE / '_refshim'))
    pkg = types.ModuleType('romcomma')
    pkg.__path__ = [str(REFERENCE)]
    sys.modules['romcomma'] = pkg
    for sub in ('gpf', 'base', 'data', 'gpr', 'gsa'):
        setattr(pkg, sub, importlib.import_module('romcomma.' + sub))
    import tensorflow as tf
    import torch
    from romcomma.data.storage import Fold, Repository
    from romcomma.gpr.models import MOGP
    from romcomma.gsa.calibrators import ClosedSobol, ClosedSobolWithError
    from romcomma.gsa.models import GSA, Sobol

    for mod in (pkg.gpf.models, pkg.gsa.calibrators, pkg.gpr.models):
        assert str(Path(mod.__file__).resolve()).startswith(str(REFERENCE)), mod.__file__

    def to_np(x):
        if hasattr(x, '__tf_tensor__'):
            x = x.__tf_tensor__()
        if isinstance(x, torch.Tensor):
            return x.detach().as_subclass(torch.Tensor).numpy().copy()
        return np.array(x, dtype=np.float64)

    def _named(model):
        from tests_golden_scenario import _walk_named
        return _walk_named(model, None)

    def variable_order(model):
        named = _named(model)
        by_id = {id(p.unconstrained_variable): key for key, p in named.items()}
        return [by_id[id(v)] for v in model.trainable_variables]

    def loss_and_grads(model, params):
        loss = model.training_loss()
        grads = torch.autograd.grad(loss, [p.unconstrained_variable for p in params])
        return float(loss.detach()), [g.detach().as_subclass(torch.Tensor).numpy() for g in grads]

    def sobol_results(gp, kind, is_error_calculated, **kwargs):
        gsa = Sobol(gp, kind, m=-1, is_error_calculated=is_error_calculated, **kwargs)
        captured = {}
        original = gsa._compose_and_save

        def capture(results):
            captured.update({key: to_np(value) for key, value in results.items()})
            return original(results)

        gsa._compose_and_save = capture                                 # instance attribute: observes, then delegates to the reference method
        gsa.calibrate()
        return captured

    return SimpleNamespace(Repository=Repository, Fold=Fold, MOGP=MOGP, ClosedSobol=ClosedSobol, ClosedSobolWithError=ClosedSobolWithError, GSA=GSA, to_np=to_np,
                           variable_order=variable_order, loss_and_grads=loss_and_grads, sobol_results=sobol_results,
                           slice_arg=lambda s: tf.constant(list(s), dtype=tf.int32))


def main():
    spec = importlib.util.spec_from_file_location('tests_golden_scenario', HERE / 'scenario.py')
    scenario = importlib.util.module_from_spec(spec)
    sys.modules['tests_golden_scenario'] = scenario
    spec.loader.exec_module(scenario)
    api = load_reference()
    only = sys.argv[1:]
    with tempfile.TemporaryDirectory() as tmp:
        for name in scenario.CASES:
            if only and name not in only:
                continue
            out = scenario.run(api, name, Path(tmp))
            np.savez_compressed(HERE / f'{name}.npz', **out)
            print(f'{name}: {len(out)} arrays, lml {out["lml"]}, fit.lml {out["fit.lml"].ravel()}')


if __name__ == '__main__':
    main()
