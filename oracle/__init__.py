"""CPU oracle for the rom-comma dense-GP hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, in float64 numpy/scipy (and, for tiny shapes, as a literal torch-CPU
transliteration of the reference's broadcasting code), the algorithm of the reference path

    romcomma/gpf/{base,kernels,likelihoods,models}.py   (gram, LML, predict)
    romcomma/gpr/models.py:332-384,427-463               (variant/covariant dispatch, K_cho, K_inv_Y)
    romcomma/gsa/{base,calibrators,models}.py            (closed Sobol contractions, slice lists)
    romcomma/data/storage.py:162-204                     (K-fold index assignment)

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it - as the checker or the timed CPU baseline, never as the product path.
The product (``rom-comma_b200/``) never imports ``oracle`` and has no CPU fallback.

PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures for this path, and it cannot be
imported here (tensorflow / gpflow / tensorflow_probability / SALib are absent and there is no network),
so the oracle is anchored on (i) a literal transliteration of the reference code (``oracle/literal.py``),
(ii) finite differences / autograd for the gradients, (iii) Gauss-Hermite quadrature for the Sobol
integrals and (iv) the reference's own self-check identity ``MOGP.check_K_inv_Y``
(romcomma/gpr/models.py:446-463).  gpflow (pinned ``>=2.2.1,<=2.5.2`` in the reference's pyproject.toml:36) is
restated from its published algorithm: ``multivariate_normal``, ``base_conditional``, ``square_distance``,
``positive()`` = softplus (+ shift), ``Scipy`` optimizer packing order.
"""
from . import gp, sobol, folds  # noqa: F401
