""" User-level entry points: run.gpr / run.gsa, sampling of test functions, result collection, contexts."""
from . import contexts, functions, sample, results, run  # noqa: F401
