""" Global sensitivity analysis: closed-form Sobol indices of a GP."""
from . import base, calibrators, models  # noqa: F401
