""" Gaussian process regression models backed by folders of csv files."""
from . import kernels, models  # noqa: F401
