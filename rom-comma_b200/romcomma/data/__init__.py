""" Data storage: Repository / Fold / Normalization."""
from . import storage  # noqa: F401
