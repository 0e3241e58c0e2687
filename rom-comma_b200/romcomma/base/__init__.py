""" Type definitions and the storage-backed Model / Data / Frame base classes."""
from . import definitions, classes  # noqa: F401
