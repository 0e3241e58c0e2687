"""A torch-CPU stand-in for the sliver of gpflow (pinned >=2.2.1,<=2.5.2 by the reference, pyproject.toml:36) that rom-comma's hot
path imports.  TEST INFRASTRUCTURE ONLY - see ../tensorflow/__init__.py.  Each piece restates gpflow 2.5.2's published algorithm
(file named in its docstring); nothing here is on a product import path.
"""
from . import config, utilities, base, logdensities, conditionals, mean_functions, kernels, likelihoods, models, optimizers  # noqa: F401
from .base import Module, Parameter, set_trainable  # noqa: F401

default_float = config.default_float
default_int = config.default_int
