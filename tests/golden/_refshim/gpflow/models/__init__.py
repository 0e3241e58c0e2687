"""gpflow.models: GPModel, GPR (gpflow/models/model.py, gpr.py, 2.5.2)."""
from . import model, training_mixins, util  # noqa: F401
from .model import GPModel, BayesianModel  # noqa: F401
from .gpr import GPR  # noqa: F401
