"""B200-native counterparts of the reference's ``romcomma.gpf`` (extensions of gpflow for multi-output GPs)."""
from . import base, kernels, likelihoods, mean_functions, models  # noqa: F401
