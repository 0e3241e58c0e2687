"""gpflow.likelihoods: Likelihood/QuadratureLikelihood skeleton and Gaussian (gpflow/likelihoods/base.py, scalar_continuous.py, 2.5.2).
The 2-argument `_predict_mean_and_var(Fmu, Fvar)` signature is the <=2.5.2 one (the reason for the reference's version pin)."""
import tensorflow as tf

from .base import Module, Parameter
from .utilities import positive


class Likelihood(Module):
    def __init__(self, latent_dim=None, observation_dim=None):
        super().__init__()
        self.latent_dim, self.observation_dim = latent_dim, observation_dim

    def predict_mean_and_var(self, Fmu, Fvar):
        return self._predict_mean_and_var(Fmu, Fvar)


class QuadratureLikelihood(Likelihood):
    def __init__(self, latent_dim=None, observation_dim=None, *, quadrature=None):
        super().__init__(latent_dim=latent_dim, observation_dim=observation_dim)


class ScalarLikelihood(QuadratureLikelihood):
    def __init__(self, **kwargs):
        super().__init__(latent_dim=None, observation_dim=None, **kwargs)


class Gaussian(ScalarLikelihood):
    DEFAULT_VARIANCE_LOWER_BOUND = 1e-6

    def __init__(self, variance=1.0, variance_lower_bound=DEFAULT_VARIANCE_LOWER_BOUND, **kwargs):
        super().__init__(**kwargs)
        if variance <= variance_lower_bound:
            raise ValueError(f'The variance of the Gaussian likelihood must be strictly greater than {variance_lower_bound}')
        self.variance = Parameter(variance, transform=positive(lower=variance_lower_bound))

    def _predict_mean_and_var(self, Fmu, Fvar):
        return tf.identity(Fmu), Fvar + self.variance
