"""gpflow.models.model: BayesianModel / GPModel (predict_y = likelihood.predict_mean_and_var(predict_f))."""
from typing import Any, Tuple

import tensorflow as tf

from ..base import Module

InputData = Any
OutputData = Any
RegressionData = Tuple[Any, Any]
MeanAndVariance = Tuple[Any, Any]


class BayesianModel(Module):
    def log_prior_density(self):
        return tf.constant(0.0, tf.float64)          # the reference sets no priors

    def log_posterior_density(self, *args, **kwargs):
        return self.maximum_log_likelihood_objective(*args, **kwargs) + self.log_prior_density()

    def _training_loss(self, *args, **kwargs):
        return -(self.maximum_log_likelihood_objective(*args, **kwargs) + self.log_prior_density())


class GPModel(BayesianModel):
    def __init__(self, kernel, likelihood, mean_function=None, num_latent_gps=None):
        super().__init__()
        assert num_latent_gps is not None, 'GPModel requires specification of num_latent_gps'
        self.num_latent_gps = num_latent_gps
        if mean_function is None:
            from ..mean_functions import Zero
            mean_function = Zero()
        self.mean_function = mean_function
        self.kernel = kernel
        self.likelihood = likelihood

    def predict_y(self, Xnew, full_cov=False, full_output_cov=False):
        if full_cov or full_output_cov:
            raise NotImplementedError('The predict_y method currently supports only the argument values full_cov=False and full_output_cov=False')
        f_mean, f_var = self.predict_f(Xnew, full_cov=full_cov, full_output_cov=full_output_cov)
        return self.likelihood.predict_mean_and_var(f_mean, f_var)
