"""gpflow.models.GPR (gpflow/models/gpr.py GPR_deprecated, 2.5.2)."""
import tensorflow as tf

from .. import likelihoods
from ..conditionals import base_conditional
from ..logdensities import multivariate_normal
from .model import GPModel
from .training_mixins import InternalDataTrainingLossMixin
from .util import data_input_to_tensor


def add_noise_cov(K, likelihood_variance):
    k_diag = tf.linalg.diag_part(K)
    s_diag = tf.fill(k_diag.shape, likelihood_variance.__tf_tensor__() if hasattr(likelihood_variance, '__tf_tensor__') else likelihood_variance)
    return tf.linalg.set_diag(K, k_diag + s_diag)


class GPR(GPModel, InternalDataTrainingLossMixin):
    def __init__(self, data, kernel, mean_function=None, noise_variance=1.0):
        likelihood = likelihoods.Gaussian(noise_variance)
        _, Y_data = data
        super().__init__(kernel, likelihood, mean_function, num_latent_gps=Y_data.shape[-1])
        self.data = data_input_to_tensor(data)

    def maximum_log_likelihood_objective(self):
        return self.log_marginal_likelihood()

    def log_marginal_likelihood(self):
        X, Y = self.data
        K = self.kernel(X)
        ks = add_noise_cov(K, self.likelihood.variance)
        L = tf.linalg.cholesky(ks)
        m = self.mean_function(X)
        log_prob = multivariate_normal(Y, m, L)
        return tf.reduce_sum(log_prob)

    def predict_f(self, Xnew, full_cov=False, full_output_cov=False):
        X, Y = self.data
        Xnew = tf.convert_to_tensor(Xnew, tf.float64)
        err = Y - self.mean_function(X)
        kmm = self.kernel(X)
        knn = self.kernel(Xnew, full_cov=full_cov)
        kmn = self.kernel(X, Xnew)
        kmm_plus_s = add_noise_cov(kmm, self.likelihood.variance)
        f_mean_zero, f_var = base_conditional(kmn, kmm_plus_s, knn, err, full_cov=full_cov, white=False)
        f_mean = f_mean_zero + self.mean_function(Xnew)
        return f_mean, f_var
