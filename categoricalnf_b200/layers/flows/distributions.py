"""Prior distributions of the flows (reference layers/flows/distributions.py:16-200).

``LogisticDistribution`` is the prior of every experiment and of the categorical encodings; its
sampling (uniform -> squeezed logit, :139-152) and log-density (:154-163) are single kernels
(``cnf_logistic_sample`` / ``cnf_logistic_logprob``, csrc/elementwise.cu).  The reference draws
its uniforms on the CPU and lets callers move them (``.to(device)``); here samples are produced on
the module's CUDA device directly, so that ``.to`` is a no-op.
"""
import math
import sys

import numpy as np
import torch
import torch.nn as nn

from ... import functional as CF
from ... import ops


def _param(params, key, default=None):
    val = params.get(key, default) if params is not None else default
    return default if val is None else val


class PriorDistribution(nn.Module):

    GAUSSIAN = 0
    LOGISTIC = 1

    def __init__(self, **kwargs):
        super().__init__()
        # device anchor that follows .to()/.cuda() without entering the state dict
        self.register_buffer("_anchor", torch.empty(0), persistent=False)
        self.distribution = self._create_distribution(**kwargs)

    def _create_distribution(self, **kwargs):
        raise NotImplementedError

    def _device(self):
        if self._anchor.is_cuda:
            return self._anchor.device
        if torch.cuda.is_available():
            return torch.device("cuda", torch.cuda.current_device())
        return self._anchor.device

    def forward(self, shape=None):
        return self.sample(shape=shape)

    def sample(self, shape=None):
        return self.distribution.sample() if shape is None else self.distribution.sample(sample_shape=shape)

    def log_prob(self, x):
        return self.distribution.log_prob(x)

    def prob(self, x):
        return self.log_prob(x).exp()

    def icdf(self, x):
        assert ((x < 0) | (x > 1)).sum() == 0, \
            "[!] ERROR: Found values outside the range of 0 to 1 as input to the inverse cumulative distribution function."
        return self.distribution.icdf(x)

    def cdf(self, x):
        return self.distribution.cdf(x)

    def info(self):
        raise NotImplementedError

    @staticmethod
    def get_string_of_distributions():
        return "%i - Gaussian, %i - Logistic" % (PriorDistribution.GAUSSIAN, PriorDistribution.LOGISTIC)


class GaussianDistribution(PriorDistribution):
    """Alternative prior (:73-86); not on the hot path, stays on ``torch.distributions``."""

    def __init__(self, mu=0.0, sigma=1.0, **kwargs):
        super().__init__(mu=mu, sigma=sigma, **kwargs)
        self.mu, self.sigma = mu, sigma

    def _create_distribution(self, mu=0.0, sigma=1.0, **kwargs):
        return torch.distributions.normal.Normal(loc=mu, scale=sigma)

    def info(self):
        return "Gaussian distribution with mu=%f and sigma=%f" % (self.mu, self.sigma)


class LogisticDistribution(PriorDistribution):

    def __init__(self, mu=0.0, sigma=1.0, eps=1e-4, **kwargs):
        sigma = sigma / 1.81   # a unit logistic has a standard deviation of about 1.81 (:95)
        super().__init__(mu=mu, sigma=sigma)
        self.mu, self.sigma, self.log_sigma, self.eps = mu, sigma, float(np.log(sigma)), eps

    def _create_distribution(self, mu=0.0, sigma=1.0, **kwargs):
        return torch.distributions.uniform.Uniform(low=0.0, high=1.0)

    # -- static helpers (:117-136), element-wise on whatever device x lives --------------------------
    @staticmethod
    def shift_x(x, mu, sigma, log_sigma=None):
        """uniform (0,1) value -> logistic sample and the log-density change, logit in float64."""
        if log_sigma is None:
            log_sigma = sigma.log() if isinstance(sigma, torch.Tensor) else math.log(sigma)
        xd = x.double()
        z = (-torch.log(xd.reciprocal() - 1.0)).float() * sigma + mu
        ldj = (-torch.log(xd) - torch.log(1.0 - xd)).float() - log_sigma
        return z, ldj

    @staticmethod
    def unshift_x(x, mu, sigma, log_sigma=None):
        if log_sigma is None:
            log_sigma = sigma.log() if isinstance(sigma, torch.Tensor) else math.log(sigma)
        v = (x - mu) / sigma
        return torch.sigmoid(v), nn.functional.softplus(v) + nn.functional.softplus(-v) + log_sigma

    def _shift_x(self, x):
        return LogisticDistribution.shift_x(x, self.mu, self.sigma, self.log_sigma)

    def _unshift_x(self, x):
        return LogisticDistribution.unshift_x(x, self.mu, self.sigma, self.log_sigma)

    def sample(self, shape=None, return_ldj=False, temp=1.0):
        """Logistic(mu, sigma*temp) draws of ``shape`` on the CUDA device (one kernel).  The
        reference's ``temp != 1`` branch names an undefined class (App. B #7); the evident intent,
        a logistic with scale sigma*temp, is what is implemented."""
        shape = tuple(shape) if shape is not None else (1,)
        dev = self._device()
        sigma = self.sigma * temp
        u = torch.rand(shape, device=dev)
        z = ops.logistic_sample(shape, dev, noise=u, mu=self.mu, sigma=sigma, eps=self.eps)
        if not return_ldj:
            return z
        # -log u - log(1-u) - log sigma at u = sigmoid(v) equals -log_prob(z) - 2 log sigma
        ldj = -CF.logistic_logprob(z, mu=self.mu, sigma=sigma) - 2.0 * (self.log_sigma + math.log(temp))
        return z, ldj

    def log_prob(self, x):
        # element-wise log-density; CPU tensors are rejected by the op (no fallback)
        return CF.logistic_logprob(x, mu=self.mu, sigma=self.sigma)

    def icdf(self, x, return_ldj=False):
        assert ((x < 0) | (x > 1)).sum() == 0, \
            "[!] ERROR: Found values outside the range of 0 to 1 as input to the inverse cumulative distribution function."
        z, ldj = self._shift_x(x)
        return (z, ldj) if return_ldj else z

    def cdf(self, x, return_ldj=False):
        z, ldj = self._unshift_x(x)
        return (z, ldj) if return_ldj else z

    def info(self):
        return "Sigmoid Uniform distribution with mu=%.2f and sigma=%.2f" % (self.mu, self.sigma)


def create_prior_distribution(distribution_params):
    kind = _param(distribution_params, "distribution_type", PriorDistribution.LOGISTIC)
    given = {k: v for k, v in distribution_params.items() if v is not None}
    if kind == PriorDistribution.GAUSSIAN:
        return GaussianDistribution(**given)
    if kind == PriorDistribution.LOGISTIC:
        return LogisticDistribution(**given)
    print("[!] ERROR: Unknown distribution type %s" % str(kind))
    sys.exit(1)


def add_prior_distribution_parameters(parser, add_name=""):
    """argparse flags of the prior (:203-213)."""
    parser.add_argument("--%sprior_dist_type" % add_name, type=int, default=PriorDistribution.LOGISTIC,
                        help="Prior distribution. Options: " + PriorDistribution.get_string_of_distributions())
    parser.add_argument("--%sprior_dist_mu" % add_name, type=float, default=None, help="Center of the distribution.")
    parser.add_argument("--%sprior_dist_sigma" % add_name, type=float, default=None, help="Scale of the distribution.")
    parser.add_argument("--%sprior_dist_start_x" % add_name, type=float, default=None,
                        help="Start position of a bounded, shifted distribution.")
    parser.add_argument("--%sprior_dist_stop_x" % add_name, type=float, default=None,
                        help="End position of a bounded, shifted distribution.")
    return parser


def prior_distribution_args_to_params(args, add_name=""):
    return {
        "distribution_type": getattr(args, "%sprior_dist_type" % add_name),
        "mu": getattr(args, "%sprior_dist_mu" % add_name),
        "sigma": getattr(args, "%sprior_dist_sigma" % add_name),
        "start_x": getattr(args, "%sprior_dist_start_x" % add_name),
        "stop_x": getattr(args, "%sprior_dist_stop_x" % add_name),
    }
