"""
Gaussian observation model: one hyper `log sigma`
(pygp/likelihoods/gaussian.py:23-45, likelihoods/_base.py:22-46).
The noise variance itself is applied on the device (pgp_exact_update adds
exp(2 log sigma) on the diagonal).
"""

import abc

import numpy as np

from ..utils.models import Parameterized, printable
from ..utils.random import rstate

__all__ = ['Likelihood', 'RealLikelihood', 'Gaussian']


class Likelihood(Parameterized):
    @abc.abstractmethod
    def transform(self, y):
        """Format observations as an array."""

    @abc.abstractmethod
    def sample(self, f, rng=None):
        """Draw noisy observations of latent values f."""


class RealLikelihood(Likelihood):
    def transform(self, y):
        return np.array(y, ndmin=1, dtype=float)


@printable
class Gaussian(RealLikelihood):
    def __init__(self, sigma):
        self._logsigma = np.log(float(sigma))
        self.nhyper = 1

    def _params(self):
        return [('sigma', 1, True)]

    @property
    def s2(self):
        return np.exp(self._logsigma * 2)

    def get_hyper(self):
        return np.r_[self._logsigma]

    def set_hyper(self, hyper):
        self._logsigma = float(hyper[0])

    def sample(self, f, rng=None):
        rng = rstate(rng)
        return f + rng.normal(size=len(f), scale=np.exp(self._logsigma))
