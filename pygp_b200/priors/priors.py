"""
Priors over hyper-parameters with the reference's interface
(`sample(size, rng)`, `logprior(theta)`, `ndim`): pygp/priors/priors.py:20-150.
Scalar host work; used by learning.sample / meta.MCMC.
"""

import numpy as np
import scipy.stats as ss

from ..utils.random import rstate

__all__ = ['Uniform', 'Gaussian', 'Gamma', 'LogNormal', 'Horseshoe']


def _vec(x):
    return np.array(x, dtype=float, ndmin=1)


class Uniform(object):
    def __init__(self, a, b):
        self._a, self._b = _vec(a), _vec(b)
        self.ndim = len(self._a)
        if len(self._a) != len(self._b):
            raise ValueError("bound sizes don't match")
        if np.any(self._b < self._a):
            raise ValueError('malformed upper/lower bounds')

    def sample(self, size=1, rng=None):
        rng = rstate(rng)
        return self._a + (self._b - self._a) * rng.rand(size, self.ndim)

    def logprior(self, theta):
        theta = _vec(theta)
        inside = all(a <= t <= b for a, b, t in zip(self._a, self._b, theta))
        return 0.0 if inside else -np.inf


class Gaussian(object):
    def __init__(self, mu, var):
        self._mu, self._s2 = _vec(mu), np.array(var, dtype=float, ndmin=1)
        self.ndim = len(self._mu)
        if self._s2.ndim == 1:
            self._std = np.sqrt(self._s2)
        elif self._s2.ndim == 2:
            self._std = np.linalg.cholesky(self._s2)
        else:
            raise ValueError('Argument `var` can be at most a rank 2 array.')

    def sample(self, size=1, rng=None):
        rng = rstate(rng)
        z = rng.randn(size, self.ndim)
        return self._mu + (self._std * z if self._std.ndim == 1 else np.dot(z, self._std))

    def logprior(self, theta):
        theta = _vec(theta)
        if self._s2.ndim == 1:
            return -0.5 * (np.sum(np.log(self._s2)) + np.sum((theta - self._mu)**2 / self._s2)
                           + self.ndim * np.log(2*np.pi))
        return ss.multivariate_normal.logpdf(theta, mean=self._mu, cov=self._s2)


class Gamma(object):
    def __init__(self, k, scale, min=0.):
        self._k, self._scale, self._min = _vec(k), _vec(scale), min
        self.ndim = len(self._k)

    def sample(self, size=1, rng=None):
        rng = rstate(rng)
        return np.vstack([self._min + rng.gamma(k, s, size=size)
                          for k, s in zip(self._k, self._scale)]).T

    def logprior(self, theta):
        theta = _vec(theta)
        if np.any(theta <= self._min):
            return -np.inf
        # argument order as the reference (priors.py:101-104)
        return ss.gamma.logpdf(self._k, theta, scale=self._scale, loc=self._min).sum()


class LogNormal(object):
    def __init__(self, mu=0., sigma=1., min=0.):
        self._mu, self._sigma, self._min = _vec(mu), _vec(sigma), min
        self.ndim = len(self._mu)

    def sample(self, size=1, rng=None):
        rng = rstate(rng)
        return np.vstack([self._min + rng.lognormal(m, s, size=size)
                          for m, s in zip(self._mu, self._sigma)]).T

    def logprior(self, theta):
        theta = _vec(theta)
        if np.any(theta <= self._min):
            return -np.inf
        return ss.lognorm.logpdf(theta, self._sigma, scale=np.exp(self._mu), loc=self._min).sum()


class Horseshoe(object):
    def __init__(self, scale=1., min=0.):
        self._scale, self._min = _vec(scale), min
        self.ndim = len(self._scale)

    def sample(self, size=1, rng=None):
        raise NotImplementedError

    def logprior(self, theta):
        theta = _vec(theta)
        if np.any(theta <= self._min):
            return -np.inf
        return np.log(np.log(1 + (self._scale / (theta - self._min))**2)).sum()
