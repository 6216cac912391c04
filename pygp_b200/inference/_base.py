"""
GP inference interface: the reference's contract (pygp/inference/_base.py:27-242)
with the data and the sufficient statistics resident on the device.

Hyper vector layout `[likelihood | kernel | mean]` (_base.py:91-105);
`set_hyper` re-factorises when data is present (_base.py:107-108); `add_data`
tries the incremental update first and falls back to a full one on
NotImplementedError (_base.py:133-141).
"""

import abc

import numpy as np

from ..utils.models import Parameterized

__all__ = ['GP']


class GP(Parameterized):
    def __init__(self, likelihood, kernel, mean):
        self._likelihood = likelihood
        self._kernel = kernel
        self._mean = float(mean)
        self._X = None
        self._y = None
        self.nhyper = self._likelihood.nhyper + self._kernel.nhyper + 1

    def reset(self):
        """Remove all data from the model."""
        self._X = None
        self._y = None

    def __repr__(self):
        def indent(pre, text):
            return pre + ('\n' + ' ' * len(pre)).join(text.splitlines())
        return indent(type(self).__name__ + '(', ',\n'.join([
            indent('likelihood=', repr(self._likelihood)),
            indent('kernel=', repr(self._kernel)),
            indent('mean=', str(self._mean))]) + ')')

    def _params(self):
        out = [('like.%s' % p[0],) + tuple(p[1:]) for p in self._likelihood._params()]
        out += [('kern.%s' % p[0],) + tuple(p[1:]) for p in self._kernel._params()]
        out += [('mean', 1, False)]
        return out

    @classmethod
    @abc.abstractmethod
    def from_gp(cls, gp):
        """New model of this class with the likelihood, kernel, mean and data of `gp`."""

    def get_hyper(self):
        return np.r_[self._likelihood.get_hyper(), self._kernel.get_hyper(), self._mean]

    def set_hyper(self, hyper):
        hyper = np.asarray(hyper, dtype=float)
        a, b = self._likelihood.nhyper, self._kernel.nhyper
        self._likelihood.set_hyper(hyper[:a])
        self._kernel.set_hyper(hyper[a:a + b])
        self._mean = float(hyper[-1])
        if self.ndata > 0:
            self._update()

    @property
    def ndata(self):
        return 0 if self._X is None else self._X.shape[0]

    @property
    def data(self):
        return (self._X, self._y)

    def add_data(self, X, y):
        X = self._kernel.transform(X)
        y = self._likelihood.transform(y)
        if self._X is None:
            self._X, self._y = X.copy(), y.copy()
            self._update()
            return
        try:
            self._updateinc(X, y)
            self._X, self._y = np.r_[self._X, X], np.r_[self._y, y]
        except NotImplementedError:
            self._X, self._y = np.r_[self._X, X], np.r_[self._y, y]
            self._update()

    def posterior(self, X, grad=False):
        """Marginal posterior mean and variance at the rows of X."""
        return self._marg_posterior(self._kernel.transform(X), grad)

    def sample(self, X, m=None, latent=True, rng=None):
        """Sample values from the posterior at the points X (_base.py:143-177):
        an n-vector, or an (m, n) array of m samples; `latent=False` adds
        observation noise.  The normal variates come from the host rng in the
        reference's order; the Cholesky of the joint covariance and the
        transform run on the device."""
        from .. import _lib
        from ..utils.random import rstate
        X = self._kernel.transform(X)
        flatten = (m is None)
        m = 1 if flatten else m
        n = len(X)
        rng = rstate(rng)
        mu, Sigma = self._full_posterior(X)
        Z = _lib.as_f64(rng.normal(size=(m, n)), 2)
        f = np.empty((m, n))
        ctx = _lib.context()
        mu, Sigma = _lib.as_f64(mu), _lib.as_f64(Sigma, 2)
        _lib.check(ctx, _lib.lib().pgp_mvn_transform(ctx.handle, _lib.ptr(mu), _lib.ptr(Sigma), n, 1e-10,
                                                     _lib.ptr(Z), m, _lib.ptr(f)))
        if not latent:
            f = self._likelihood.sample(f.ravel(), rng).reshape(m, n)
        return f.ravel() if flatten else f

    def sample_fourier(self, N, rng=None):
        raise NotImplementedError('GP.sample_fourier is outside the B200 hot path')

    @abc.abstractmethod
    def _update(self):
        """Recompute the sufficient statistics from all the data."""

    def _updateinc(self, X, y):
        raise NotImplementedError

    def _full_posterior(self, X):
        raise NotImplementedError('_full_posterior is not provided for this model (next: N3)')

    @abc.abstractmethod
    def _marg_posterior(self, X, grad=False):
        """Marginal posterior at X."""

    @abc.abstractmethod
    def loglikelihood(self, grad=False):
        """Log marginal likelihood, optionally with its hyper-gradient."""
