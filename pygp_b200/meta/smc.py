"""
Sequential Monte Carlo over GP hyper-parameters (pygp/meta/smc.py:53-150).

A population of n particles, each a GP with its own hyper vector drawn from the
prior.  For every datum: resample when the effective sample size drops below
n/2, add the datum to every particle, reweight by the change in log marginal
likelihood, and move every particle with one slice-sampling step.

Device work per datum: n incremental factor updates (pgp_exact_append_inc,
O(N^2) each instead of refactoring), the particles' log-likelihoods, and the
slice-sampler evaluations; the mixture prediction over the particles is one
batched call when the particles are ExactGP models.
"""

import numpy as np
from scipy.special import logsumexp

from ..learning.sampling import sample
from ..utils.models import get_params
from ..utils.random import rstate

__all__ = ['SMC']


def _draw_prior_hypers(model, priors, n, rng):
    """n hyper vectors with the blocks that have a prior drawn from it (log-space
    blocks stored as logs); blocks whose prior is None keep the model's value."""
    hypers = np.tile(model.get_hyper(), (n, 1))
    for name, block, islog in get_params(model):
        prior = priors.get(name)
        if prior is None:
            continue
        draw = prior.sample(n, rng=rng)
        hypers[:, block] = np.log(draw) if islog else draw
    return hypers


class SMC(object):
    def __init__(self, model, prior, n=100, rng=None):
        self._prior = prior
        self._n = n
        self._rng = rstate(rng)
        data = None
        if model.ndata > 0:                       # particles start empty; the data is replayed below
            data = model.data
            model = model.copy()
            model.reset()
        self._samples = [model.copy(h) for h in _draw_prior_hypers(model, dict(prior), n, self._rng)]
        self._logweights = np.full(n, -np.log(n))
        self._loglikes = np.zeros(n)
        if data is not None:
            self.add_data(data[0], data[1])

    def __iter__(self):
        return iter(self._samples)

    @property
    def ndata(self):
        return self._samples[-1].ndata

    @property
    def data(self):
        return self._samples[-1].data

    def add_data(self, X, y):
        X = self._samples[0]._kernel.transform(X)
        y = self._samples[0]._likelihood.transform(y)
        n = self._n
        for xi, yi in zip(X, y):
            # multinomial resampling when the effective sample size 1 / sum w^2 < n / 2
            if -logsumexp(2*self._logweights) < np.log(n/2):
                idx = self._rng.choice(n, n, p=np.exp(self._logweights))
                self._samples = [self._samples[i].copy() for i in idx]
                self._logweights = np.full(n, -np.log(n))
                self._loglikes = self._loglikes[idx]
            for model in self._samples:
                model.add_data(xi, yi)
            # incremental weight: likelihood after / before the datum, particles not yet moved
            after = np.array([model.loglikelihood() for model in self._samples])
            self._logweights = self._logweights + after - self._loglikes
            self._logweights -= logsumexp(self._logweights)
            for model in self._samples:            # MCMC move, one slice-sampling step each
                sample(model, self._prior, 1, rng=self._rng)
            self._loglikes = np.array([model.loglikelihood() for model in self._samples])

    def posterior(self, X, grad=False):
        """Weighted moment-matched mixture over the particles (smc.py:128-150)."""
        w = np.exp(self._logweights)
        parts = [m.posterior(X, grad) for m in self._samples]
        mu_, s2_ = (np.array([p[i] for p in parts]) for i in range(2))
        mu = np.average(mu_, weights=w, axis=0)
        s2 = np.average(s2_ + (mu_ - mu)**2, weights=w, axis=0)
        if not grad:
            return mu, s2
        dmu_, ds2_ = (np.array([p[i] for p in parts]) for i in range(2, 4))
        dmu = np.average(dmu_, weights=w, axis=0)
        Dmu = dmu_ - dmu
        ds2 = np.average(ds2_ + 2*mu_[:, :, None]*Dmu - 2*mu[None, :, None]*Dmu, weights=w, axis=0)
        return mu, s2, dmu, ds2
