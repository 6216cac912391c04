"""
Hyper-parameter marginalisation by MCMC (pygp/meta/mcmc.py:22-93).

The reference keeps n deep copies of the GP (one factorisation each) and loops
over them in `posterior`.  Here the n sampled hyper vectors are kept as a
matrix and the mixture prediction is ONE batched device call
(pgp_batched_predict: n Gram builds, n Cholesky factorisations, n solves in the
same launches).  Model objects are materialised only when iterated.
"""

import ctypes as C

import numpy as np

from .. import _lib
from ..inference.exact import ExactGP
from ..learning.sampling import sample
from ..utils.random import rstate

__all__ = ['MCMC']


class MCMC(object):
    # Opt-in multi-GPU mixture: set `sharded = True` (or a process group)
    # in a one-process-per-GPU job where EVERY rank holds the same chain (same data, same
    # rng seed) and calls `posterior` in lockstep; the n sampled models are then split
    # across the ranks.  Never inferred from the world size: a call on a subset of ranks,
    # or chains that differ per rank (rng=None), would hang or mix different posteriors.
    sharded = False

    def __init__(self, model, prior, n=100, burn=100, rng=None):
        self._model = model.copy()
        self._prior = prior
        self._hypers = np.empty((0, self._model.nhyper))
        self._models = None
        self._n = n
        self._burn = burn
        self._rng = rstate(rng)
        if self._model.ndata > 0:
            if self._burn > 0:
                sample(self._model, self._prior, self._burn, rng=self._rng)
            self._resample()

    def _resample(self):
        self._hypers = sample(self._model, self._prior, self._n, raw=True, rng=self._rng)
        self._models = None

    @property
    def _samples(self):
        """List of model copies, one per sampled hyper vector (built lazily)."""
        if self._models is None:
            self._models = [self._model.copy(h) for h in self._hypers]
        return self._models

    def __iter__(self):
        return iter(self._samples)

    @property
    def ndata(self):
        return self._model.ndata

    @property
    def data(self):
        return self._model.data

    def add_data(self, X, y):
        nprev = self._model.ndata
        self._model.add_data(X, y)
        if self._model.ndata > 2*nprev and self._burn > 0:
            sample(self._model, self._prior, self._burn, rng=self._rng)
        self._resample()

    def _component_posteriors(self, X):
        model = self._model
        if isinstance(model, ExactGP) and model.ndata > 0 and len(self._hypers) > 0:
            X = _lib.as_f64(model._kernel.transform(X), 2)
            Xd, yd = _lib.as_f64(model._X, 2), _lib.as_f64(model._y, 1)
            H = _lib.as_f64(self._hypers, 2)
            B, ms = len(H), len(X)
            mu, s2 = np.empty((B, ms)), np.empty((B, ms))
            info = np.zeros(B, dtype=np.int32)
            ctx = _lib.context()
            _lib.check(ctx, _lib.lib().pgp_batched_predict(
                ctx.handle, model._kernel._spec(), _lib.ptr(Xd), _lib.ptr(yd), len(Xd),
                _lib.ptr(H), B, _lib.ptr(X), ms, _lib.ptr(mu), _lib.ptr(s2),
                info.ctypes.data_as(C.POINTER(C.c_int32))))
            if np.any(info):
                raise np.linalg.LinAlgError('sampled hyper-parameters give a non positive definite kernel matrix')
            return mu, s2
        parts = [m.posterior(X) for m in self._samples]
        return np.array([p[0] for p in parts]), np.array([p[1] for p in parts])

    def posterior(self, X, grad=False):
        """Moment-matched mixture over the sampled models (mcmc.py:75-93)."""
        if grad:
            parts = [m.posterior(X, True) for m in self._samples]
            mu_, s2_, dmu_, ds2_ = (np.array([p[i] for p in parts]) for i in range(4))
            mu = np.mean(mu_, axis=0)
            s2 = np.mean(s2_ + (mu_ - mu)**2, axis=0)
            dmu = np.mean(dmu_, axis=0)
            Dmu = dmu_ - dmu
            ds2 = np.mean(ds2_ + 2*mu_[:, :, None]*Dmu - 2*mu[None, :, None]*Dmu, axis=0)
            return mu, s2, dmu, ds2
        from .. import sharding
        model = self._model
        group = None if self.sharded is True else self.sharded
        if (self.sharded and sharding.world(group)[1] > 1 and isinstance(model, ExactGP) and model.ndata > 0
                and len(self._hypers) > 0):
            # one process per GPU, identical chains (same rng) on every rank: the
            # n sampled models are split across the ranks
            return sharding.sharded_mixture_posterior(model, self._hypers, model._kernel.transform(X), group)
        mu_, s2_ = self._component_posteriors(X)
        mu = np.mean(mu_, axis=0)
        s2 = np.mean(s2_ + (mu_ - mu)**2, axis=0)
        return mu, s2
