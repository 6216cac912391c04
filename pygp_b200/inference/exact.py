"""
Exact GP regression on the device (pygp/inference/exact.py:20-143).

`_update`, `loglikelihood` and `_marg_posterior(grad=False)` are single C-ABI
calls (pgp_exact_update / _loglike / _predict); X, y, the factor L = R^T and
a = R^-T (y - mean) stay in HBM between calls, so an optimiser step moves only
the hyper vector in and (lZ, dlZ) out.
"""

import ctypes as C

import numpy as np

from .. import _lib
from ..likelihoods import Gaussian
from ._base import GP

__all__ = ['ExactGP']


class _DeviceModel(object):
    """Owner of a `pgp_model*`; deep-copy clones the device buffers
    (Parameterized.copy semantics, utils/models.py:47-55)."""

    def __init__(self, ctx, handle):
        self.ctx, self.handle = ctx, handle

    @classmethod
    def create(cls, kernel, X, y):
        ctx = _lib.context()
        X, y = _lib.as_f64(X, 2), _lib.as_f64(y, 1)
        h = C.c_void_p()
        _lib.check(ctx, _lib.lib().pgp_exact_create(ctx.handle, kernel._spec(), _lib.ptr(X),
                                                    _lib.ptr(y), len(X), C.byref(h)))
        return cls(ctx, h)

    def __deepcopy__(self, memo):
        h = C.c_void_p()
        _lib.check(self.ctx, _lib.lib().pgp_model_clone(self.handle, C.byref(h)))
        return _DeviceModel(self.ctx, h)

    def __del__(self):
        try:
            if self.handle:
                _lib.lib().pgp_model_destroy(self.handle)
                self.handle = None
        except Exception:       # interpreter shutdown
            pass


class ExactGP(GP):
    # Opt-in multi-GPU evaluation: set to e.g. {'nb': 512, 'min_n': 32768} in a
    # one-process-per-GPU job whose ranks all hold the same model and call
    # set_hyper / loglikelihood(True) in lockstep (replicated optimiser); `_update` then runs the
    # 1-D block-column distributed Cholesky and `loglikelihood(True)` the block-column gradient
    # (pygp_b200/csrc/dist.cu) for ndata >= min_n.  'force': also with a single rank (tests).
    distributed = None

    def __init__(self, likelihood, kernel, mean):
        if not isinstance(likelihood, Gaussian):
            raise ValueError('exact inference requires a Gaussian likelihood')
        super(ExactGP, self).__init__(likelihood, kernel, mean)
        self._dev = None
        self._ndev = 0          # rows already resident on the device

    @classmethod
    def from_gp(cls, gp):
        newgp = cls(gp._likelihood.copy(), gp._kernel.copy(), gp._mean)
        if gp.ndata > 0:
            X, y = gp.data
            newgp.add_data(X, y)
        return newgp

    def reset(self):
        self._dev = None
        self._ndev = 0
        super(ExactGP, self).reset()

    # -- device state ----------------------------------------------------------
    def _update(self):
        L = _lib.lib()
        n = self.ndata
        if self._dev is None:
            self._dev = _DeviceModel.create(self._kernel, self._X, self._y)
        elif self._ndev < n:
            Xn, yn = _lib.as_f64(self._X[self._ndev:], 2), _lib.as_f64(self._y[self._ndev:], 1)
            _lib.check(self._dev.ctx, L.pgp_exact_append(self._dev.handle, _lib.ptr(Xn), _lib.ptr(yn), len(Xn)))
        self._ndev = n
        cfg = self._dist_cfg()
        if cfg:
            from .. import distchol
            distchol.distributed_update(self, nb=cfg.get('nb', 512), group=cfg.get('group'))
            return
        hyp = _lib.as_f64(self.get_hyper())
        _lib.check(self._dev.ctx, L.pgp_exact_update(self._dev.handle, _lib.ptr(hyp)))

    def _updateinc(self, X, y):
        """Grow the factor by the new rows (exact.py:57-62): O(n^2 m) on the device."""
        if self._dev is None or self._ndev != self.ndata:
            raise NotImplementedError
        cfg = self.distributed
        if cfg and self.ndata + len(X) >= cfg.get('min_n', 32768):
            raise NotImplementedError          # the distributed path refactors
        Xn, yn = _lib.as_f64(X, 2), _lib.as_f64(y, 1)
        try:
            _lib.check(self._dev.ctx, _lib.lib().pgp_exact_append_inc(self._dev.handle, _lib.ptr(Xn), _lib.ptr(yn),
                                                                      len(Xn)))
        except Exception:
            # the device copy may already hold the new rows while the host does not:
            # drop it, the next _update uploads the host's data again
            self._dev, self._ndev = None, 0
            raise
        self._ndev += len(Xn)

    def _ensure_dev(self):
        """Rebuild the device state from the host data if it was dropped (a failed
        incremental update): the reference keeps working from its old factor, here the
        model is refactored on next use."""
        if self._dev is None and self.ndata > 0:
            self._update()

    def _factor(self):
        self._ensure_dev()
        n = self.ndata
        R, a = np.empty((n, n)), np.empty(n)
        _lib.check(self._dev.ctx, _lib.lib().pgp_exact_get_factor(self._dev.handle, _lib.ptr(R), _lib.ptr(a)))
        return R, a

    @property
    def _R(self):
        """Upper Cholesky factor as the reference stores it (exact.py:54)."""
        return None if self._dev is None else self._factor()[0]

    @property
    def _a(self):
        return None if self._dev is None else self._factor()[1]

    # -- GP interface ------------------------------------------------------------
    def _dist_cfg(self):
        """The opt-in distributed configuration if it applies to this model right now."""
        cfg = self.distributed
        if cfg and self.ndata >= cfg.get('min_n', 32768):
            from .. import sharding
            if sharding.world(cfg.get('group'))[1] > 1 or cfg.get('force'):
                return cfg
        return None

    def loglikelihood(self, grad=False):
        self._ensure_dev()
        cfg = self._dist_cfg()
        if cfg and grad:
            from .. import distchol
            return distchol.distributed_loglikelihood(self, True, nb=cfg.get('nb', 512), group=cfg.get('group'))
        lZ = C.c_double()
        dlZ = np.empty(self.nhyper) if grad else None
        _lib.check(self._dev.ctx, _lib.lib().pgp_exact_loglike(
            self._dev.handle, int(bool(grad)), C.byref(lZ), None if dlZ is None else _lib.ptr(dlZ)))
        return (lZ.value, dlZ) if grad else lZ.value

    def _full_posterior(self, X):
        """Joint posterior mean and covariance at the rows of X (exact.py:64-79)."""
        X = _lib.as_f64(X, 2)
        if self._X is None:
            return np.full(X.shape[0], self._mean), self._kernel.get(X)
        self._ensure_dev()
        mu, Sigma = np.empty(len(X)), np.empty((len(X), len(X)))
        _lib.check(self._dev.ctx, _lib.lib().pgp_exact_full_posterior(
            self._dev.handle, _lib.ptr(X), len(X), _lib.ptr(mu), _lib.ptr(Sigma)))
        return mu, Sigma

    def _marg_posterior(self, X, grad=False):
        X = _lib.as_f64(X, 2)
        if self._X is None:
            # prior: mean and k(x, x); constant mean and stationary kernel -> zero gradients (exact.py:83-104)
            out = (np.full(X.shape[0], self._mean), self._kernel.dget(X))
            return out + (np.zeros_like(X), np.zeros_like(X)) if grad else out
        if X.shape[1] != self._kernel.ndim:
            raise ValueError('test inputs have the wrong number of columns')
        self._ensure_dev()
        mu, s2 = np.empty(len(X)), np.empty(len(X))
        if not grad:
            _lib.check(self._dev.ctx, _lib.lib().pgp_exact_predict(
                self._dev.handle, _lib.ptr(X), len(X), _lib.ptr(mu), _lib.ptr(s2)))
            return mu, s2
        dmu, ds2 = np.empty(X.shape), np.empty(X.shape)
        _lib.check(self._dev.ctx, _lib.lib().pgp_exact_predict_grad(
            self._dev.handle, _lib.ptr(X), len(X), _lib.ptr(mu), _lib.ptr(s2), _lib.ptr(dmu), _lib.ptr(ds2)))
        return mu, s2, dmu, ds2
