"""
Kernel interface (pygp/kernels/_base.py:22-64, _real.py:29-73) on top of the
C ABI.  A kernel object holds only its log-space hyper-parameters and a
description of its structure; `get`, `grad`, `dget`, `dgrad` serialise that
description into a `pgp_kernel_spec` (leaves + postfix program) and run on the
device.  No covariance arithmetic happens in Python.
"""

import abc

import numpy as np

from .. import _lib
from ..utils.models import Parameterized

__all__ = ['Kernel', 'RealKernel']


class Kernel(Parameterized):
    def __call__(self, x1, x2):
        return self.get(x1[None], x2[None])[0]

    # -- structure -> pgp_kernel_spec ---------------------------------------
    @abc.abstractmethod
    def _emit(self, parts, ops, offset):
        """Append this kernel's leaves and postfix ops; return the next hyper offset."""

    def _spec(self):
        parts, ops = [], []
        end = self._emit(parts, ops, 0)
        if len(parts) > _lib.MAX_PARTS or len(ops) > _lib.MAX_OPS:
            raise ValueError('composite kernel too large for the device descriptor '
                             '(%d leaves, %d ops)' % (len(parts), len(ops)))
        if self.ndim > _lib.MAX_DIM or end > _lib.MAX_HYPER:
            raise ValueError('too many input dimensions / hyper-parameters for the device descriptor')
        spec = _lib.KernelSpec()
        spec.ndim, spec.nhyper = int(self.ndim), int(end)
        spec.n_parts, spec.n_ops = len(parts), len(ops)
        for i, (typ, iso, off, nh) in enumerate(parts):
            spec.parts[i] = _lib.Part(typ, iso, off, nh)
        for i, (op, arg) in enumerate(ops):
            spec.ops[i] = _lib.Op(op, arg)
        return spec

    def _hyp(self):
        return _lib.as_f64(self.get_hyper())

    def _inputs(self, X1, X2):
        X1 = _lib.as_f64(X1, 2)
        if X1.shape[1] != self.ndim:
            raise ValueError('inputs have %d columns, kernel expects %d' % (X1.shape[1], self.ndim))
        if X2 is not None:
            X2 = _lib.as_f64(X2, 2)
            if X2.shape[1] != self.ndim:
                raise ValueError('inputs have %d columns, kernel expects %d' % (X2.shape[1], self.ndim))
        return X1, X2

    # -- Kernel interface ------------------------------------------------------
    def get(self, X1, X2=None):
        """Covariances between X1 and X2 (X1 itself if X2 is None)."""
        X1, X2 = self._inputs(X1, X2)
        n1, n2 = len(X1), (len(X1) if X2 is None else len(X2))
        out = np.empty((n1, n2))
        ctx, L, hyp, spec = _lib.context(), _lib.lib(), self._hyp(), self._spec()
        _lib.check(ctx, L.pgp_gram(ctx.handle, spec, _lib.ptr(hyp), _lib.ptr(X1), n1,
                                   None if X2 is None else _lib.ptr(X2), n2, _lib.ptr(out)))
        return out

    def grad(self, X1, X2=None):
        """Iterator over d get / d hyper_k, one (n1, n2) matrix per hyper."""
        X1, X2 = self._inputs(X1, X2)
        n1, n2 = len(X1), (len(X1) if X2 is None else len(X2))
        ctx, L, hyp, spec = _lib.context(), _lib.lib(), self._hyp(), self._spec()
        for k in range(self.nhyper):
            out = np.empty((n1, n2))
            _lib.check(ctx, L.pgp_gram_grad(ctx.handle, spec, _lib.ptr(hyp), _lib.ptr(X1), n1,
                                            None if X2 is None else _lib.ptr(X2), n2, k, _lib.ptr(out)))
            yield out

    def dget(self, X):
        """Self-covariances k(x_i, x_i)."""
        X, _ = self._inputs(X, None)
        out = np.empty(len(X))
        ctx, L, hyp, spec = _lib.context(), _lib.lib(), self._hyp(), self._spec()
        _lib.check(ctx, L.pgp_dget(ctx.handle, spec, _lib.ptr(hyp), _lib.ptr(X), len(X), _lib.ptr(out)))
        return out

    def dgrad(self, X):
        """Iterator over the hyper-gradients of the self-covariances."""
        X, _ = self._inputs(X, None)
        out = np.empty((self.nhyper, len(X)))
        ctx, L, hyp, spec = _lib.context(), _lib.lib(), self._hyp(), self._spec()
        _lib.check(ctx, L.pgp_dgrad(ctx.handle, spec, _lib.ptr(hyp), _lib.ptr(X), len(X), _lib.ptr(out)))
        for row in out:
            yield row

    @abc.abstractmethod
    def transform(self, X):
        """Format the inputs X as arrays."""


class RealKernel(Kernel):
    """Kernel over real vectors; `+` and `*` build Sum/Product composites
    (flattening same-type nesting, _real.py:32-36, _combo.py:151-160)."""

    def __add__(self, other):
        from ._combo import SumKernel, combine
        return SumKernel(*combine(SumKernel, self, other))

    def __mul__(self, other):
        from ._combo import ProductKernel, combine
        return ProductKernel(*combine(ProductKernel, self, other))

    def transform(self, X):
        return np.array(X, ndmin=2, dtype=float)

    # Input-gradients: on the device as well.
    def _gradx(self, X1, X2, wrt_y):
        X1, X2 = self._inputs(X1, X2)
        n1, n2 = len(X1), (len(X1) if X2 is None else len(X2))
        out = np.empty((n1, n2, self.ndim))
        ctx, L, hyp, spec = _lib.context(), _lib.lib(), self._hyp(), self._spec()
        _lib.check(ctx, L.pgp_gram_gradx(ctx.handle, spec, _lib.ptr(hyp), _lib.ptr(X1), n1,
                                         None if X2 is None else _lib.ptr(X2), n2, wrt_y, _lib.ptr(out)))
        return out

    def gradx(self, X1, X2=None):
        """d k(x1, x2) / d x1: (n1, n2, ndim)."""
        return self._gradx(X1, X2, 0)

    def grady(self, X1, X2=None):
        """d k(x1, x2) / d x2: (n1, n2, ndim)."""
        return self._gradx(X1, X2, 1)

    def gradxy(self, X1, X2=None):
        """d2 k(x1, x2) / d x1 d x2: (n1, n2, ndim, ndim).  Defined, as in the
        reference, for SE kernels and their sums / products only."""
        spec = self._spec()
        if any(spec.parts[i].type != _lib.SE for i in range(spec.n_parts)):
            raise NotImplementedError
        X1, X2 = self._inputs(X1, X2)
        n1, n2 = len(X1), (len(X1) if X2 is None else len(X2))
        out = np.empty((n1, n2, self.ndim, self.ndim))
        ctx, L, hyp = _lib.context(), _lib.lib(), self._hyp()
        _lib.check(ctx, L.pgp_gram_gradxy(ctx.handle, spec, _lib.ptr(hyp), _lib.ptr(X1), n1,
                                          None if X2 is None else _lib.ptr(X2), n2, _lib.ptr(out)))
        return out

    def sample_spectrum(self, N, rng=None):
        raise NotImplementedError('sample_spectrum is outside the B200 hot path')


class _ARDLeaf(RealKernel):
    """Shared constructor logic of SE / Matern / RQ: scalar `ell` + `ndim`
    means an isotropic kernel, a vector `ell` means ARD (se.py:26-38)."""

    def _init_scales(self, sf, ell, ndim, extra):
        self._logsf = np.log(float(sf))
        self._logell = np.log(np.array(ell, dtype=float))
        self._iso = False
        self.ndim = int(np.size(self._logell))
        if ndim is not None:
            if np.size(self._logell) != 1:
                raise ValueError('ndim only usable with scalar lengthscales')
            self._logell = float(self._logell)
            self._iso = True
            self.ndim = int(ndim)
        else:
            self._logell = np.atleast_1d(self._logell)
        self.nhyper = 1 + int(np.size(self._logell)) + extra

    @property
    def _nell(self):
        return 1 if self._iso else self.ndim
