"""BasicGP convenience front-end (pygp/inference/basic.py:20-70): Gaussian
noise plus an SE or Matern kernel, with the short parameter names
`sn, sf, ell, mu` that `optimize` / `sample` priors key on."""

import numpy as np

from ..kernels import SE, Matern
from ..likelihoods import Gaussian
from ..utils.models import printable
from .exact import ExactGP

__all__ = ['BasicGP']

_MATERN = {'matern1': 1, 'matern3': 3, 'matern5': 5}


@printable
class BasicGP(ExactGP):
    def __init__(self, sn, sf, ell, mu=0, ndim=None, kernel='se'):
        if kernel == 'se':
            kern = SE(sf, ell, ndim)
        elif kernel in _MATERN:
            kern = Matern(sf, ell, _MATERN[kernel], ndim)
        else:
            raise ValueError('Unknown kernel type')
        super(BasicGP, self).__init__(Gaussian(sn), kern, mu)

    def _params(self):
        return [('sn', 1, True)] + self._kernel._params() + [('mu', 1, False)]

    @classmethod
    def from_gp(cls, gp):
        if not isinstance(gp._likelihood, Gaussian):
            raise ValueError('BasicGP instances must have Gaussian likelihood')
        if isinstance(gp._kernel, SE):
            kernel = 'se'
        elif isinstance(gp._kernel, Matern):
            kernel = 'matern%d' % gp._kernel._d
        else:
            raise ValueError('BasicGP instances must have a SE/Matern kernel')
        sn = np.sqrt(gp._likelihood.s2)
        sf = np.exp(gp._kernel._logsf)
        ell = np.exp(gp._kernel._logell)
        # (the reference drops `ndim` here, basic.py:66; keeping it preserves iso kernels)
        ndim = gp._kernel.ndim if gp._kernel._iso else None
        newgp = cls(sn, sf, ell, gp._mean, ndim, kernel)
        if gp.ndata > 0:
            X, y = gp.data
            newgp.add_data(X, y)
        return newgp
