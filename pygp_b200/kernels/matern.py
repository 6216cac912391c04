"""Matern kernels with nu = d/2, d in {1, 3, 5}, iso or ARD
(pygp/kernels/matern.py:24-98).  Evaluated by gram.cu (PGP_MATERN{1,3,5})."""

import numpy as np

from .. import _lib
from ..utils.models import printable
from ._base import _ARDLeaf

__all__ = ['Matern']

_TYPE = {1: _lib.MATERN1, 3: _lib.MATERN3, 5: _lib.MATERN5}


@printable
class Matern(_ARDLeaf):
    def __init__(self, sf, ell, d=3, ndim=None):
        self._init_scales(sf, ell, ndim, 0)
        self._d = d
        if d not in _TYPE:
            raise ValueError('d must be one of 1, 3, or 5')

    def _params(self):
        return [('sf', 1, True), ('ell', self.nhyper - 1, True)]

    def get_hyper(self):
        return np.r_[self._logsf, self._logell]

    def set_hyper(self, hyper):
        self._logsf = hyper[0]
        self._logell = hyper[1] if self._iso else np.array(hyper[1:], dtype=float)

    def _emit(self, parts, ops, offset):
        ops.append((_lib.OP_PUSH, len(parts)))
        parts.append((_TYPE[self._d], int(self._iso), offset, self.nhyper))
        return offset + self.nhyper
