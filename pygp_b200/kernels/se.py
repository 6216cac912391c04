"""Squared-exponential kernel, iso or ARD (pygp/kernels/se.py:24-74):
k = sf^2 exp(-|x/ell - y/ell|^2 / 2).  Evaluated by gram.cu (PGP_SE)."""

import numpy as np

from .. import _lib
from ..utils.models import printable
from ._base import _ARDLeaf

__all__ = ['SE']


@printable
class SE(_ARDLeaf):
    def __init__(self, sf, ell, ndim=None):
        self._init_scales(sf, ell, ndim, 0)

    def _params(self):
        return [('sf', 1, True), ('ell', self.nhyper - 1, True)]

    def get_hyper(self):
        return np.r_[self._logsf, self._logell]

    def set_hyper(self, hyper):
        self._logsf = hyper[0]
        self._logell = hyper[1] if self._iso else np.array(hyper[1:], dtype=float)

    def _emit(self, parts, ops, offset):
        ops.append((_lib.OP_PUSH, len(parts)))
        parts.append((_lib.SE, int(self._iso), offset, self.nhyper))
        return offset + self.nhyper
