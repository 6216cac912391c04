"""Rational-quadratic kernel, iso or ARD (pygp/kernels/rq.py:23-93):
k = sf^2 (1 + |x/ell - y/ell|^2 / (2 alpha))^-alpha.  Evaluated by gram.cu."""

import numpy as np

from .. import _lib
from ..utils.models import printable
from ._base import _ARDLeaf

__all__ = ['RQ']


@printable
class RQ(_ARDLeaf):
    def __init__(self, sf, ell, alpha, ndim=None):
        self._init_scales(sf, ell, ndim, 1)
        self._logalpha = np.log(float(alpha))

    def _params(self):
        return [('sf', 1, True), ('ell', self.nhyper - 2, True), ('alpha', 1, True)]

    def get_hyper(self):
        return np.r_[self._logsf, self._logell, self._logalpha]

    def set_hyper(self, hyper):
        self._logsf = hyper[0]
        self._logell = hyper[1] if self._iso else np.array(hyper[1:-1], dtype=float)
        self._logalpha = hyper[-1]

    def _emit(self, parts, ops, offset):
        ops.append((_lib.OP_PUSH, len(parts)))
        parts.append((_lib.RQ, int(self._iso), offset, self.nhyper))
        return offset + self.nhyper
