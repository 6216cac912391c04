"""Periodic kernel on one input dimension (pygp/kernels/periodic.py:23-82):
k = sf^2 exp(-2 sin^2(|x - y| pi / p) / ell^2).  Evaluated by gram.cu."""

import numpy as np

from .. import _lib
from ..utils.models import printable
from ._base import RealKernel

__all__ = ['Periodic']


@printable
class Periodic(RealKernel):
    def __init__(self, sf, ell, p):
        self._logsf = np.log(float(sf))
        self._logell = np.log(float(ell))
        self._logp = np.log(float(p))
        self.ndim = 1
        self.nhyper = 3

    def _params(self):
        return [('sf', 1, True), ('ell', 1, True), ('p', 1, True)]

    def get_hyper(self):
        return np.r_[self._logsf, self._logell, self._logp]

    def set_hyper(self, hyper):
        self._logsf, self._logell, self._logp = hyper[0], hyper[1], hyper[2]

    def _emit(self, parts, ops, offset):
        ops.append((_lib.OP_PUSH, len(parts)))
        parts.append((_lib.PERIODIC, 1, offset, 3))
        return offset + 3
