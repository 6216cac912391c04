"""Sum and product composites (pygp/kernels/_combo.py:54-160,
_real.py:76-117).  A composite owns deep copies of its parts and concatenates
their hyper vectors; on the device it is a postfix program over leaf values
(spec.cuh), so nested trees cost one kernel launch like a leaf."""

import numpy as np

from .. import _lib
from ._base import RealKernel

__all__ = ['ComboKernel', 'SumKernel', 'ProductKernel', 'combine']


def combine(cls, *parts):
    """Flatten operands that already are a `cls` composite (associativity)."""
    flat = []
    for part in parts:
        flat.extend(part._parts if isinstance(part, cls) else [part])
    return flat


class ComboKernel(RealKernel):
    _op = None
    _verb = 'combine'

    def __init__(self, *parts):
        if not (all(isinstance(p, RealKernel) for p in parts)
                and all(p.ndim == parts[0].ndim for p in parts)):
            raise ValueError('cannot %s mismatched kernels' % self._verb)
        self._parts = [p.copy() for p in parts]
        self.nhyper = sum(p.nhyper for p in self._parts)
        self.ndim = self._parts[0].ndim

    def __repr__(self):
        head = type(self).__name__ + '('
        body = ',\n'.join(repr(p) for p in self._parts) + ')'
        return ('\n' + ' ' * len(head)).join((head + body).splitlines())

    def _params(self):
        # flat list over the leaves, named part<i>.<param> (_combo.py:74-88)
        out, leaves = [], []
        stack = list(reversed(self._parts))
        while stack:
            part = stack.pop()
            if isinstance(part, ComboKernel):
                stack.extend(reversed(part._parts))
            else:
                leaves.append(part)
        for i, leaf in enumerate(leaves):
            out.extend(('part%d.%s' % (i, p[0]),) + tuple(p[1:]) for p in leaf._params())
        return out

    def get_hyper(self):
        return np.hstack([p.get_hyper() for p in self._parts])

    def set_hyper(self, hyper):
        a = 0
        for p in self._parts:
            p.set_hyper(hyper[a:a + p.nhyper])
            a += p.nhyper

    def _emit(self, parts, ops, offset):
        for p in self._parts:
            offset = p._emit(parts, ops, offset)
        ops.append((self._op, len(self._parts)))
        return offset


class SumKernel(ComboKernel):
    _op = _lib.OP_SUM
    _verb = 'add'


class ProductKernel(ComboKernel):
    _op = _lib.OP_PROD
    _verb = 'multiply'
