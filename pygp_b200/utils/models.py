"""
Host-side parameter interface: same contract as pygp/utils/models.py:21-93
(`_params`, `get_hyper`, `set_hyper`, `copy(hyper)`, `printable`, `get_params`).
No arithmetic lives here.
"""

import abc
import copy as _copy

import numpy as np

__all__ = ['Parameterized', 'printable', 'get_params']


class Parameterized(abc.ABC):
    """Object described by a flat hyper-parameter vector."""

    @abc.abstractmethod
    def _params(self):
        """List of `(name, size, islog)` blocks in get_hyper() order."""

    @abc.abstractmethod
    def get_hyper(self):
        """Flat float64 vector of hyper-parameters (log space where islog)."""

    @abc.abstractmethod
    def set_hyper(self, hyper):
        """Assign the hyper-parameters from a flat vector."""

    def copy(self, hyper=None):
        """Deep copy (device state included); optionally with new hypers
        (utils/models.py:47-55)."""
        other = _copy.deepcopy(self)
        if hyper is not None:
            other.set_hyper(hyper)
        return other


def get_params(obj):
    """Yield `(name, slice, islog)` for each block of obj._params()
    (utils/models.py:83-93)."""
    start = 0
    for name, size, islog in obj._params():
        yield name, slice(start, start + size), islog
        start += size


def printable(cls):
    """Class decorator: repr as `Name(key=value, ...)` with log-space blocks
    shown exponentiated (utils/models.py:58-76)."""
    def __repr__(self):
        hyper = self.get_hyper()
        items = []
        for name, block, islog in get_params(self):
            value = hyper[block]
            value = value[0] if len(value) == 1 else value
            items.append('%s=%s' % (name, np.exp(value) if islog else value))
        return '%s(%s)' % (type(self).__name__, ', '.join(items))
    cls.__repr__ = __repr__
    return cls
