"""GP inference objects: same names as pygp.inference (DTC is not provided)."""
from .exact import ExactGP
from .basic import BasicGP

__all__ = ['ExactGP', 'BasicGP']
try:
    from .fitc import FITC
    __all__.append('FITC')
except ImportError:      # pragma: no cover
    pass
