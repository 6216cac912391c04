"""GP inference objects: same names as pygp.inference."""
from .exact import ExactGP
from .basic import BasicGP
from .fitc import FITC
from .dtc import DTC

__all__ = ['ExactGP', 'BasicGP', 'FITC', 'DTC']
