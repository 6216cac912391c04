"""Covariance kernels: same constructors and methods as pygp.kernels."""
from .se import SE
from .matern import Matern
from .periodic import Periodic
from .rq import RQ
from ._combo import SumKernel, ProductKernel

__all__ = ['SE', 'Matern', 'Periodic', 'RQ']
