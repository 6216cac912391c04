"""Meta-models marginalising the hypers: same names as pygp.meta."""
from .mcmc import MCMC
from .smc import SMC

__all__ = ['MCMC', 'SMC']
