"""Meta-models marginalising the hypers (meta.SMC is not provided this round)."""
from .mcmc import MCMC

__all__ = ['MCMC']
