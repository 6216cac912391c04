"""Likelihood models (host side: they only supply the noise variance)."""
from .gaussian import Gaussian, Likelihood, RealLikelihood

__all__ = ['Gaussian', 'Likelihood', 'RealLikelihood']
