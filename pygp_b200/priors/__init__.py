"""Hyper-parameter priors (host side, O(nhyper) scalar work)."""
from .priors import Uniform, Gaussian, Gamma, LogNormal, Horseshoe

__all__ = ['Uniform', 'Gaussian', 'Gamma', 'LogNormal', 'Horseshoe']
