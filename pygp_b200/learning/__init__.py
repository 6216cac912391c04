"""Hyper-parameter fitting: same entry points as pygp.learning."""
from .optimization import optimize
from .sampling import sample

__all__ = ['optimize', 'sample']
