"""
pygp_b200 -- the exact-GP hot path of mwhoffman/pygp on NVIDIA B200.

Same public names as `pygp` (BasicGP, optimize, kernels, inference, learning,
likelihoods, meta, priors); every covariance / factorisation / solve runs in
hand-written sm_100a kernels behind the C ABI of include/pygp_b200.h.
"""

from . import utils
from . import kernels
from . import likelihoods
from . import priors
from . import inference
from . import learning
from . import meta

from .inference import BasicGP
from .learning import optimize

__all__ = ['BasicGP', 'optimize']
__version__ = '0.1.0'
