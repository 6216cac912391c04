"""Parameter plumbing shared by kernels, likelihoods and GP objects."""
from . import models
from . import random
