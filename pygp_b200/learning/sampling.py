"""
Slice sampling of GP hyper-parameters (pygp/learning/sampling.py:24-146).

The chain is sequential by construction; every `logprob` evaluation is one
device factorisation (`set_hyper` -> pgp_exact_update) plus the cached lZ.
"""

import numpy as np

from ..utils.models import get_params
from ..utils.random import rstate

__all__ = ['sample']


def _slice_sample(logprob, x0, sigma=1.0, step_out=True, max_steps_out=1000, rng=None):
    """One random-direction slice-sampling step with step-out and shrinkage
    (sampling.py:24-74; same order of rng draws)."""
    rng = rstate(rng)
    direction = rng.randn(x0.shape[0])
    direction = direction / np.sqrt(np.sum(direction**2))

    def along(z):
        return logprob(direction*z + x0)

    upper = sigma*rng.rand()
    lower = upper - sigma
    level = np.log(rng.rand()) + along(0.0)

    if step_out:
        steps = 0
        while along(lower) > level and steps < max_steps_out:
            steps += 1
            lower -= sigma
        steps = 0
        while along(upper) > level and steps < max_steps_out:
            steps += 1
            upper += sigma

    while True:
        z = (upper - lower)*rng.rand() + lower
        value = along(z)
        if np.isnan(value):
            raise Exception('Slice sampler got a NaN')
        if value > level:
            return z*direction + x0
        if z < 0:
            lower = z
        elif z > 0:
            upper = z
        else:
            raise Exception('Slice sampler shrank to zero!')


def sample(gp, priors, n, raw=True, rng=None):
    """Draw n hyper-parameter samples; `raw=False` returns model copies."""
    rng = rstate(rng)
    priors = dict(priors)
    active = np.ones(gp.nhyper, dtype=bool)
    logged = np.ones(gp.nhyper, dtype=bool)
    blocks = []
    for name, block, islog in get_params(gp):
        logged[block] = islog
        if name in priors and priors[name] is None:
            active[block] = False
        elif name in priors:
            blocks.append((block, priors[name]))

    hyper0 = gp.get_hyper()
    hyper0[logged] = np.exp(hyper0[logged])

    def logprob(x):
        hyper = hyper0.copy()
        hyper[active] = x
        total = 0
        for block, prior in blocks:
            total += prior.logprior(hyper[block])
            if np.isinf(total):
                return total
        hyper[logged] = np.log(hyper[logged])
        gp.set_hyper(hyper)
        return total + gp.loglikelihood()

    hypers = np.tile(hyper0, (n, 1))
    x = hyper0.copy()[active]
    for i in range(n):
        x = _slice_sample(logprob, x, rng=rng)
        hypers[i][active] = x
    hypers[:, logged] = np.log(hypers[:, logged])
    gp.set_hyper(hypers[-1])
    return hypers if raw else [gp.copy(h) for h in hypers]
