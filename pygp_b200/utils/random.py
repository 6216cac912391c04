"""Random-state helper (the reference takes it from the un-vendored mwhutils)."""

import numpy as np

__all__ = ['rstate']


def rstate(rng=None):
    """None -> numpy's global state; int -> seeded RandomState; RandomState -> itself."""
    if rng is None:
        return np.random.mtrand._rand
    if isinstance(rng, np.random.RandomState):
        return rng
    if isinstance(rng, (int, np.integer)):
        return np.random.RandomState(int(rng))
    raise ValueError('unknown seed: %r' % (rng,))
