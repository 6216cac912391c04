"""Helpers shared by the -m gpu tests: build product objects from oracle spec
tuples, and tolerances (BASELINE.json north_star: relative 1e-10 on lZ, mu, s2;
1e-8 on gradients)."""

import numpy as np

LZ_RTOL = 1e-10
PRED_RTOL = 1e-10
GRAD_RTOL = 1e-8


def product_kernel(spec):
    import pygp_b200 as pygp
    pk = pygp.kernels
    tag = spec[0]
    if tag == 'se':
        return pk.SE(*spec[1:])
    if tag == 'matern':
        return pk.Matern(*spec[1:])
    if tag == 'periodic':
        return pk.Periodic(*spec[1:])
    if tag == 'rq':
        return pk.RQ(*spec[1:])
    parts = [product_kernel(s) for s in spec[1:]]
    out = parts[0]
    for p in parts[1:]:
        out = (out + p) if tag == 'sum' else (out * p)
    return out


def assert_grad_close(g, g0, rtol=GRAD_RTOL):
    """1e-8 relative to the gradient's scale (entries near zero are compared
    against the largest entry, as approx-equality of a vector quantity)."""
    g, g0 = np.asarray(g), np.asarray(g0)
    scale = np.max(np.abs(g0))
    np.testing.assert_allclose(g, g0, rtol=rtol, atol=rtol*scale)


def assert_pred_close(mu, s2, mu0, s20, yscale=1.0, sf2=1.0, rtol=PRED_RTOL):
    """1e-10 relative; the absolute floor is 1e-10 of the output scale
    (|y| for the mean, sf^2 for the variance) -- SURVEY.md section 7.3 (1)."""
    np.testing.assert_allclose(mu, mu0, rtol=rtol, atol=rtol*yscale)
    np.testing.assert_allclose(s2, s20, rtol=rtol, atol=rtol*sf2)
