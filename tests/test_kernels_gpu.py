"""Parity of Kernel.get / grad / dget / dgrad (CUDA, through the C ABI) with the
oracle and with the committed reference outputs.  Mirrors the reference's
tests/test_kernels.py (same kernels, same rng(0) inputs) and adds shapes that
cross the 64-wide device tiles."""

import numpy as np
import numpy.testing as nt
import pytest

from oracle.cases import KERNEL_CASES, kernel_inputs
from oracle.pygp_oracle import make_kernel
from gpu_util import product_kernel

pytestmark = pytest.mark.gpu

RT, AT = 1e-12, 1e-14


@pytest.mark.parametrize('name', sorted(KERNEL_CASES))
def test_vs_golden(name, golden):
    g = golden['kernels']
    k = product_kernel(KERNEL_CASES[name])
    x1, x2 = kernel_inputs(k.ndim)
    nt.assert_array_equal(k.get_hyper(), g[name + '/hyper'])
    nt.assert_allclose(k.get(x1, x2), g[name + '/get12'], rtol=RT, atol=AT)
    nt.assert_allclose(k.get(x1), g[name + '/get11'], rtol=RT, atol=AT)
    nt.assert_allclose(np.array(list(k.grad(x1, x2))), g[name + '/grad12'], rtol=RT, atol=AT)
    nt.assert_allclose(np.array(list(k.grad(x1))), g[name + '/grad11'], rtol=RT, atol=AT)
    nt.assert_allclose(k.dget(x1), g[name + '/dget'], rtol=RT)
    nt.assert_allclose(np.array(list(k.dgrad(x1))), g[name + '/dgrad'], rtol=RT, atol=AT)
    k2 = k.copy(g[name + '/hyper2'])
    nt.assert_allclose(k2.get(x1, x2), g[name + '/get12_h2'], rtol=RT, atol=AT)
    nt.assert_allclose(np.array(list(k2.grad(x1, x2))), g[name + '/grad12_h2'], rtol=RT, atol=AT)


@pytest.mark.parametrize('name', sorted(KERNEL_CASES))
def test_self_consistency(name):
    # reference tests/test_kernels.py:53-67,82-85: transpose, self, dgrad
    k = product_kernel(KERNEL_CASES[name])
    x1, x2 = kernel_inputs(k.ndim)
    nt.assert_allclose(k.get(x1, x2), k.get(x2, x1).T)
    G1 = np.array(list(k.grad(x1, x2)))
    G2 = np.array(list(k.grad(x2, x1))).swapaxes(1, 2)
    nt.assert_allclose(G1, G2)
    nt.assert_array_equal(k.get(x1), k.get(x1, x1))
    nt.assert_allclose(list(k.dgrad(x1)), [np.diag(_) for _ in k.grad(x1)])
    K = k.get(x1)
    nt.assert_array_equal(K, K.T)              # exact symmetry (direct differences)
    _ = repr(k), k.copy(), k(x1[0], x2[0])


@pytest.mark.parametrize('name,n1,n2', [('se_ard8', 130, 67), ('matern5_16', 65, 129), ('matern1_ard', 200, 64),
                                       ('se_plus_per', 257, 63), ('maunaloa', 100, 131), ('prod_mixed', 70, 70),
                                       ('rq_ard', 1, 300), ('periodic', 129, 1)])
def test_tiles_vs_oracle(name, n1, n2):
    spec = KERNEL_CASES[name]
    k, ok = product_kernel(spec), make_kernel(spec)
    rng = np.random.RandomState(3)
    x1, x2 = rng.rand(n1, k.ndim) * 2, rng.rand(n2, k.ndim) * 2
    x2[0] = x1[0]                               # an exact coincidence (r = 0 guards)
    nt.assert_allclose(k.get(x1, x2), ok.get(x1, x2), rtol=RT, atol=AT)
    nt.assert_allclose(k.get(x1), ok.get(x1), rtol=RT, atol=AT)
    nt.assert_allclose(np.array(list(k.grad(x1, x2))), np.array(ok.grad(x1, x2)), rtol=RT, atol=AT)
    nt.assert_allclose(np.array(list(k.grad(x2))), np.array(ok.grad(x2)), rtol=RT, atol=AT)


def test_empty_and_errors():
    import pygp_b200 as pygp
    k = pygp.kernels.SE(1.0, [0.5, 0.5])
    assert k.get(np.zeros((0, 2)), np.zeros((4, 2))).shape == (0, 4)
    assert k.dget(np.zeros((0, 2))).shape == (0,)
    with pytest.raises(ValueError):
        k.get(np.zeros((3, 3)))


@pytest.mark.parametrize('name', sorted(KERNEL_CASES))
def test_input_gradients_vs_golden(name, golden):
    """Kernel.gradx / grady (reference tests/test_kernels.py:97-128) against the
    reference's own outputs; grady == -gradx, and gradx(X) == gradx(X, X)."""
    g = golden['kernels']
    k = product_kernel(KERNEL_CASES[name])
    x1, x2 = kernel_inputs(k.ndim)
    nt.assert_allclose(k.gradx(x1, x2), g[name + '/gradx12'], rtol=1e-12, atol=1e-13)
    nt.assert_allclose(k.grady(x1, x2), g[name + '/grady12'], rtol=1e-12, atol=1e-13)
    nt.assert_allclose(k.gradx(x1), g[name + '/gradx11'], rtol=1e-12, atol=1e-13)
    # gradxy: SE and its sums / products (tests/test_kernels.py:130-146); NotImplementedError otherwise
    if name + '/gradxy12' in g.files:
        nt.assert_allclose(k.gradxy(x1, x2), g[name + '/gradxy12'], rtol=1e-11, atol=1e-12)
        nt.assert_allclose(k.gradxy(x1), make_kernel(KERNEL_CASES[name]).gradxy(x1), rtol=1e-11, atol=1e-12)
    else:
        with pytest.raises(NotImplementedError):
            k.gradxy(x1, x2)


@pytest.mark.parametrize('name', ['matern3_ard', 'maunaloa', 'prod_mixed', 'periodic'])
def test_input_gradients_tiles_vs_oracle(name):
    """ragged sizes crossing the 64 x 64 tile, and the finite-difference check
    of the reference's test_gradx."""
    import scipy.optimize as spop
    spec = KERNEL_CASES[name]
    k, ok = product_kernel(spec), make_kernel(spec)
    rng = np.random.RandomState(4)
    x1, x2 = rng.rand(70, k.ndim), rng.rand(131, k.ndim)
    G = k.gradx(x1, x2)
    nt.assert_allclose(G, ok.gradx(x1, x2), rtol=1e-12, atol=1e-13)
    f = lambda a, b: ok.get(a[None], b[None])[0, 0]
    G2 = np.array([spop.approx_fprime(a, f, 1e-8, b) for a in x1[:3] for b in x2[:4]]).reshape(3, 4, -1)
    nt.assert_allclose(G[:3, :4], G2, rtol=1e-5, atol=1e-5)
