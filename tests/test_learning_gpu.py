"""optimize / sample / meta.MCMC on top of the device path (reference
tests/test_learning.py:24-36, tests/test_meta.py:24-42) and the batched
small-N path against a loop of single models."""

import os

import numpy as np
import numpy.testing as nt
import pytest

from oracle.pygp_oracle import make_kernel, OExactGP, synthetic_problem
from gpu_util import product_kernel, assert_grad_close, assert_pred_close, LZ_RTOL

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_optimization_constraint_and_trajectory():
    import pygp_b200 as pygp
    data = np.load(os.path.join(HERE, 'golden', 'xy.npz'))
    X, y = data['X'], data['y']
    gp = pygp.BasicGP(sn=.1, sf=1, ell=.1, mu=0)
    gp.add_data(X, y)
    lZ0 = gp.loglikelihood()
    # every objective evaluation of the optimiser's trajectory matches the oracle
    ogp = OExactGP(.1, make_kernel(('se', 1, .1)), 0.)
    ogp.add_data(X, y)
    import scipy.optimize as so
    traj = []
    h0 = ogp.get_hyper()

    def obj(x):
        h = h0.copy()
        h[1:] = x
        ogp.set_hyper(h)
        traj.append(h.copy())
        lZ, dlZ = ogp.loglikelihood(True)
        return -lZ, -dlZ[1:]
    so.fmin_l_bfgs_b(obj, h0[1:])
    for h in traj[:25]:
        gp.set_hyper(h)
        ogp.set_hyper(h)
        lZ, dlZ = gp.loglikelihood(True)
        olZ, odlZ = ogp.loglikelihood(True)
        # the optimiser visits badly conditioned hypers: two correct FP64
        # implementations then differ by ~eps * cond(K~), so the 1e-10 / 1e-8 bars
        # are widened by that factor where it exceeds them
        cond = np.linalg.cond(ogp._R)**2
        slack = max(1.0, 20*np.finfo(float).eps*cond/LZ_RTOL)
        nt.assert_allclose(lZ, olZ, rtol=LZ_RTOL*slack, atol=1e-10*slack)
        assert_grad_close(dlZ, odlZ, rtol=1e-8*slack)
    gp.set_hyper(h0)
    pygp.optimize(gp, {'sn': None})
    nt.assert_equal(gp.get_hyper()[0], np.log(0.1))       # tests/test_learning.py:36
    assert gp.loglikelihood() > lZ0


def _mcmc_setup(n=10):
    import pygp_b200 as pygp
    ndim = 2
    prior = {'sn': pygp.priors.Uniform(0.01, 1.0), 'sf': pygp.priors.Uniform(0.01, 5.0),
             'ell': pygp.priors.Uniform([0.01]*ndim, [1.0]*ndim),
             'mu': pygp.priors.Uniform([-2.0], [2.0])}
    model = pygp.meta.MCMC(pygp.BasicGP(0.5, 1, [1]*ndim), prior, n=n, burn=0, rng=0)
    rng = np.random.RandomState(0)
    model.add_data(rng.rand(10, ndim), rng.rand(10))
    return model, rng.rand(10, ndim)


def test_mcmc_batched_equals_loop():
    model, Xs = _mcmc_setup()
    mu, s2 = model.posterior(Xs)                              # one batched device call
    parts = [m.posterior(Xs) for m in model]                  # n single models
    mu_ = np.array([p[0] for p in parts])
    s2_ = np.array([p[1] for p in parts])
    mu0 = mu_.mean(0)
    nt.assert_allclose(mu, mu0, rtol=1e-12)
    nt.assert_allclose(s2, np.mean(s2_ + (mu_ - mu0)**2, axis=0), rtol=1e-11)
    assert model.ndata == 10 and len(list(model)) == 10
    # and every component against the oracle
    X, y = model.data
    for m, (mu_i, s2_i) in zip(model, parts):
        h = m.get_hyper()
        og = OExactGP(np.exp(h[0]), make_kernel(('se', np.exp(h[1]), list(np.exp(h[2:4])))), h[4])
        og.add_data(X, y)
        omu, os2 = og.posterior(Xs)
        assert_pred_close(mu_i, s2_i, omu, os2)


def test_sampler_chain_matches_oracle_chain():
    """The slice sampler is driven only by loglikelihood(): with the same rng the
    device-backed chain reproduces a chain run on the oracle."""
    import pygp_b200 as pygp
    from pygp_b200.learning.sampling import sample
    rng = np.random.RandomState(4)
    X, y = rng.rand(25, 1), rng.rand(25)
    prior = {'sn': pygp.priors.Uniform(0.01, 1.0), 'sf': pygp.priors.Uniform(0.01, 5.0),
             'ell': pygp.priors.Uniform([0.01], [1.0]), 'mu': None}
    gp = pygp.BasicGP(0.5, 1, [0.3])
    gp.add_data(X, y)
    H = sample(gp, prior, 8, rng=7)

    class Shim(OExactGP):                                  # oracle with BasicGP's names
        def _params(self):
            return [('sn', 1, True), ('sf', 1, True), ('ell', 1, True), ('mu', 1, False)]
    og = Shim(0.5, make_kernel(('se', 1, [0.3])), 0.)
    og.add_data(X, y)
    H0 = sample(og, prior, 8, rng=7)
    nt.assert_allclose(H, H0, rtol=1e-7, atol=1e-9)


@pytest.mark.parametrize('spec,n,d,B', [(('se', 1.0, [0.8]*8), 300, 8, 37), (('matern', 1.0, [0.7, 0.9], 3), 129, 2, 5)])
def test_batched_loglike_vs_single(spec, n, d, B):
    import ctypes as C
    import pygp_b200 as pygp
    from pygp_b200 import _lib
    X, y, Xs = synthetic_problem(n, d, 40)
    k = product_kernel(spec)
    gp = pygp.inference.ExactGP(pygp.likelihoods.Gaussian(0.1), k, 0.0)
    gp.add_data(X, y)
    h0 = gp.get_hyper()
    H = h0 + np.random.RandomState(2).uniform(-0.5, 0.5, (B, len(h0)))
    lZ, info = np.empty(B), np.zeros(B, dtype=np.int32)
    ctx = _lib.context()
    dlZ = np.empty((B, len(h0)))
    _lib.check(ctx, _lib.lib().pgp_batched_loglike(ctx.handle, k._spec(), _lib.ptr(X), _lib.ptr(y), n,
                                                   _lib.ptr(H), B, _lib.ptr(lZ), _lib.ptr(dlZ),
                                                   info.ctypes.data_as(C.POINTER(C.c_int32))))
    assert not info.any()
    lZ_only = np.empty(B)                          # dlZ = NULL: likelihood only, same values
    _lib.check(ctx, _lib.lib().pgp_batched_loglike(ctx.handle, k._spec(), _lib.ptr(X), _lib.ptr(y), n,
                                                   _lib.ptr(H), B, _lib.ptr(lZ_only), None,
                                                   info.ctypes.data_as(C.POINTER(C.c_int32))))
    nt.assert_array_equal(lZ_only, lZ)
    mu, s2 = np.empty((B, 40)), np.empty((B, 40))
    _lib.check(ctx, _lib.lib().pgp_batched_predict(ctx.handle, k._spec(), _lib.ptr(X), _lib.ptr(y), n,
                                                   _lib.ptr(H), B, _lib.ptr(Xs), 40, _lib.ptr(mu), _lib.ptr(s2),
                                                   info.ctypes.data_as(C.POINTER(C.c_int32))))
    for b in range(B):
        gp.set_hyper(H[b])
        l1, g1 = gp.loglikelihood(True)
        nt.assert_allclose(lZ[b], l1, rtol=1e-13)
        nt.assert_allclose(dlZ[b], g1, rtol=1e-10, atol=1e-10*np.abs(g1).max())      # batched gradient == pgp_exact_loglike
        m1, v1 = gp.posterior(Xs)
        nt.assert_allclose(mu[b], m1, rtol=1e-12, atol=1e-13)
        nt.assert_allclose(s2[b], v1, rtol=1e-12, atol=1e-14)
    og = OExactGP(0.1, make_kernel(spec), 0.0)
    og.add_data(X, y)
    for b in (0, B-1):
        og.set_hyper(H[b])
        nt.assert_allclose(lZ[b], og.loglikelihood(), rtol=LZ_RTOL)


def test_sharding_single_rank_uses_device_path():
    """pygp_b200.sharding without a process group (world size 1) runs the real
    batched device calls; the N > 1 logic is covered on CPU (tests/test_sharding.py)
    and on GPUs by tools/dist_check.py."""
    import pygp_b200 as pygp
    from pygp_b200 import sharding
    X, y, Xs = synthetic_problem(150, 3, 17)
    spec = ('se', 1.0, [0.5, 0.6, 0.7])
    gp = pygp.inference.ExactGP(pygp.likelihoods.Gaussian(0.1), product_kernel(spec), 0.0)
    gp.add_data(X, y)
    H = gp.get_hyper() + np.random.RandomState(2).uniform(-0.3, 0.3, size=(5, gp.nhyper))
    lZ = sharding.sharded_batched_loglike(gp, H)
    ref = []
    for h in H:
        o = OExactGP(0.1, make_kernel(spec), 0.0)
        o.add_data(X, y)
        o.set_hyper(h)
        ref.append((o.loglikelihood(),) + o.posterior(Xs))
    nt.assert_allclose(lZ, [r[0] for r in ref], rtol=LZ_RTOL)
    lZg, dlZg = sharding.sharded_batched_loglike(gp, H, grad=True)     # batched gradient vs the oracle, one by one
    nt.assert_array_equal(lZg, lZ)
    for h, g in zip(H, dlZg):
        o = OExactGP(0.1, make_kernel(spec), 0.0)
        o.add_data(X, y)
        o.set_hyper(h)
        assert_grad_close(g, o.loglikelihood(True)[1])
    mu, s2 = sharding.sharded_mixture_posterior(gp, H, Xs)
    mu_ = np.array([r[1] for r in ref])
    s2_ = np.array([r[2] for r in ref])
    mu0 = mu_.mean(0)
    nt.assert_allclose(mu, mu0, rtol=1e-10, atol=1e-10)
    nt.assert_allclose(s2, np.mean(s2_ + (mu_ - mu0)**2, axis=0), rtol=1e-9, atol=1e-10)
    m2, v2 = sharding.sharded_posterior(gp, Xs)
    m0, v0 = gp.posterior(Xs)
    nt.assert_array_equal(m2, m0)
    nt.assert_array_equal(v2, v0)


def test_mcmc_posterior_input_gradients():
    """meta.MCMC.posterior(X, grad=True) (mcmc.py:84-93; reference tests/test_meta.py:44-54):
    finite differences of the mixture mean and variance."""
    import scipy.optimize as spop
    import pygp_b200 as pygp
    rng = np.random.RandomState(1)
    X = rng.rand(10, 2)
    y = rng.rand(10)
    gp = pygp.BasicGP(0.5, 1, [1, 1])
    gp.add_data(X, y)
    priors = {'sn': pygp.priors.Uniform(0.01, 1.0), 'sf': pygp.priors.Uniform(0.01, 5.0),
              'ell': pygp.priors.Uniform([0.01, 0.01], [1.0, 1.0]), 'mu': pygp.priors.Uniform(-2, 2)}
    model = pygp.meta.MCMC(gp, priors, n=6, burn=0, rng=0)
    Xq = rng.rand(4, 2)
    mu, s2, dmu, ds2 = model.posterior(Xq, grad=True)
    f_mu = lambda x: model.posterior(x[None])[0][0]
    f_s2 = lambda x: model.posterior(x[None])[1][0]
    nt.assert_allclose(dmu, [spop.approx_fprime(x, f_mu, 1e-7) for x in Xq], rtol=1e-5, atol=1e-5)
    nt.assert_allclose(ds2, [spop.approx_fprime(x, f_s2, 1e-7) for x in Xq], rtol=1e-4, atol=1e-5)


def test_smc_particles():
    """meta.SMC (pygp/meta/smc.py:53-150; the reference only smoke-tests it,
    tests/test_plotting.py:62-63): weights stay normalised, every particle's
    incrementally grown factor equals a from-scratch model at its hypers, the
    mixture posterior is the weighted moment match of the particles, and its
    input-gradients agree with finite differences."""
    import scipy.optimize as spop
    import pygp_b200 as pygp
    rng = np.random.RandomState(3)
    X = rng.rand(12, 2)
    y = np.sin(3*X.sum(1)) + 0.1*rng.randn(12)
    priors = {'sn': pygp.priors.Uniform(0.05, 1.0), 'sf': pygp.priors.Uniform(0.2, 5.0),
              'ell': pygp.priors.Uniform([0.05, 0.05], [2.0, 2.0]), 'mu': pygp.priors.Uniform(-2, 2)}
    model = pygp.meta.SMC(pygp.BasicGP(0.5, 1, [1, 1]), priors, n=6, rng=0)
    assert model.ndata == 0
    model.add_data(X[:7], y[:7])
    model.add_data(X[7:], y[7:])
    assert model.ndata == 12
    nt.assert_allclose(np.exp(model._logweights).sum(), 1.0, rtol=1e-12)
    for p in model:
        fresh = pygp.BasicGP(0.5, 1, [1, 1])
        fresh.set_hyper(p.get_hyper())
        fresh.add_data(*p.data)
        nt.assert_allclose(p.loglikelihood(), fresh.loglikelihood(), rtol=1e-9)
    Xq = rng.rand(3, 2)
    mu, s2, dmu, ds2 = model.posterior(Xq, grad=True)
    w = np.exp(model._logweights)
    parts = [p.posterior(Xq) for p in model]
    mu_ = np.array([p[0] for p in parts])
    s2_ = np.array([p[1] for p in parts])
    nt.assert_allclose(mu, w @ mu_, rtol=1e-12)
    nt.assert_allclose(s2, w @ (s2_ + (mu_ - mu)**2), rtol=1e-12)
    f_mu = lambda x: model.posterior(x[None])[0][0]
    f_s2 = lambda x: model.posterior(x[None])[1][0]
    nt.assert_allclose(dmu, [spop.approx_fprime(x, f_mu, 1e-7) for x in Xq], rtol=1e-5, atol=1e-5)
    nt.assert_allclose(ds2, [spop.approx_fprime(x, f_s2, 1e-7) for x in Xq], rtol=1e-4, atol=1e-5)
    # a model that already holds data replays it (smc.py:62-76)
    gp = pygp.BasicGP(0.5, 1, [1, 1])
    gp.add_data(X[:5], y[:5])
    m2 = pygp.meta.SMC(gp, priors, n=4, rng=1)
    assert m2.ndata == 5 and gp.ndata == 5
