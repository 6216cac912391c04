"""Parity of the FITC hot path (pgp_fitc_update / _loglike / _predict,
pygp/inference/fitc.py:66-232) with the oracle, the committed reference
outputs and the survey's known answer.  Tolerances: 1e-10 relative on lZ / mu /
s2, 1e-8 on gradients (BASELINE.json north_star)."""


import numpy as np
import numpy.testing as nt
import pytest

from oracle.cases import GP_CASES, SURVEY_KAT, GP_SN, GP_MEAN, gp_inputs
from oracle.pygp_oracle import make_kernel, OFITC, synthetic_problem
from gpu_util import product_kernel, assert_grad_close, assert_pred_close, LZ_RTOL

pytestmark = pytest.mark.gpu

FITC_CASES = sorted(n for n, c in GP_CASES.items() if c[3])


def build(name):
    import pygp_b200 as pygp
    spec, N, d, _ = GP_CASES[name]
    X, y, Xs, U = gp_inputs(N, d, True)
    gp = pygp.inference.FITC(pygp.likelihoods.Gaussian(GP_SN), product_kernel(spec), GP_MEAN, U)
    gp.add_data(X, y)
    return gp, Xs


@pytest.mark.parametrize('name', FITC_CASES)
def test_vs_golden(name, golden):
    g = golden['gp']
    gp, Xs = build(name)
    lZ, dlZ = gp.loglikelihood(True)
    assert gp.loglikelihood() == lZ
    mu, s2 = gp.posterior(Xs)
    nt.assert_allclose(lZ, g[name + '/lZ'], rtol=LZ_RTOL)
    assert_grad_close(dlZ, g[name + '/dlZ'])
    assert_pred_close(mu, s2, g[name + '/mu'], g[name + '/s2'])
    gp.set_hyper(g[name + '/hyper2'])
    lZ, dlZ = gp.loglikelihood(True)
    mu, s2 = gp.posterior(Xs)
    nt.assert_allclose(lZ, g[name + '/lZ_h2'], rtol=LZ_RTOL)
    assert_grad_close(dlZ, g[name + '/dlZ_h2'])
    assert_pred_close(mu, s2, g[name + '/mu_h2'], g[name + '/s2_h2'])


def test_vs_survey_kat():
    lZ0, dlZ0, mu0, s20 = SURVEY_KAT['fitc_se_2d']
    gp, Xs = build('fitc_se_2d')
    lZ, dlZ = gp.loglikelihood(True)
    mu, s2 = gp.posterior(Xs)
    nt.assert_allclose(lZ, lZ0, rtol=LZ_RTOL)
    assert_grad_close(dlZ, dlZ0)
    assert_pred_close(mu, s2, mu0, s20)


@pytest.mark.parametrize('spec,N,d,p,m', [
    (('se', 1.0, [0.5*np.sqrt(8)]*8), 3000, 8, 128, 257),       # C5-FITC shape, scaled down
    (('matern', 1.0, [0.8, 0.9, 0.7], 5), 5001, 3, 65, 100),      # ragged p and n, Matern guard path
    (('sum', ('se', 1.0, [0.15, 0.2]), ('rq', 0.5, [0.2, 0.25], 0.8)), 1200, 2, 96, 64),   # composite
    (('se', 1.0, 0.05, 1), 70000, 1, 33, 40),                     # n >> p: split-K accumulation, 2 Kxu chunks
])
def test_vs_oracle_medium(spec, N, d, p, m):
    """Against the oracle at sizes it finishes in seconds.  FITC's mean goes
    through chol(Kuu + sn2/1e6 I), whose condition number (1e6..1e9 here) makes
    ANY reordering of the float64 arithmetic move mu by more than 1e-10: the
    floor of the tolerance is therefore the oracle's own reordering sensitivity,
    measured with the numpy model of the device formulation (oracle/fitc_model.py)."""
    import pygp_b200 as pygp
    from oracle import fitc_model as fm
    X, y, Xs = synthetic_problem(N, d, m)
    U = np.random.RandomState(3).rand(p, d) if d > 1 else np.linspace(0, 1, p)[:, None]
    gp = pygp.inference.FITC(pygp.likelihoods.Gaussian(0.1), product_kernel(spec), 0.1, U)
    gp.add_data(X, y)
    ok = make_kernel(spec)
    ogp = OFITC(0.1, ok, 0.1, U)
    ogp.add_data(X, y)
    lZ, dlZ = gp.loglikelihood(True)
    olZ, odlZ = ogp.loglikelihood(True)
    mu, s2 = gp.posterior(Xs)
    omu, os2 = ogp.posterior(Xs)
    st = fm.fitc_update(ok, 0.01, 0.1, U, X, y)
    mlZ, mdlZ = fm.fitc_loglike(ok, st, U, X, True)
    mmu, ms2 = fm.fitc_predict(ok, st, U, Xs)
    sens = lambda a, b: 20*float(np.max(np.abs(np.asarray(a) - np.asarray(b))))
    nt.assert_allclose(lZ, olZ, rtol=LZ_RTOL, atol=sens(mlZ, olZ))
    gs = np.abs(odlZ).max()
    nt.assert_allclose(dlZ, odlZ, rtol=1e-8, atol=max(1e-8*gs, sens(mdlZ, odlZ)))
    sf2 = float(np.max(ok.dget(Xs[:1])))
    nt.assert_allclose(mu, omu, rtol=1e-10, atol=max(1e-10*np.abs(y).max(), sens(mmu, omu)))
    nt.assert_allclose(s2, os2, rtol=1e-10, atol=max(1e-10*sf2, sens(ms2, os2)))


def test_reference_interface():
    """tests/test_inference.py:174-207 -- FITC ctor, from_gp, reset, add_data in
    two halves == all at once, copy(hyper) semantics."""
    import pygp_b200 as pygp
    spec, N, d, _ = GP_CASES['fitc_ard4_500']
    X, y, Xs, U = gp_inputs(N, d, True)
    mk = lambda: pygp.inference.FITC(pygp.likelihoods.Gaussian(GP_SN), product_kernel(spec), GP_MEAN, U)
    gp = mk()
    # prior
    mu, s2 = gp.posterior(Xs)
    nt.assert_allclose(mu, GP_MEAN)
    nt.assert_allclose(s2, gp._kernel.dget(Xs))
    gp.add_data(X, y)
    g2 = mk()
    g2.add_data(X[:200], y[:200])
    g2.add_data(X[200:], y[200:])
    nt.assert_allclose(g2.loglikelihood(), gp.loglikelihood(), rtol=1e-12)
    nt.assert_allclose(g2.posterior(Xs)[0], gp.posterior(Xs)[0], rtol=1e-10)
    # copy(hyper): deep copy + set_hyper re-factorises on its own device state
    h2 = gp.get_hyper() + 0.05
    g3 = gp.copy(h2)
    g4 = mk()
    g4.set_hyper(h2)
    g4.add_data(X, y)
    nt.assert_allclose(g3.loglikelihood(), g4.loglikelihood(), rtol=1e-12)
    nt.assert_allclose(gp.get_hyper(), mk().get_hyper())          # original untouched
    # copy() WITHOUT a hyper argument is fully usable (utils/models.py:47-55): the copy has the
    # data but rebuilds its own device state on first use
    g7 = gp.copy()
    nt.assert_allclose(g7.posterior(Xs)[0], gp.posterior(Xs)[0], rtol=1e-12)
    nt.assert_allclose(g7.posterior(Xs)[1], gp.posterior(Xs)[1], rtol=1e-12)
    nt.assert_allclose(gp.copy().loglikelihood(True)[1], gp.loglikelihood(True)[1], rtol=1e-12)
    assert gp.copy().sample(Xs[:3], rng=0).shape == (3,)
    # from_gp (fitc.py:53-64)
    g5 = pygp.inference.FITC.from_gp(gp)
    nt.assert_allclose(g5.loglikelihood(), gp.loglikelihood(), rtol=1e-12)
    ex = pygp.inference.ExactGP(pygp.likelihoods.Gaussian(GP_SN), product_kernel(spec), GP_MEAN)
    with pytest.raises(ValueError):
        pygp.inference.FITC.from_gp(ex)
    g6 = pygp.inference.FITC.from_gp(ex, U)
    assert g6.ndata == 0 and g6.pseudoinputs.shape == U.shape
    gp.reset()
    assert gp.ndata == 0
    nt.assert_allclose(gp.posterior(Xs)[0], GP_MEAN)


def test_fitc_equals_exact_when_u_is_x():
    """With U = X FITC is the exact GP up to the su2 jitter: a size-independent
    property the domain offers (lZ within 1e-5 for su2 = sn2 / 1e6)."""
    import pygp_b200 as pygp
    X, y, Xs = synthetic_problem(200, 2, 20)
    k = ('se', 1.0, [0.7, 0.8])
    f = pygp.inference.FITC(pygp.likelihoods.Gaussian(0.3), product_kernel(k), 0.0, X)
    e = pygp.inference.ExactGP(pygp.likelihoods.Gaussian(0.3), product_kernel(k), 0.0)
    f.add_data(X, y)
    e.add_data(X, y)
    nt.assert_allclose(f.loglikelihood(), e.loglikelihood(), rtol=1e-5)
    nt.assert_allclose(f.posterior(Xs)[0], e.posterior(Xs)[0], rtol=1e-4, atol=1e-5)


def test_not_positive_definite_raises():
    import pygp_b200 as pygp
    X, y, _ = synthetic_problem(50, 1, 0)
    U = np.r_[X[:4], X[:4]]                      # duplicated pseudo-inputs, tiny jitter
    gp = pygp.inference.FITC(pygp.likelihoods.Gaussian(1e-9), pygp.kernels.SE(1.0, 5.0, ndim=1), 0.0, U)
    with pytest.raises(np.linalg.LinAlgError):
        gp.add_data(X, y)


@pytest.mark.parametrize('name', FITC_CASES)
def test_posterior_input_gradients(name, golden):
    """FITC.posterior(X, grad=True) (fitc.py:144-165) against the reference's outputs."""
    g = golden['gp']
    gp, Xs = build(name)
    mu, s2, dmu, ds2 = gp.posterior(Xs, grad=True)
    assert_pred_close(mu, s2, g[name + '/mu'], g[name + '/s2'])
    nt.assert_allclose(dmu, g[name + '/dmu'], rtol=1e-8, atol=1e-9*max(1.0, np.abs(g[name + '/dmu']).max()))
    nt.assert_allclose(ds2, g[name + '/ds2'], rtol=1e-7, atol=1e-9*max(1.0, np.abs(g[name + '/ds2']).max()))


@pytest.mark.parametrize('name', FITC_CASES)
def test_full_posterior_and_sample(name, golden):
    """_full_posterior and GP.sample (_base.py:143-177) against the reference's own
    joint covariance and its draws for the same seed (the normal variates come from
    the host rng in the reference's order; Cholesky + transform run on the device)."""
    g = golden['gp']
    gp, _ = build(name)
    Xj = g[name + '/Xj']
    mu, Sigma = gp._full_posterior(Xj)
    nt.assert_allclose(mu, g[name + '/full_mu'], rtol=1e-10, atol=1e-10)
    nt.assert_allclose(Sigma, g[name + '/full_Sigma'], rtol=1e-8, atol=1e-11)
    nt.assert_allclose(Sigma, Sigma.T, rtol=0, atol=1e-14)
    # marginals of the joint == the marginal posterior
    m1, s1 = gp.posterior(Xj)
    nt.assert_allclose(np.diag(Sigma), s1, rtol=1e-9, atol=1e-12)
    nt.assert_allclose(mu, m1, rtol=1e-12)
    f = gp.sample(Xj, 3, latent=False, rng=5)
    # the draws go through chol(Sigma + 1e-10 I): the factor of a covariance with 1e-4..1e-2
    # eigenvalues amplifies 1e-11 differences in Sigma; 1e-6 is the reference-vs-oracle level too
    nt.assert_allclose(f, g[name + '/sample'], rtol=1e-6, atol=1e-7)
    assert gp.sample(Xj, rng=0).shape == (len(Xj),)
    gp.reset()                                         # prior draws
    assert gp.sample(Xj, 2, rng=1).shape == (2, len(Xj))
