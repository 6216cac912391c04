"""Parity of the DTC path (pgp_dtc_create + pgp_fitc_*; pygp/inference/dtc.py:54-199)
with the committed reference outputs and the oracle: lZ, gradient, posterior with
input-gradients, joint posterior.  Tolerances as for FITC (1e-10 / 1e-8 with the
oracle's reordering sensitivity as the floor for ill-conditioned Kuu)."""

import numpy as np
import numpy.testing as nt
import pytest

from oracle.cases import DTC_CASES, GP_SN, GP_MEAN, gp_inputs
from oracle.pygp_oracle import make_kernel, ODTC, synthetic_problem
from gpu_util import product_kernel, assert_grad_close, assert_pred_close, LZ_RTOL

pytestmark = pytest.mark.gpu


def build(name):
    import pygp_b200 as pygp
    spec, N, d = DTC_CASES[name]
    X, y, Xs, U = gp_inputs(N, d, True)
    gp = pygp.inference.DTC(pygp.likelihoods.Gaussian(GP_SN), product_kernel(spec), GP_MEAN, U)
    gp.add_data(X, y)
    return gp, Xs


@pytest.mark.parametrize('name', sorted(DTC_CASES))
def test_vs_golden(name, golden):
    g = golden['gp']
    gp, Xs = build(name)
    for tag in ('', '_h2'):
        if tag:
            gp.set_hyper(g[name + '/hyper2'])
        lZ, dlZ = gp.loglikelihood(True)
        assert gp.loglikelihood() == lZ
        mu, s2, dmu, ds2 = gp.posterior(Xs, grad=True)
        nt.assert_allclose(lZ, g[name + '/lZ' + tag], rtol=LZ_RTOL)
        assert_grad_close(dlZ, g[name + '/dlZ' + tag])
        assert_pred_close(mu, s2, g[name + '/mu' + tag], g[name + '/s2' + tag])
        nt.assert_allclose(dmu, g[name + '/dmu' + tag], rtol=1e-8, atol=1e-9*max(1.0, np.abs(g[name + '/dmu' + tag]).max()))
        nt.assert_allclose(ds2, g[name + '/ds2' + tag], rtol=1e-7, atol=1e-9*max(1.0, np.abs(g[name + '/ds2' + tag]).max()))
        fmu, fS = gp._full_posterior(Xs)
        nt.assert_allclose(fmu, mu, rtol=1e-12)
        nt.assert_allclose(fS, g[name + '/full_Sigma' + tag], rtol=1e-7, atol=1e-10)


@pytest.mark.parametrize('spec,N,d,p,m', [
    (('se', 1.0, [0.5*np.sqrt(8)]*8), 3000, 8, 128, 100),
    (('sum', ('se', 1.0, [0.15, 0.2]), ('rq', 0.5, [0.2, 0.25], 0.8)), 1200, 2, 96, 64),
    (('se', 1.0, 0.05, 1), 70000, 1, 33, 40),
])
def test_vs_oracle_medium(spec, N, d, p, m):
    import pygp_b200 as pygp
    from oracle import fitc_model as fm
    X, y, Xs = synthetic_problem(N, d, m)
    U = np.random.RandomState(3).rand(p, d) if d > 1 else np.linspace(0, 1, p)[:, None]
    gp = pygp.inference.DTC(pygp.likelihoods.Gaussian(0.1), product_kernel(spec), 0.1, U)
    gp.add_data(X, y)
    ok = make_kernel(spec)
    ogp = ODTC(0.1, ok, 0.1, U)
    ogp.add_data(X, y)
    lZ, dlZ = gp.loglikelihood(True)
    olZ, odlZ = ogp.loglikelihood(True)
    mu, s2 = gp.posterior(Xs)
    omu, os2 = ogp.posterior(Xs)
    st = fm.dtc_update(ok, 0.01, 0.1, U, X, y)
    mlZ, mdlZ = fm.dtc_loglike(ok, st, U, X, True)
    mmu, ms2 = fm.dtc_predict(ok, st, U, Xs)
    sens = lambda a, b: 20*float(np.max(np.abs(np.asarray(a) - np.asarray(b))))
    nt.assert_allclose(lZ, olZ, rtol=LZ_RTOL, atol=sens(mlZ, olZ))
    gs = np.abs(odlZ).max()
    nt.assert_allclose(dlZ, odlZ, rtol=1e-8, atol=max(1e-8*gs, sens(mdlZ, odlZ)))
    sf2 = float(np.max(ok.dget(Xs[:1])))
    nt.assert_allclose(mu, omu, rtol=1e-10, atol=max(1e-10*np.abs(y).max(), sens(mmu, omu)))
    nt.assert_allclose(s2, os2, rtol=1e-10, atol=max(1e-10*sf2, sens(ms2, os2)))


def test_reference_interface():
    import pygp_b200 as pygp
    spec, N, d = DTC_CASES['dtc_se_2d']
    X, y, Xs, U = gp_inputs(N, d, True)
    gp = pygp.inference.DTC(pygp.likelihoods.Gaussian(GP_SN), product_kernel(spec), GP_MEAN, U)
    assert isinstance(gp, pygp.inference.DTC) and gp.pseudoinputs.shape == U.shape
    mu, s2 = gp.posterior(Xs)                       # prior
    nt.assert_allclose(mu, GP_MEAN)
    gp.add_data(X, y)
    g2 = pygp.inference.DTC.from_gp(gp)
    assert isinstance(g2, pygp.inference.DTC)
    nt.assert_allclose(g2.loglikelihood(), gp.loglikelihood(), rtol=1e-12)
    f = pygp.inference.FITC.from_gp(gp)             # a different approximation: a different likelihood
    assert abs(f.loglikelihood() - gp.loglikelihood()) > 1e-6
    assert gp.sample(Xs, 2, rng=0).shape == (2, len(Xs))


def test_copy_without_hyper_is_usable():
    """`DTC.copy()` (utils/models.py:47-55) gives a model whose posterior / likelihood work at once."""
    gp, Xs = build(sorted(DTC_CASES)[0])
    g2 = gp.copy()
    nt.assert_allclose(g2.posterior(Xs)[0], gp.posterior(Xs)[0], rtol=1e-12)
    nt.assert_allclose(g2.loglikelihood(), gp.loglikelihood(), rtol=1e-12)
