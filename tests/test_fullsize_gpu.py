"""BASELINE.json's configurations pinned AT FULL SIZE against the oracle (the CPU restatement of the
reference, itself pinned to the live reference by tests/golden + oracle/make_golden.py):

  C2   ExactGP SE-ARD d=8, N=8192 (pygp/inference/exact.py:50-55,81-97,118-143): lZ 1e-10, dlZ 1e-8,
       posterior 1e-10 -- the oracle needs ~1-2 minutes of CPU for its one evaluation
  C4   64 of the 4096 hyper vectors x N=2048 through pgp_batched_loglike (values AND gradients) against
       the oracle one by one (learning/sampling.py:146, meta/mcmc.py:75-93)
  FITC N=65536, M=512 against OFITC (pygp/inference/fitc.py:66-232)
  C3 (N=32768, oracle needs > 72 GiB) is covered by size-independent properties in
  tests/test_exact_gpu.py::test_full_size_properties_n32768; C5 (N=65536, distributed == single GPU)
  by tests/test_multigpu.py.

Synthetic inputs and base hypers as SURVEY.md 8d / bench.py."""

import ctypes as C

import numpy as np
import numpy.testing as nt
import pytest

from oracle.pygp_oracle import make_kernel, OExactGP, OFITC, synthetic_problem
from gpu_util import product_kernel, assert_grad_close, assert_pred_close, LZ_RTOL

pytestmark = pytest.mark.gpu


@pytest.mark.timeout(1200)
def test_c2_se_ard_n8192_vs_oracle():
    import pygp_b200 as pygp
    n, d, m = 8192, 8, 512
    X, y, Xs = synthetic_problem(n, d, m)
    spec = ('se', 1.0, [0.5*np.sqrt(d)]*d)
    gp = pygp.inference.ExactGP(pygp.likelihoods.Gaussian(0.1), product_kernel(spec), 0.0)
    gp.add_data(X, y)
    h = gp.get_hyper() + 0.03*np.random.RandomState(5).randn(gp.nhyper)    # an optimiser iterate, not the start point
    gp.set_hyper(h)
    lZ, dlZ = gp.loglikelihood(True)
    mu, s2 = gp.posterior(Xs)
    ogp = OExactGP(0.1, make_kernel(spec), 0.0)
    ogp.add_data(X, y)
    ogp.set_hyper(h)
    olZ, odlZ = ogp.loglikelihood(True)
    omu, os2 = ogp.posterior(Xs)
    nt.assert_allclose(lZ, olZ, rtol=LZ_RTOL)
    assert_grad_close(dlZ, odlZ)
    assert_pred_close(mu, s2, omu, os2, yscale=np.abs(y).max(), sf2=np.exp(2*h[1]))


@pytest.mark.timeout(900)
def test_c4_batched_64_of_4096_n2048_vs_oracle():
    import pygp_b200 as pygp
    from pygp_b200 import _lib
    n, d, B_all, B = 2048, 8, 4096, 64
    X, y, _ = synthetic_problem(n, d, 0, seed=2)
    spec = ('se', 1.0, [0.5*np.sqrt(d)]*d)
    k = product_kernel(spec)
    gp = pygp.inference.ExactGP(pygp.likelihoods.Gaussian(0.1), k, 0.0)
    gp.add_data(X, y)
    H_all = gp.get_hyper() + np.random.RandomState(2).uniform(-0.5, 0.5, size=(B_all, gp.nhyper))
    H = np.ascontiguousarray(H_all[::B_all//B])                           # every 64th of the 4096 samples
    lZ, dlZ, info = np.empty(B), np.empty((B, gp.nhyper)), np.zeros(B, dtype=np.int32)
    ctx = _lib.context()
    _lib.check(ctx, _lib.lib().pgp_batched_loglike(ctx.handle, k._spec(), _lib.ptr(X), _lib.ptr(y), n, _lib.ptr(H), B,
                                                   _lib.ptr(lZ), _lib.ptr(dlZ), info.ctypes.data_as(C.POINTER(C.c_int32))))
    assert not info.any()
    og = OExactGP(0.1, make_kernel(spec), 0.0)
    og.add_data(X, y)
    for b in range(B):
        og.set_hyper(H[b])
        if b % 8 == 0:                                                    # gradient on every 8th (the oracle's is ~2 s each)
            olZ, odlZ = og.loglikelihood(True)
            assert_grad_close(dlZ[b], odlZ)
        else:
            olZ = og.loglikelihood()
        nt.assert_allclose(lZ[b], olZ, rtol=LZ_RTOL)


@pytest.mark.timeout(1500)
def test_fitc_n65536_m512_vs_oracle():
    import pygp_b200 as pygp
    n, d, p, m = 65536, 8, 512, 256
    X, y, Xs = synthetic_problem(n, d, m)
    U = np.random.RandomState(3).rand(p, d)
    spec = ('se', 1.0, [0.5*np.sqrt(d)]*d)
    gp = pygp.inference.FITC(pygp.likelihoods.Gaussian(0.1), product_kernel(spec), 0.0, U)
    gp.add_data(X, y)
    lZ, dlZ = gp.loglikelihood(True)
    mu, s2 = gp.posterior(Xs)
    og = OFITC(0.1, make_kernel(spec), 0.0, U)
    og.add_data(X, y)
    olZ, odlZ = og.loglikelihood(True)
    omu, os2 = og.posterior(Xs)
    # FITC's Kuu + 1e-6 sn2 I is ill-conditioned by construction (fitc.py:68): ANY reordering of the float64
    # arithmetic moves the results by more than 1e-10.  As in tests/test_fitc_gpu.py the floor of the tolerance is
    # the oracle's own reordering sensitivity, measured with the numpy model of the device formulation.
    from oracle import fitc_model as fm
    ok = make_kernel(spec)
    st = fm.fitc_update(ok, 0.01, 0.0, U, X, y)
    mlZ, mdlZ = fm.fitc_loglike(ok, st, U, X, True)
    mmu, ms2 = fm.fitc_predict(ok, st, U, Xs)
    sens = lambda a, b: 20*float(np.max(np.abs(np.asarray(a) - np.asarray(b))))
    nt.assert_allclose(lZ, olZ, rtol=LZ_RTOL, atol=sens(mlZ, olZ))
    gs = np.abs(odlZ).max()
    nt.assert_allclose(dlZ, odlZ, rtol=1e-8, atol=max(1e-8*gs, sens(mdlZ, odlZ)))
    nt.assert_allclose(mu, omu, rtol=1e-10, atol=max(1e-10*np.abs(y).max(), sens(mmu, omu)))
    nt.assert_allclose(s2, os2, rtol=1e-10, atol=max(1e-10, sens(ms2, os2)))
