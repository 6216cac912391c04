"""Parity of the ExactGP hot path (pgp_exact_update / _loglike / _predict)
with the oracle, the committed reference outputs and the survey's known
answers.  Tolerances: 1e-10 relative on lZ / mu / s2, 1e-8 on gradients
(BASELINE.json north_star)."""

import numpy as np
import numpy.testing as nt
import pytest
import scipy.optimize as spop

from oracle.cases import GP_CASES, SURVEY_KAT, GP_SN, GP_MEAN, gp_inputs
from oracle.pygp_oracle import make_kernel, OExactGP, synthetic_problem
from gpu_util import product_kernel, assert_grad_close, assert_pred_close, LZ_RTOL

pytestmark = pytest.mark.gpu

EXACT = sorted(n for n, c in GP_CASES.items() if not c[3])


def build(name):
    import pygp_b200 as pygp
    spec, N, d, _ = GP_CASES[name]
    X, y, Xs, _ = gp_inputs(N, d)
    gp = pygp.inference.ExactGP(pygp.likelihoods.Gaussian(GP_SN), product_kernel(spec), GP_MEAN)
    gp.add_data(X, y)
    return gp, Xs


@pytest.mark.parametrize('name', EXACT)
def test_vs_golden(name, golden):
    g = golden['gp']
    gp, Xs = build(name)
    lZ, dlZ = gp.loglikelihood(True)
    assert gp.loglikelihood() == lZ
    mu, s2 = gp.posterior(Xs)
    nt.assert_allclose(lZ, g[name + '/lZ'], rtol=LZ_RTOL)
    assert_grad_close(dlZ, g[name + '/dlZ'])
    assert_pred_close(mu, s2, g[name + '/mu'], g[name + '/s2'])
    R, a = gp._R, gp._a
    if R.shape[0] <= 64:
        nt.assert_allclose(R, g[name + '/R'], rtol=1e-10, atol=1e-12)
    nt.assert_allclose(a, g[name + '/a'], rtol=1e-9, atol=1e-10)
    gp.set_hyper(g[name + '/hyper2'])          # the optimiser's access path
    lZ, dlZ = gp.loglikelihood(True)
    mu, s2 = gp.posterior(Xs)
    nt.assert_allclose(lZ, g[name + '/lZ_h2'], rtol=LZ_RTOL)
    assert_grad_close(dlZ, g[name + '/dlZ_h2'])
    assert_pred_close(mu, s2, g[name + '/mu_h2'], g[name + '/s2_h2'])


@pytest.mark.parametrize('name', sorted(set(SURVEY_KAT) & set(EXACT)))
def test_vs_survey_kat(name):
    lZ0, dlZ0, mu0, s20 = SURVEY_KAT[name]
    gp, Xs = build(name)
    lZ, dlZ = gp.loglikelihood(True)
    mu, s2 = gp.posterior(Xs)
    nt.assert_allclose(lZ, lZ0, rtol=LZ_RTOL)
    assert_grad_close(dlZ, dlZ0)
    assert_pred_close(mu, s2, mu0, s20)


@pytest.mark.parametrize('spec,N,d,m', [
    (('se', 1.0, 0.1, 1), 1000, 1, 500),                      # config C1
    (('se', 1.0, [0.5*np.sqrt(8)]*8), 1537, 8, 300),          # C2-like, ragged N
    (('matern', 1.0, [2.0]*16, 5), 2048, 16, 257),            # C3-like
    (('matern', 1.0, [0.3, 0.2], 1), 700, 2, 64),
    (('sum', ('se', 1.0, 0.5), ('periodic', 0.5, 1.0, 0.25)), 901, 1, 100),   # C5-like
    (('rq', 0.8, [0.7, 0.9, 1.1], 1.5), 450, 3, 33),
])
def test_vs_oracle_medium(spec, N, d, m):
    import pygp_b200 as pygp
    X, y, Xs = synthetic_problem(N, d, m)
    if spec[0] == 'sum':
        X, Xs = np.sort(X, axis=0)*8, Xs*8
    gp = pygp.inference.ExactGP(pygp.likelihoods.Gaussian(0.1), product_kernel(spec), 0.0)
    gp.add_data(X, y)
    ogp = OExactGP(0.1, make_kernel(spec), 0.0)
    ogp.add_data(X, y)
    lZ, dlZ = gp.loglikelihood(True)
    olZ, odlZ = ogp.loglikelihood(True)
    nt.assert_allclose(lZ, olZ, rtol=LZ_RTOL)
    assert_grad_close(dlZ, odlZ)
    mu, s2 = gp.posterior(Xs)
    omu, os2 = ogp.posterior(Xs)
    assert_pred_close(mu, s2, omu, os2, yscale=np.abs(y).max(), sf2=ogp._kernel.dget(Xs[:1])[0])
    nt.assert_allclose(gp._a, ogp._a, rtol=1e-8, atol=1e-9*np.abs(ogp._a).max())


def test_properties_large():
    """Size-independent properties at a size the oracle would take minutes for:
    L L^T reproduces K + sn2 I, a solves L a = r, and the analytic gradient
    matches a central finite difference of lZ."""
    import pygp_b200 as pygp
    N, d = 4096 + 77, 4
    X, y, _ = synthetic_problem(N, d)
    k = pygp.kernels.SE(1.0, [0.6]*d)
    gp = pygp.inference.ExactGP(pygp.likelihoods.Gaussian(0.1), k, 0.1)
    gp.add_data(X, y)
    R, a = gp._R, gp._a
    assert np.all(np.triu(R) == R)
    K = k.get(X) + 0.1**2*np.eye(N)
    nt.assert_allclose(R.T @ R, K, rtol=1e-12, atol=1e-12)
    nt.assert_allclose(R.T @ a, y - 0.1, rtol=1e-10, atol=1e-11)
    lZ, dlZ = gp.loglikelihood(True)
    nt.assert_allclose(lZ, -0.5*a@a - 0.5*N*np.log(2*np.pi) - np.log(np.diag(R)).sum(), rtol=1e-12)
    h = gp.get_hyper()
    for i in (0, 1, 3, len(h)-1):
        e = np.zeros_like(h)
        e[i] = 1e-5
        fd = (gp.copy(h+e).loglikelihood() - gp.copy(h-e).loglikelihood()) / 2e-5
        nt.assert_allclose(dlZ[i], fd, rtol=2e-6, atol=1e-6*np.abs(dlZ).max())


def test_not_positive_definite_raises():
    # the reference raises LinAlgError from scipy.linalg.cholesky (exact.py:54); no jitter
    import pygp_b200 as pygp
    X = np.r_[np.linspace(0, 1, 40), np.linspace(0, 1, 40)][:, None]     # duplicated rows
    gp = pygp.inference.ExactGP(pygp.likelihoods.Gaussian(1e-12), pygp.kernels.SE(1.0, 5.0, ndim=1), 0.0)
    with pytest.raises(np.linalg.LinAlgError):
        gp.add_data(X, np.sin(X[:, 0]))


def test_add_data_reset_copy_prior():
    # reference tests/test_inference.py:39-79,132-145 (full-update branch of add_data)
    import pygp_b200 as pygp
    rng = np.random.RandomState(1)
    X, y = rng.rand(10, 2), rng.rand(10)
    X2, y2 = rng.rand(10, 2), rng.rand(10)
    gp = pygp.inference.ExactGP(pygp.likelihoods.Gaussian(1), pygp.kernels.SE(1, 1, ndim=2), 0.0)
    mu, s2 = gp.posterior(X2)                   # prior
    nt.assert_array_equal(mu, np.zeros(10))
    nt.assert_allclose(s2, np.ones(10))
    gp.add_data(X, y)
    gp1 = gp.copy()
    gp1.add_data(X2, y2)
    gp2 = pygp.inference.ExactGP(pygp.likelihoods.Gaussian(1), pygp.kernels.SE(1, 1, ndim=2), 0.0)
    gp2.add_data(np.r_[X, X2], np.r_[y, y2])
    nt.assert_allclose(gp1.posterior(X2), gp2.posterior(X2), rtol=1e-12)
    nt.assert_allclose(gp1.loglikelihood(), gp2.loglikelihood(), rtol=1e-12)
    assert gp.ndata == 10 and gp1.ndata == 20
    p0 = gp.posterior(X2)
    nt.assert_allclose(gp.copy().posterior(X2), p0, rtol=0, atol=0)
    gp3 = gp.copy()
    gp3.reset()
    assert gp3.ndata == 0 and gp3.data == (None, None)
    gp3.add_data(X, y)
    nt.assert_allclose(gp3.posterior(X2), p0)
    gp4 = pygp.inference.ExactGP.from_gp(gp)
    nt.assert_allclose(gp4.posterior(X2), p0)
    h = gp.get_hyper()
    gp.set_hyper(h)
    nt.assert_allclose(gp.get_hyper(), h)


def test_loglikelihood_fd_basic():
    # reference tests/test_inference.py:105-112 with BasicGP(1,1,1,0,ndim=2)
    import pygp_b200 as pygp
    rng = np.random.RandomState(1)
    gp = pygp.BasicGP(1, 1, 1, 0, ndim=2)
    X = rng.rand(10, 2)
    gp.add_data(X, gp._likelihood.sample(rng.rand(10), rng))
    x = gp.get_hyper()
    _, g1 = gp.loglikelihood(grad=True)
    g2 = spop.approx_fprime(x, lambda x_: gp.copy(x_).loglikelihood(), 1e-8)
    nt.assert_allclose(g1, g2, rtol=1e-5, atol=1e-5)
    for kern in ('se', 'matern1', 'matern3', 'matern5'):
        _ = pygp.BasicGP.from_gp(pygp.BasicGP(1, 1, 1, 0, 2, kern))


@pytest.mark.parametrize('N,nb', [(700, 128), (513, 64), (256, 256)])
def test_distributed_update_single_rank(N, nb):
    """pgp_dist_exact_update / pgp_dist_exact_loglike (csrc/dist.cu: block columns built and
    factored in place, staircase solves, block-column trace) on one GPU without a process
    group == pgp_exact_update / pgp_exact_loglike.  The N > 1 schedule is covered on CPU by its
    numpy model (tests/test_distchol.py) and on GPUs by tests/test_multigpu.py."""
    import pygp_b200 as pygp
    from pygp_b200.distchol import distributed_update, distributed_loglikelihood
    X, y, Xs = synthetic_problem(N, 4, 50)
    spec = ('matern', 1.0, [0.7, 0.8, 0.9, 1.0], 5)
    mk = lambda: pygp.inference.ExactGP(pygp.likelihoods.Gaussian(0.1), product_kernel(spec), 0.1)
    ref = mk()
    ref.add_data(X, y)
    gp = mk()
    gp.add_data(X, y)
    h2 = gp.get_hyper() + 0.05
    ref.set_hyper(h2)
    # set the new hypers on the host objects only, then factor through the distributed path
    a, b = gp._likelihood.nhyper, gp._kernel.nhyper
    gp._likelihood.set_hyper(h2[:a]); gp._kernel.set_hyper(h2[a:a + b]); gp._mean = float(h2[-1])
    distributed_update(gp, nb=nb)
    lZ0, dlZ0 = ref.loglikelihood(True)
    lZ, dlZ = distributed_loglikelihood(gp, True, nb=nb)      # block-column gradient
    nt.assert_allclose(lZ, lZ0, rtol=1e-11)
    assert_grad_close(dlZ, dlZ0, rtol=1e-9)
    lZ, dlZ = gp.loglikelihood(True)                          # replicated gradient on the adopted factor
    nt.assert_allclose(lZ, lZ0, rtol=1e-11)
    assert_grad_close(dlZ, dlZ0, rtol=1e-9)
    mu, s2 = gp.posterior(Xs)
    mu0, s20 = ref.posterior(Xs)
    nt.assert_allclose(mu, mu0, rtol=1e-10, atol=1e-11)
    nt.assert_allclose(s2, s20, rtol=1e-9, atol=1e-12)


def test_distributed_option_through_the_public_api():
    """`ExactGP.distributed` (opt-in, here forced on a single rank): add_data / set_hyper / loglikelihood(True) /
    posterior go through the block-column path and agree with the default path."""
    import pygp_b200 as pygp
    X, y, Xs = synthetic_problem(450, 3, 20)
    spec = ('se', 1.0, [0.6, 0.7, 0.8])
    mk = lambda: pygp.inference.ExactGP(pygp.likelihoods.Gaussian(0.1), product_kernel(spec), 0.0)
    ref, gp = mk(), mk()
    gp.distributed = {'nb': 128, 'min_n': 100, 'force': True}
    ref.add_data(X, y)
    gp.add_data(X, y)
    h = ref.get_hyper() + 0.04
    ref.set_hyper(h)
    gp.set_hyper(h)
    lZ0, dlZ0 = ref.loglikelihood(True)
    lZ, dlZ = gp.loglikelihood(True)
    nt.assert_allclose(lZ, lZ0, rtol=1e-11)
    assert_grad_close(dlZ, dlZ0, rtol=1e-9)
    nt.assert_allclose(gp.loglikelihood(), lZ0, rtol=1e-11)
    mu, s2 = gp.posterior(Xs)
    mu0, s20 = ref.posterior(Xs)
    nt.assert_allclose(mu, mu0, rtol=1e-10, atol=1e-11)
    nt.assert_allclose(s2, s20, rtol=1e-9, atol=1e-12)


def test_distributed_update_not_positive_definite():
    """A matrix that is not positive definite comes back from the distributed path as the same LinAlgError
    the one-GPU path raises (scipy.linalg.cholesky at exact.py:54), with the failing minor reported."""
    import pygp_b200 as pygp
    from pygp_b200.distchol import distributed_update
    X, y, _ = synthetic_problem(300, 2, 0)
    X[200] = X[10]                                   # duplicated input + (almost) no noise: singular K
    gp = pygp.inference.ExactGP(pygp.likelihoods.Gaussian(0.1), pygp.kernels.SE(1.0, [0.5, 0.5]), 0.0)
    gp.add_data(X, y)
    gp._likelihood.set_hyper(np.array([np.log(1e-12)]))
    with pytest.raises(np.linalg.LinAlgError):
        distributed_update(gp, nb=64)
    with pytest.raises(np.linalg.LinAlgError):
        gp.set_hyper(gp.get_hyper())
    gp._likelihood.set_hyper(np.array([np.log(0.1)]))   # and the model recovers on both paths
    distributed_update(gp, nb=64)
    lZ = gp.loglikelihood()
    gp.set_hyper(gp.get_hyper())
    nt.assert_allclose(gp.loglikelihood(), lZ, rtol=1e-12)


def test_full_size_properties_n32768():
    """BASELINE configs[2] at full size (Matern-5/2 ARD d=16, N=32768), where the
    oracle cannot run (> 72 GiB): size-independent properties of the same
    device state instead.
      (1) L L^T reproduces K + sn2 I on sampled rows (factor correctness);
      (2) a = L^-1 r: |a|^2 and sum log diag L reproduce lZ (exact.py:119-121);
      (3) posterior mean at training inputs equals y - sn2 alpha, with alpha taken
          from dlZ's mean entry sum(alpha) (ties predict, solve and gradient paths);
      (4) directional finite difference of lZ matches dlZ (tests/test_inference.py:105-112)."""
    import pygp_b200 as pygp
    n, d = 32768, 16
    X, y, _ = synthetic_problem(n, d, 0)
    ell = [0.5*np.sqrt(d)]*d
    kern = pygp.kernels.Matern(1.0, ell, 5)
    gp = pygp.inference.ExactGP(pygp.likelihoods.Gaussian(0.1), kern, 0.0)
    gp.add_data(X, y)
    lZ, dlZ = gp.loglikelihood(True)
    assert np.isfinite(lZ) and np.all(np.isfinite(dlZ))
    sn2 = 0.01
    # (1) sampled rows of L L^T against the kernel, through the factor as the reference stores it
    R, a = gp._factor()                      # upper R = L^T (8 GiB on the host), a = R^-T r
    rows = [0, 1, 63, 64, 4097, 20000, n - 1]
    Krows = kern.get(X[rows], X)
    for r_i, krow in zip(rows, Krows):
        rec = R[:r_i + 1, r_i] @ R[:r_i + 1, :]          # (L L^T)[r_i, :] = sum_k R[k, r_i] R[k, :]
        ref = krow.copy()
        ref[r_i] += sn2
        nt.assert_allclose(rec[r_i:], ref[r_i:], rtol=1e-11, atol=1e-12)
    # (2) lZ from its definition
    lZ_def = -0.5*a @ a - np.sum(np.log(np.diag(R))) - 0.5*n*np.log(2*np.pi)
    nt.assert_allclose(lZ, lZ_def, rtol=1e-12)
    # (3) mu(X_i) = mean + k_i^T alpha = y_i - sn2 alpha_i ; alpha = R^-1 a on a slice via back substitution
    import scipy.linalg as sla
    tail = slice(n - 512, n)
    alpha_tail = sla.solve_triangular(R[tail, tail], a[tail])      # last block of R^-1 a
    mu, s2 = gp.posterior(X[tail])
    nt.assert_allclose(mu, y[tail] - sn2*alpha_tail, rtol=1e-8, atol=1e-8)
    assert np.all(s2 > 0) and np.all(s2 < 1.0 + 1e-12)
    del R
    # (4) directional derivative
    h = gp.get_hyper()
    v = np.random.RandomState(5).randn(len(h))
    v /= np.linalg.norm(v)
    eps = 1e-5
    gp.set_hyper(h + eps*v)
    lp = gp.loglikelihood()
    gp.set_hyper(h - eps*v)
    lm = gp.loglikelihood()
    nt.assert_allclose((lp - lm)/(2*eps), dlZ @ v, rtol=2e-6)


@pytest.mark.parametrize('name', EXACT)
def test_posterior_input_gradients(name, golden):
    """posterior(X, grad=True) (exact.py:99-116, reference tests/test_inference.py:159-169)
    against the reference's outputs, and a second test-point set against the oracle."""
    g = golden['gp']
    gp, Xs = build(name)
    mu, s2, dmu, ds2 = gp.posterior(Xs, grad=True)
    assert_pred_close(mu, s2, g[name + '/mu'], g[name + '/s2'])
    sc = max(1.0, np.abs(g[name + '/dmu']).max())
    nt.assert_allclose(dmu, g[name + '/dmu'], rtol=1e-9, atol=1e-10*sc)
    sc = max(1.0, np.abs(g[name + '/ds2']).max())
    nt.assert_allclose(ds2, g[name + '/ds2'], rtol=1e-8, atol=1e-10*sc)
    spec, N, d, _ = GP_CASES[name]
    X, y, _, _ = gp_inputs(N, d)
    ogp = OExactGP(GP_SN, make_kernel(spec), GP_MEAN)
    ogp.add_data(X, y)
    Xt = np.random.RandomState(8).rand(77, d)             # crosses the 64-row tile; d + 1 rows per point
    out = gp.posterior(Xt, grad=True)
    ref = ogp.posterior(Xt, grad=True)
    for a, b, tol in zip(out, ref, (1e-10, 1e-10, 1e-9, 1e-8)):
        nt.assert_allclose(a, b, rtol=tol, atol=tol*max(1.0, np.abs(b).max()))
    # before any data: zero gradients (exact.py:100-104)
    gp.reset()
    out = gp.posterior(Xt[:5], grad=True)
    assert np.all(out[2] == 0) and np.all(out[3] == 0)


@pytest.mark.parametrize('spec,d,n0,adds', [
    (('se', 1.0, [0.5, 0.6]), 2, 10, [10]),                       # reference tests/test_inference.py:65-79
    (('matern', 1.0, [0.7, 0.8, 0.9], 5), 3, 301, [1, 1, 64, 130, 3]),   # odd sizes, across tile edges
    (('sum', ('se', 1.0, 0.5), ('periodic', 0.5, 1.0, 0.25)), 1, 1000, [24]),
    # one to eight new rows on n >= 512: the few-rows TRSM (chol.cu: trsm_fewrows), in place after the first growth
    (('se', 1.0, [0.5, 0.6, 0.7, 0.8]), 4, 700, [1, 1, 5, 8, 3, 2]),
    (('matern', 1.0, [0.4, 0.5], 3), 2, 1537, [1, 7, 1]),
])
def test_incremental_update_equals_full(spec, d, n0, adds):
    """_updateinc (exact.py:57-62 + mwhutils.linalg.chol_update): growing the
    factor by new rows == refactoring from scratch (the reference's test_add_data),
    for lZ, gradient, posterior and the factor itself; and against the oracle."""
    import pygp_b200 as pygp
    ntot = n0 + sum(adds)
    X, y, Xs = synthetic_problem(ntot, d, 40)
    if spec[0] == 'sum':
        X, Xs = X*8, Xs*8
    mk = lambda: pygp.inference.ExactGP(pygp.likelihoods.Gaussian(0.1), product_kernel(spec), 0.1)
    inc = mk()
    inc.add_data(X[:n0], y[:n0])
    lo = n0
    for a in adds:
        inc.add_data(X[lo:lo + a], y[lo:lo + a])      # -> _updateinc -> pgp_exact_append_inc
        lo += a
    full = mk()
    full.add_data(X, y)
    assert inc.ndata == full.ndata == ntot
    lZ, dlZ = inc.loglikelihood(True)
    lZ0, dlZ0 = full.loglikelihood(True)
    nt.assert_allclose(lZ, lZ0, rtol=1e-10)
    assert_grad_close(dlZ, dlZ0)
    mu, s2 = inc.posterior(Xs)
    mu0, s20 = full.posterior(Xs)
    nt.assert_allclose(mu, mu0, rtol=1e-9, atol=1e-10)
    nt.assert_allclose(s2, s20, rtol=1e-8, atol=1e-11)
    nt.assert_allclose(inc._a, full._a, rtol=1e-8, atol=1e-9)
    nt.assert_allclose(inc._R, full._R, rtol=1e-8, atol=1e-10)      # factor buffer with append slack (ld != lead_dim(n))
    cp = inc.copy()                                                  # a clone is an exact fit again
    nt.assert_allclose(cp.loglikelihood(), lZ0, rtol=1e-10)
    ogp = OExactGP(0.1, make_kernel(spec), 0.1)
    ogp.add_data(X, y)
    nt.assert_allclose(lZ, ogp.loglikelihood(), rtol=1e-10)
    # hypers can still be changed afterwards (full refactorisation of the grown data set)
    inc.set_hyper(inc.get_hyper() + 0.01)
    full.set_hyper(full.get_hyper() + 0.01)
    nt.assert_allclose(inc.loglikelihood(), full.loglikelihood(), rtol=1e-12)


def test_incremental_update_not_positive_definite():
    import pygp_b200 as pygp
    X, y, _ = synthetic_problem(50, 1, 0)
    gp = pygp.inference.ExactGP(pygp.likelihoods.Gaussian(1e-9), pygp.kernels.SE(1.0, 5.0, ndim=1), 0.0)
    gp.add_data(X[:3], y[:3])
    with pytest.raises(np.linalg.LinAlgError):
        gp.add_data(np.r_[X[:3], X[:3]], np.r_[y[:3], y[:3]])      # duplicates, no noise: singular
    # the model recovers: a later well-posed call refactors from the host's data
    gp2 = pygp.inference.ExactGP(pygp.likelihoods.Gaussian(0.1), pygp.kernels.SE(1.0, 0.5, ndim=1), 0.0)
    gp2.add_data(X[:20], y[:20])
    gp2.add_data(X[20:], y[20:])
    assert np.isfinite(gp2.loglikelihood())


@pytest.mark.parametrize('name', EXACT)
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


def test_plain_c_client_matches_python_host(tmp_path):
    """tests/c_abi/smoke.c drives the C ABI with no Python in the process; the
    Python host package on the same inputs must give the same numbers."""
    import os
    import subprocess
    import pygp_b200 as pygp
    from pygp_b200 import _lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.dirname(_lib.LIB_PATH)
    exe = str(tmp_path / 'c_smoke')
    subprocess.check_call(['gcc', '-std=c99', '-I', os.path.join(root, 'include'), os.path.join(root, 'tests', 'c_abi', 'smoke.c'),
                           '-o', exe, '-L', libdir, '-lpygp_b200', '-Wl,-rpath,' + libdir, '-lm'])
    p = subprocess.run([exe], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    assert 'not positive definite' in p.stderr            # the singular update came back as info > 0
    vals = np.array([float(v) for v in p.stdout.split()])
    # the client's LCG inputs, regenerated here
    N, D, M = 200, 3, 3
    s, out = 12345, []
    for _ in range(N*D + M*D):
        s = (s*1664525 + 1013904223) % 2**32
        out.append((s >> 8)/16777216.0)
    X, Xs = np.array(out[:N*D]).reshape(N, D), np.array(out[N*D:]).reshape(M, D)
    y = np.sin(3.0*X.sum(1))
    gp = pygp.inference.ExactGP(pygp.likelihoods.Gaussian(0.1), pygp.kernels.SE(1.2, [0.5, 0.6, 0.7]), 0.05)
    gp.add_data(X, y)
    lZ, dlZ = gp.loglikelihood(True)
    mu, s2 = gp.posterior(Xs)
    ref = np.r_[lZ, dlZ, mu, s2]
    nt.assert_allclose(vals, ref, rtol=1e-12, atol=1e-13)
    ogp = OExactGP(0.1, make_kernel(('se', 1.2, [0.5, 0.6, 0.7])), 0.05)
    ogp.add_data(X, y)
    nt.assert_allclose(vals[0], ogp.loglikelihood(), rtol=LZ_RTOL)
