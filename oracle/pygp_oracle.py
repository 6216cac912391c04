"""
TEST INFRASTRUCTURE ONLY -- CPU restatement of the pygp exact-GP hot path.

Every function cites the reference file:line (relative to /root/reference) whose
arithmetic it follows.  The restatement keeps the reference's ORDER of floating
point operations wherever that is observable (direct-difference squared
distances through scipy's cdist, division by the length-scales before the
distance, upper-triangular scipy Cholesky, mean through (R^-T K*)^T a, ...).

Parity status: PINNED.  ``oracle/make_golden.py`` checks this file against the
unmodified reference imported in the build container and commits the
reference's outputs under ``tests/golden/``; ``tests/test_oracle.py`` replays
them, plus the known answers of SURVEY.md section 8c.

Kernels are described by a *spec* (nested tuples) so that the tests can build
the oracle object and the product object from the same description::

    ('se', sf, ell[, ndim])            ('matern', sf, ell, d[, ndim])
    ('periodic', sf, ell, p)           ('rq', sf, ell, alpha[, ndim])
    ('sum', spec, spec, ...)           ('prod', spec, spec, ...)
"""

import numpy as np
import scipy.linalg as sla
import scipy.spatial.distance as ssd

__all__ = ['make_kernel', 'OSE', 'OMatern', 'OPeriodic', 'ORQ', 'OSum',
           'OProduct', 'OExactGP', 'OFITC', 'ODTC', 'synthetic_problem']


# -- distances: pygp/kernels/_distances.py:17-52 ------------------------------

def _rescale(ell, X1, X2):
    # _distances.py:17-23 -- true division of both inputs by ell.
    X1 = X1 / ell
    X2 = X2 / ell if (X2 is not None) else None
    return X1, X2


def _sqdist(X1, X2=None):
    # _distances.py:35-41 -- cdist 'sqeuclidean' = direct sum of (x-y)^2.
    X2 = X1 if (X2 is None) else X2
    return ssd.cdist(X1, X2, 'sqeuclidean')


def _diff(X1, X2=None):
    # _distances.py:26-32 -- pairwise differences X1[:, None, :] - X2[None, :, :]
    X2 = X1 if (X2 is None) else X2
    return X1[:, None, :] - X2[None, :, :]


def _sqdist_foreach(X1, X2=None):
    # _distances.py:44-52 -- one squared-difference matrix per input dimension.
    X2 = X1 if (X2 is None) else X2
    for i in range(X1.shape[1]):
        yield ssd.cdist(X1[:, i, None], X2[:, i, None], 'sqeuclidean')


class _OKernel(object):
    def copy_with(self, hyper):
        import copy
        k = copy.deepcopy(self)
        k.set_hyper(np.asarray(hyper, dtype=float))
        return k


# -- SE: pygp/kernels/se.py:25-74 ---------------------------------------------

class OSE(_OKernel):
    def __init__(self, sf, ell, ndim=None):
        # se.py:26-38
        self._logsf = np.log(float(sf))
        self._logell = np.log(ell)
        self._iso = False
        self.ndim = np.size(self._logell)
        self.nhyper = 1 + np.size(self._logell)
        if ndim is not None:
            if np.size(self._logell) == 1:
                self._logell = float(self._logell)
                self._iso = True
                self.ndim = ndim
            else:
                raise ValueError('ndim only usable with scalar lengthscales')

    def get_hyper(self):
        return np.r_[self._logsf, self._logell]          # se.py:46-47

    def set_hyper(self, hyper):
        self._logsf = hyper[0]                           # se.py:49-51
        self._logell = hyper[1] if self._iso else hyper[1:]

    def get(self, X1, X2=None):
        # se.py:53-55
        X1, X2 = _rescale(np.exp(self._logell), X1, X2)
        return np.exp(self._logsf*2 - _sqdist(X1, X2)/2)

    def grad(self, X1, X2=None):
        # se.py:57-66
        X1, X2 = _rescale(np.exp(self._logell), X1, X2)
        D = _sqdist(X1, X2)
        K = np.exp(self._logsf*2 - D/2)
        out = [2*K]
        if self._iso:
            out.append(K*D)
        else:
            for D in _sqdist_foreach(X1, X2):
                out.append(K*D)
        return out

    def gradx(self, X1, X2=None):
        # se.py:76-83
        ell = np.exp(self._logell)
        X1, X2 = _rescale(ell, X1, X2)
        D = _diff(X1, X2)
        K = np.exp(self._logsf*2 - np.sum(D**2, axis=-1)/2)
        return -K[:, :, None] * D / ell

    def grady(self, X1, X2=None):
        return -self.gradx(X1, X2)      # se.py:85-86

    def gradxy(self, X1, X2=None):
        # se.py:88-99
        ell = np.exp(self._logell)
        X1, X2 = _rescale(ell, X1, X2)
        D = _diff(X1, X2)
        d = D.shape[2]
        K = np.exp(self._logsf*2 - np.sum(D**2, axis=-1)/2)
        D = D / ell
        M = np.eye(d)/ell**2 - D[:, :, None] * D[:, :, :, None]
        return M * K[:, :, None, None]

    def dget(self, X1):
        return np.exp(self._logsf*2) * np.ones(len(X1))  # se.py:68-69

    def dgrad(self, X):
        # se.py:71-74
        return [2 * self.dget(X)] + [np.zeros(len(X))
                                     for _ in range(self.nhyper-1)]


# -- Matern: pygp/kernels/matern.py:25-98 -------------------------------------

class OMatern(_OKernel):
    def __init__(self, sf, ell, d=3, ndim=None):
        # matern.py:26-42
        self._logsf = np.log(float(sf))
        self._logell = np.log(ell)
        self._d = d
        self._iso = False
        self.ndim = np.size(self._logell)
        self.nhyper = 1 + np.size(self._logell)
        if ndim is not None:
            if np.size(self._logell) == 1:
                self._logell = float(self._logell)
                self._iso = True
                self.ndim = ndim
            else:
                raise ValueError('ndim only usable with scalar lengthscales')
        if self._d not in {1, 3, 5}:
            raise ValueError('d must be one of 1, 3, or 5')

    def _f(self, r):
        # matern.py:44-48
        return (1 if (self._d == 1) else
                1+r if (self._d == 3) else
                1+r*(1+r/3.))

    def _df(self, r):
        # matern.py:50-54
        return (1 if (self._d == 1) else
                r if (self._d == 3) else
                r*(1+r)/3.)

    def get_hyper(self):
        return np.r_[self._logsf, self._logell]

    def set_hyper(self, hyper):
        self._logsf = hyper[0]
        self._logell = hyper[1] if self._iso else hyper[1:]

    def get(self, X1, X2=None):
        # matern.py:69-74
        X1, X2 = _rescale(np.exp(self._logell)/np.sqrt(self._d), X1, X2)
        D = np.sqrt(_sqdist(X1, X2))
        S = np.exp(self._logsf*2 - D)
        return S * self._f(D)

    def grad(self, X1, X2=None):
        # matern.py:76-90
        X1, X2 = _rescale(np.exp(self._logell)/np.sqrt(self._d), X1, X2)
        D = np.sqrt(_sqdist(X1, X2))
        S = np.exp(self._logsf*2 - D)
        K = S * self._f(D)
        M = S * self._df(D)
        out = [2*K]
        if self._iso:
            out.append(M*D)
        else:
            for D_ in _sqdist_foreach(X1, X2):
                with np.errstate(invalid='ignore', divide='ignore'):
                    out.append(np.where(D < 1e-12, 0, M*D_/D))
        return out

    def gradx(self, X1, X2=None):
        # matern.py:100-111
        ell = np.exp(self._logell) / np.sqrt(self._d)
        X1, X2 = _rescale(ell, X1, X2)
        D1 = _diff(X1, X2)
        D = np.sqrt(np.sum(D1**2, axis=-1))
        S = np.exp(self._logsf*2 - D)
        with np.errstate(invalid='ignore', divide='ignore'):
            M = np.where(D < 1e-12, 0, S * self._df(D) / D)
        return -M[:, :, None] * D1 / ell

    def grady(self, X1, X2=None):
        return -self.gradx(X1, X2)      # matern.py:113-114

    def gradxy(self, X1, X2=None):
        raise NotImplementedError       # as the reference

    def dget(self, X1):
        return np.exp(self._logsf*2) * np.ones(len(X1))  # matern.py:92-93

    def dgrad(self, X):
        return [2 * self.dget(X)] + [np.zeros(len(X))
                                     for _ in range(self.nhyper-1)]


# -- Periodic: pygp/kernels/periodic.py:24-82 ---------------------------------

class OPeriodic(_OKernel):
    def __init__(self, sf, ell, p):
        # periodic.py:33-38 (one input dimension only)
        self._logsf = np.log(float(sf))
        self._logell = np.log(float(ell))
        self._logp = np.log(float(p))
        self.ndim = 1
        self.nhyper = 3

    def get_hyper(self):
        return np.r_[self._logsf, self._logell, self._logp]

    def set_hyper(self, hyper):
        self._logsf, self._logell, self._logp = hyper[0], hyper[1], hyper[2]

    def get(self, X1, X2=None):
        # periodic.py:53-59
        sf2 = np.exp(self._logsf*2)
        ell = np.exp(self._logell)
        p = np.exp(self._logp)
        D = np.sqrt(_sqdist(X1, X2)) * np.pi / p
        return sf2 * np.exp(-2*(np.sin(D) / ell)**2)

    def grad(self, X1, X2=None):
        # periodic.py:61-74
        sf2 = np.exp(self._logsf*2)
        ell = np.exp(self._logell)
        p = np.exp(self._logp)
        D = np.sqrt(_sqdist(X1, X2)) * np.pi / p
        R = np.sin(D) / ell
        S = R**2
        E = 2 * sf2 * np.exp(-2*S)
        return [E, 2*E*S, 2*E*R*D * np.cos(D) / ell]

    def gradx(self, X1, X2=None):
        # periodic.py:84-94
        sf2 = np.exp(self._logsf*2)
        ell = np.exp(self._logell)
        p = np.exp(self._logp)
        D = _diff(X1, X2) * np.pi / p
        K = sf2 * np.exp(-2*(np.sin(D) / ell)**2)
        return -2 * np.pi / ell**2 / p * K * np.sin(2*D)

    def grady(self, X1, X2=None):
        return -self.gradx(X1, X2)      # periodic.py:96-97

    def gradxy(self, X1, X2=None):
        raise NotImplementedError       # as the reference

    def dget(self, X1):
        return np.exp(self._logsf*2) * np.ones(len(X1))  # periodic.py:76-77

    def dgrad(self, X):
        return [2 * self.dget(X), np.zeros(len(X)), np.zeros(len(X))]


# -- RQ: pygp/kernels/rq.py:24-93 ---------------------------------------------

class ORQ(_OKernel):
    def __init__(self, sf, ell, alpha, ndim=None):
        # rq.py:25-40
        self._logsf = np.log(float(sf))
        self._logell = np.log(ell)
        self._logalpha = np.log(float(alpha))
        self._iso = False
        self.ndim = np.size(self._logell)
        self.nhyper = 2 + np.size(self._logell)
        if ndim is not None:
            if np.size(self._logell) == 1:
                self._logell = float(self._logell)
                self._iso = True
                self.ndim = ndim
            else:
                raise ValueError('ndim only usable with scalar lengthscales')

    def get_hyper(self):
        return np.r_[self._logsf, self._logell, self._logalpha]

    def set_hyper(self, hyper):
        self._logsf = hyper[0]
        self._logell = hyper[1] if self._iso else hyper[1:-1]
        self._logalpha = hyper[-1]

    def get(self, X1, X2=None):
        # rq.py:56-63
        sf2 = np.exp(self._logsf*2)
        ell = np.exp(self._logell)
        alpha = np.exp(self._logalpha)
        X1, X2 = _rescale(ell, X1, X2)
        return sf2 * (1 + 0.5*_sqdist(X1, X2)/alpha) ** (-alpha)

    def grad(self, X1, X2=None):
        # rq.py:65-84
        sf2 = np.exp(self._logsf*2)
        ell = np.exp(self._logell)
        alpha = np.exp(self._logalpha)
        X1, X2 = _rescale(ell, X1, X2)
        D = _sqdist(X1, X2)
        E = 1 + 0.5*D/alpha
        K = sf2 * E**(-alpha)
        M = K*D/E
        out = [2*K]
        if self._iso:
            out.append(M)
        else:
            for D in _sqdist_foreach(X1, X2):
                out.append(K*D/E)
        out.append(0.5*M - alpha*K*np.log(E))
        return out

    def gradx(self, X1, X2=None):
        # rq.py:95-108
        sf2 = np.exp(self._logsf*2)
        ell = np.exp(self._logell)
        alpha = np.exp(self._logalpha)
        X1, X2 = _rescale(ell, X1, X2)
        D = _diff(X1, X2)
        E = 1 + np.sum(D**2, axis=-1) / 2 / alpha
        K = sf2 * E**(-alpha)
        return -(K/E)[:, :, None] * D / ell

    def grady(self, X1, X2=None):
        return -self.gradx(X1, X2)      # rq.py:110-111

    def gradxy(self, X1, X2=None):
        raise NotImplementedError       # as the reference

    def dget(self, X1):
        return np.exp(self._logsf*2) * np.ones(len(X1))  # rq.py:86-87

    def dgrad(self, X):
        return [2 * self.dget(X)] + [np.zeros(len(X))
                                     for _ in range(self.nhyper-1)]


# -- composites: pygp/kernels/_combo.py:25-160, _real.py:76-117 ---------------

def _product_but(A):
    """_combo.py:32-51: M[i] = product of every A[j] with j != i, by
    cumulative products (no division)."""
    A = list(A)
    M = np.empty_like(A)
    np.cumprod(A[:0:-1], axis=0, out=M[:-1][::-1])
    M[-1] = A[0]
    for i in range(1, len(A)-1):
        M[i] *= M[-1]
        M[-1] *= A[i]
    return M


class _OCombo(_OKernel):
    def __init__(self, *parts):
        import copy
        if not all(p.ndim == parts[0].ndim for p in parts):   # _real.py:76-83
            raise ValueError('cannot combine mismatched kernels')
        # _combo.py:151-160 / _real.py:32-36: same-type nesting is flattened.
        flat = []
        for p in parts:
            flat += p._parts if isinstance(p, type(self)) else [p]
        self._parts = [copy.deepcopy(p) for p in flat]        # _combo.py:62
        self.nhyper = sum(p.nhyper for p in self._parts)
        self.ndim = self._parts[0].ndim

    def get_hyper(self):
        return np.hstack([p.get_hyper() for p in self._parts])  # _combo.py:90-91

    def set_hyper(self, hyper):
        a = 0                                                 # _combo.py:93-98
        for p in self._parts:
            b = a + p.nhyper
            p.set_hyper(hyper[a:b])
            a = b


class OSum(_OCombo):
    # _combo.py:103-120
    def get(self, X1, X2=None):
        return sum(p.get(X1, X2) for p in self._parts)

    def gradx(self, X1, X2=None):
        return sum(p.gradx(X1, X2) for p in self._parts)        # _real.py:96-97

    def grady(self, X1, X2=None):
        return sum(p.grady(X1, X2) for p in self._parts)        # _real.py:99-100

    def gradxy(self, X1, X2=None):
        return sum(p.gradxy(X1, X2) for p in self._parts)       # _real.py:102-103

    def dget(self, X):
        return sum(p.dget(X) for p in self._parts)

    def grad(self, X1, X2=None):
        return [g for p in self._parts for g in p.grad(X1, X2)]

    def dgrad(self, X):
        return [g for p in self._parts for g in p.dgrad(X)]


class OProduct(_OCombo):
    # _combo.py:123-146
    def get(self, X1, X2=None):
        out = 1
        for p in self._parts:
            out = out * p.get(X1, X2)
        return out

    def gradx(self, X1, X2=None):
        # _real.py:119-122
        F = _product_but([p.get(X1, X2)[:, :, None] for p in self._parts])
        return sum(f*p.gradx(X1, X2) for f, p in zip(F, self._parts))

    def grady(self, X1, X2=None):
        # _real.py:124-127
        F = _product_but([p.get(X1, X2)[:, :, None] for p in self._parts])
        return sum(f*p.grady(X1, X2) for f, p in zip(F, self._parts))

    def gradxy(self, X1, X2=None):
        # _real.py:129-156
        K = [p.get(X1, X2) for p in self._parts]
        Kn = _product_but(K)
        Gx = [p.gradx(X1, X2) for p in self._parts]
        Gy = [p.grady(X1, X2) for p in self._parts]
        Gxy = [p.gradxy(X1, X2) for p in self._parts]
        grad = sum(Kni[:, :, None, None] * dKi for Kni, dKi in zip(Kn, Gxy))
        xpart = sum(dKi * Ki[:, :, None] for dKi, Ki in zip(Gx, Kn))
        ypart = sum(dKi / Ki[:, :, None] for dKi, Ki in zip(Gy, K))
        grad += xpart[:, :, :, None] * ypart[:, :, None, :]
        grad -= sum((Kni / Ki)[:, :, None, None] * dKx[:, :, :, None] * dKy[:, :, None, :]
                    for Kni, Ki, dKx, dKy in zip(Kn, K, Gx, Gy))
        return grad

    def dget(self, X):
        out = 1
        for p in self._parts:
            out = out * p.dget(X)
        return out

    def grad(self, X1, X2=None):
        F = [p.get(X1, X2) for p in self._parts]
        out = []
        for Mi, p in zip(_product_but(F), self._parts):
            for dM in p.grad(X1, X2):
                out.append(Mi*dM)
        return out

    def dgrad(self, X):
        F = [p.dget(X) for p in self._parts]
        out = []
        for Mi, p in zip(_product_but(F), self._parts):
            for dM in p.dgrad(X):
                out.append(Mi*dM)
        return out


def make_kernel(spec):
    """Build an oracle kernel from the tuple description in the module doc."""
    tag = spec[0]
    if tag == 'se':
        return OSE(*spec[1:])
    if tag == 'matern':
        return OMatern(*spec[1:])
    if tag == 'periodic':
        return OPeriodic(*spec[1:])
    if tag == 'rq':
        return ORQ(*spec[1:])
    if tag == 'sum':
        return OSum(*[make_kernel(s) for s in spec[1:]])
    if tag == 'prod':
        return OProduct(*[make_kernel(s) for s in spec[1:]])
    raise ValueError('unknown kernel tag %r' % (tag,))


def _gp_sample(gp, X, m, latent, rng):
    # _base.py:143-177 (rstate: int seed -> RandomState(seed)); Gaussian.sample gaussian.py:47-49
    X = np.array(X, ndmin=2, dtype=float)
    flatten = (m is None)
    m = 1 if flatten else m
    n = len(X)
    rng = rng if isinstance(rng, np.random.RandomState) else np.random.RandomState(rng)
    mu, Sigma = gp.full_posterior(X)
    Sigma += 1e-10 * np.eye(n)
    f = mu[None] + np.dot(rng.normal(size=(m, n)), sla.cholesky(Sigma))
    if not latent:
        f = (f.ravel() + rng.normal(size=f.size, scale=np.sqrt(gp.s2))).reshape(m, n)
    return f.ravel() if flatten else f


# -- ExactGP: pygp/inference/exact.py:20-143, _base.py:47-141 -----------------

class OExactGP(object):
    """Hyper vector layout [log sn | kernel hypers | mean] (_base.py:91-105)."""

    def __init__(self, sn, kernel, mean=0.0):
        self._logsn = np.log(float(sn))      # likelihoods/gaussian.py:27
        self._kernel = kernel
        self._mean = float(mean)
        self._X = self._y = self._R = self._a = None
        self.nhyper = 1 + kernel.nhyper + 1  # _base.py:56-57

    @property
    def s2(self):
        return np.exp(self._logsn*2)         # gaussian.py:36-39

    @property
    def ndata(self):
        return 0 if self._X is None else self._X.shape[0]

    def get_hyper(self):
        return np.r_[self._logsn, self._kernel.get_hyper(), self._mean]

    def set_hyper(self, hyper):
        # _base.py:98-108
        hyper = np.asarray(hyper, dtype=float)
        self._logsn = hyper[0]
        self._kernel.set_hyper(hyper[1:1+self._kernel.nhyper])
        self._mean = hyper[-1]
        if self.ndata > 0:
            self._update()

    def add_data(self, X, y):
        # _base.py:120-141 (full-update branch; _updateinc is out of scope)
        X = np.array(X, ndmin=2, dtype=float)
        y = np.array(y, ndmin=1, dtype=float)
        if self._X is None:
            self._X, self._y = X.copy(), y.copy()
        else:
            self._X = np.r_[self._X, X]
            self._y = np.r_[self._y, y]
        self._update()

    def _update(self):
        # exact.py:50-55
        sn2 = self.s2
        K = self._kernel.get(self._X) + sn2 * np.eye(len(self._X))
        r = self._y - self._mean
        self._R = sla.cholesky(K)
        self._a = sla.solve_triangular(self._R, r, trans=True)

    def loglikelihood(self, grad=False):
        # exact.py:118-143
        lZ = -0.5 * np.inner(self._a, self._a)
        lZ -= 0.5 * np.log(2 * np.pi) * self.ndata
        lZ -= np.sum(np.log(self._R.diagonal()))
        if not grad:
            return lZ
        alpha = sla.solve_triangular(self._R, self._a, trans=False)
        Q = sla.cho_solve((self._R, False), np.eye(self.ndata))
        Q -= np.outer(alpha, alpha)
        dlZ = np.r_[
            -self.s2 * np.trace(Q),
            [-0.5*np.sum(Q*dK) for dK in self._kernel.grad(self._X)],
            np.sum(alpha)]
        return lZ, dlZ

    def full_posterior(self, X):
        # exact.py:64-79
        X = np.array(X, ndmin=2, dtype=float)
        mu = np.full(X.shape[0], self._mean)
        Sigma = self._kernel.get(X)
        if self._X is not None:
            K = self._kernel.get(self._X, X)
            V = sla.solve_triangular(self._R, K, trans=True)
            mu += np.dot(V.T, self._a)
            Sigma -= np.dot(V.T, V)
        return mu, Sigma

    def sample(self, X, m=None, latent=True, rng=None):
        return _gp_sample(self, X, m, latent, rng)

    def posterior(self, X, grad=False):
        # exact.py:81-116, via _base.py:179-186
        X = np.array(X, ndmin=2, dtype=float)
        mu = np.full(X.shape[0], self._mean)
        s2 = self._kernel.dget(X)
        if self._X is not None:
            K = self._kernel.get(self._X, X)
            RK = sla.solve_triangular(self._R, K, trans=True)
            mu += np.dot(RK.T, self._a)
            s2 -= np.sum(RK**2, axis=0)
        if not grad:
            return mu, s2
        dmu = np.zeros_like(X)
        ds2 = np.zeros_like(X)
        if self._X is not None:
            dK = self._kernel.grady(self._X, X)
            dK = dK.reshape(self.ndata, -1)
            RdK = sla.solve_triangular(self._R, dK, trans=True)
            dmu += np.dot(RdK.T, self._a).reshape(X.shape)
            RdK = np.rollaxis(np.reshape(RdK, (-1,) + X.shape), 2)
            ds2 -= 2 * np.sum(RdK * RK, axis=1).T
        return mu, s2, dmu, ds2


# -- FITC: pygp/inference/fitc.py:19-232 --------------------------------------

class OFITC(object):
    def __init__(self, sn, kernel, mean, U):
        self._logsn = np.log(float(sn))
        self._kernel = kernel
        self._mean = float(mean)
        self._U = np.array(U, ndmin=2, dtype=float, copy=True)   # fitc.py:31
        self._X = self._y = None
        self._L = self._R = self._b = self._A = self._a = None
        self.nhyper = 1 + kernel.nhyper + 1

    @property
    def s2(self):
        return np.exp(self._logsn*2)

    @property
    def ndata(self):
        return 0 if self._X is None else self._X.shape[0]

    def get_hyper(self):
        return np.r_[self._logsn, self._kernel.get_hyper(), self._mean]

    def set_hyper(self, hyper):
        hyper = np.asarray(hyper, dtype=float)
        self._logsn = hyper[0]
        self._kernel.set_hyper(hyper[1:1+self._kernel.nhyper])
        self._mean = hyper[-1]
        if self.ndata > 0:
            self._update()

    def add_data(self, X, y):
        X = np.array(X, ndmin=2, dtype=float)
        y = np.array(y, ndmin=1, dtype=float)
        if self._X is None:
            self._X, self._y = X.copy(), y.copy()
        else:
            self._X = np.r_[self._X, X]
            self._y = np.r_[self._y, y]
        self._update()

    def _update(self):
        # fitc.py:66-100
        sn2 = self.s2
        su2 = sn2 / 1e6
        Kuu = self._kernel.get(self._U)
        p = self._U.shape[0]
        self._L = sla.cholesky(Kuu + su2*np.eye(p))
        Kux = self._kernel.get(self._U, self._X)
        kxx = self._kernel.dget(self._X)
        r = self._y - self._mean
        V = sla.solve_triangular(self._L, Kux, trans=True)
        ell = np.sqrt(kxx + sn2 - np.sum(V**2, axis=0))
        Kux /= ell
        V /= ell
        r /= ell
        self._A = np.eye(p) + np.dot(V, V.T)
        self._a = np.dot(Kux, r)
        self._R = np.dot(sla.cholesky(self._A), self._L)
        self._b = sla.solve_triangular(self._R, self._a, trans=True)

    def full_posterior(self, X):
        # fitc.py:102-120
        X = np.array(X, ndmin=2, dtype=float)
        mu = np.full(X.shape[0], self._mean)
        Sigma = self._kernel.get(X)
        if self._X is not None:
            K = self._kernel.get(self._U, X)
            LK = sla.solve_triangular(self._L, K, trans=True)
            RK = sla.solve_triangular(self._R, K, trans=True)
            mu += np.dot(RK.T, self._b)
            Sigma += np.dot(RK.T, RK) - np.dot(LK.T, LK)
        return mu, Sigma

    def sample(self, X, m=None, latent=True, rng=None):
        return _gp_sample(self, X, m, latent, rng)

    def posterior(self, X, grad=False):
        # fitc.py:122-165
        X = np.array(X, ndmin=2, dtype=float)
        mu = np.full(X.shape[0], self._mean)
        s2 = self._kernel.dget(X)
        if self._X is not None:
            K = self._kernel.get(self._U, X)
            LK = sla.solve_triangular(self._L, K, trans=True)
            RK = sla.solve_triangular(self._R, K, trans=True)
            mu += np.dot(RK.T, self._b)
            s2 += np.sum(RK**2, axis=0) - np.sum(LK**2, axis=0)
        if not grad:
            return mu, s2
        dmu = np.zeros_like(X)
        ds2 = np.zeros_like(X)
        if self._X is not None:
            p = self._U.shape[0]
            dK = self._kernel.grady(self._U, X)
            dK = dK.reshape(p, -1)
            LdK = sla.solve_triangular(self._L, dK, trans=True)
            RdK = sla.solve_triangular(self._R, dK, trans=True)
            dmu += np.dot(RdK.T, self._b).reshape(X.shape)
            LdK = np.rollaxis(np.reshape(LdK, (p,) + X.shape), 2)
            RdK = np.rollaxis(np.reshape(RdK, (p,) + X.shape), 2)
            ds2 += 2 * np.sum(RdK * RK, axis=1).T
            ds2 -= 2 * np.sum(LdK * LK, axis=1).T
        return mu, s2, dmu, ds2

    def loglikelihood(self, grad=False):
        # fitc.py:167-232
        sn2 = self.s2
        su2 = sn2 / 1e6
        Kux = self._kernel.get(self._U, self._X)
        kxx = self._kernel.dget(self._X)
        r = self._y - self._mean
        V = sla.solve_triangular(self._L, Kux, trans=True)
        ell = np.sqrt(kxx + sn2 - np.sum(V**2, axis=0))
        V /= ell
        r /= ell
        A = sla.cholesky(self._A)
        beta = sla.solve_triangular(A, V.dot(r), trans=True)
        alpha = (r - V.T.dot(sla.solve_triangular(A, beta))) / ell
        lZ = -np.sum(np.log(np.diag(A))) - np.sum(np.log(ell))
        lZ -= 0.5 * (np.inner(r, r) - np.inner(beta, beta))
        lZ -= 0.5 * ell.shape[0] * np.log(2*np.pi)
        if not grad:
            return lZ

        B = sla.solve_triangular(self._L, V*ell)
        W = sla.solve_triangular(A, V/ell, trans=True)
        w = B.dot(alpha)
        v = 2*su2*np.sum(B**2, axis=0)
        dlZ = np.zeros(self.nhyper)
        dlZ[0] = (
            - sn2 * (np.sum(1/ell**2) - np.sum(W**2) - np.inner(alpha, alpha))
            - su2 * (np.sum(w**2) + np.sum(B.dot(W.T)**2))
            + 0.5 * (
                np.inner(alpha, v*alpha) + np.inner(np.sum(W**2, axis=0), v)))
        dK = zip(self._kernel.grad(self._U),
                 self._kernel.grad(self._U, self._X),
                 self._kernel.dgrad(self._X))
        for i, (dKuu, dKux, dkxx) in enumerate(dK, 1):
            M = 2*dKux - dKuu.dot(B)
            v = dkxx - np.sum(M*B, axis=0)
            dlZ[i] = (
                - np.sum(dkxx/ell**2)
                - np.inner(w, dKuu.dot(w) - 2*dKux.dot(alpha))
                + np.inner(alpha, v*alpha) + np.inner(np.sum(W**2, axis=0), v)
                + np.sum(M.dot(W.T) * B.dot(W.T))) / 2.0
        dlZ[-1] = np.sum(alpha)
        return lZ, dlZ


# -- DTC: pygp/inference/dtc.py:20-199 ------------------------------------------

class ODTC(object):
    def __init__(self, sn, kernel, mean, U):
        self._logsn = np.log(float(sn))
        self._kernel = kernel
        self._mean = float(mean)
        self._U = np.array(U, ndmin=2, dtype=float, copy=True)      # dtc.py:31
        self._X = self._y = None
        self._Ruu = self._Rux = self._a = None
        self.nhyper = 1 + kernel.nhyper + 1

    @property
    def s2(self):
        return np.exp(self._logsn*2)

    @property
    def ndata(self):
        return 0 if self._X is None else self._X.shape[0]

    def get_hyper(self):
        return np.r_[self._logsn, self._kernel.get_hyper(), self._mean]

    def set_hyper(self, hyper):
        hyper = np.asarray(hyper, dtype=float)
        self._logsn = hyper[0]
        self._kernel.set_hyper(hyper[1:1+self._kernel.nhyper])
        self._mean = hyper[-1]
        if self.ndata > 0:
            self._update()

    def add_data(self, X, y):
        X = np.array(X, ndmin=2, dtype=float)
        y = np.array(y, ndmin=1, dtype=float)
        if self._X is None:
            self._X, self._y = X.copy(), y.copy()
        else:
            self._X = np.r_[self._X, X]
            self._y = np.r_[self._y, y]
        self._update()

    def _update(self):
        # dtc.py:54-75
        p = self._U.shape[0]
        su2 = self.s2 * 1e-6
        Kuu = self._kernel.get(self._U)
        self._Ruu = sla.cholesky(Kuu + su2 * np.eye(p))
        Kux = self._kernel.get(self._U, self._X)
        S = Kuu + np.dot(Kux, Kux.T) / self.s2
        r = self._y - self._mean
        self._Rux = sla.cholesky(S + su2 * np.eye(p))
        self._a = sla.solve_triangular(self._Rux, np.dot(Kux, r), trans=True)

    def full_posterior(self, X):
        # dtc.py:77-92
        X = np.array(X, ndmin=2, dtype=float)
        mu = np.full(X.shape[0], self._mean)
        Sigma = self._kernel.get(X)
        if self._X is not None:
            K = self._kernel.get(self._U, X)
            b = sla.solve_triangular(self._Ruu, K, trans=True)
            c = sla.solve_triangular(self._Rux, K, trans=True)
            mu += np.dot(c.T, self._a) / self.s2
            Sigma += -np.dot(b.T, b) + np.dot(c.T, c)
        return mu, Sigma

    def sample(self, X, m=None, latent=True, rng=None):
        return _gp_sample(self, X, m, latent, rng)

    def posterior(self, X, grad=False):
        # dtc.py:94-135 (the input-gradient of the mean is NOT divided by sn2 in the
        # reference, dtc.py:127 -- kept, parity is with the reference as it is)
        X = np.array(X, ndmin=2, dtype=float)
        mu = np.full(X.shape[0], self._mean)
        s2 = self._kernel.dget(X)
        if self._X is not None:
            K = self._kernel.get(self._U, X)
            b = sla.solve_triangular(self._Ruu, K, trans=True)
            c = sla.solve_triangular(self._Rux, K, trans=True)
            mu += np.dot(c.T, self._a) / self.s2
            s2 += -np.sum(b * b, axis=0) + np.sum(c * c, axis=0)
        if not grad:
            return mu, s2
        dmu = np.zeros_like(X)
        ds2 = np.zeros_like(X)
        if self._X is not None:
            dK = self._kernel.grady(self._U, X)
            dK = dK.reshape(self._U.shape[0], -1)
            db = sla.solve_triangular(self._Ruu, dK, trans=True)
            db = np.rollaxis(np.reshape(db, (-1,) + X.shape), 2)
            dc = sla.solve_triangular(self._Rux, dK, trans=True)
            dmu += np.dot(dc.T, self._a).reshape(X.shape)
            dc = np.rollaxis(np.reshape(dc, (-1,) + X.shape), 2)
            ds2 += -2 * np.sum(db * b, axis=1).T + 2 * np.sum(dc * c, axis=1).T
        return mu, s2, dmu, ds2

    def loglikelihood(self, grad=False):
        # dtc.py:137-199
        sn2 = self.s2
        su2 = sn2 * 1e-6
        ell = np.sqrt(sn2)
        Kux = self._kernel.get(self._U, self._X)
        r = self._y.copy() - self._mean
        r /= ell
        V = sla.solve_triangular(self._Ruu, Kux, trans=True)
        V /= ell
        p = self._U.shape[0]
        A = sla.cholesky(np.eye(p) + np.dot(V, V.T))
        beta = sla.solve_triangular(A, V.dot(r), trans=True)
        lZ = -np.sum(np.log(np.diag(A))) - self.ndata * np.log(ell)
        lZ -= 0.5 * (np.inner(r, r) - np.inner(beta, beta))
        lZ -= 0.5 * self.ndata * np.log(2*np.pi)
        if not grad:
            return lZ
        alpha = (r - V.T.dot(sla.solve_triangular(A, beta)))
        B = sla.solve_triangular(self._Ruu, V)
        W = sla.solve_triangular(A, V, trans=True)
        VW = np.dot(V, W.T)
        BW = np.dot(B, W.T)
        w = B.dot(alpha)
        v = V.dot(alpha)
        dlZ = np.zeros(self.nhyper)
        dlZ[0] = -(- np.inner(r, r) + np.inner(beta, beta) + np.inner(v, v) + su2 * np.inner(w, w)
                   + self.ndata - np.sum(V**2) + np.sum(VW**2) - su2 * (np.sum(B**2) - np.sum(BW**2)))
        dK = zip(self._kernel.grad(self._U), self._kernel.grad(self._U, self._X))
        for i, (dKuu, dKux) in enumerate(dK, 1):
            M = 2 * dKux / ell - dKuu.dot(B)
            dlZ[i] = -0.5 * (- np.inner(w, np.dot(M, alpha)) + np.sum(M*B) - np.sum(M.dot(W.T) * B.dot(W.T)))
        dlZ[-1] = np.sum(alpha) / ell
        return lZ, dlZ


# -- synthetic inputs of SURVEY.md section 8d ---------------------------------

def synthetic_problem(n, d, m=0, sn=0.1, seed=0, seed_test=1):
    """X = rand(n,d); y = sin(3 sum x) + sn randn; Xs = rand(m,d) (8d)."""
    rng = np.random.RandomState(seed)
    X = rng.rand(n, d)
    y = np.sin(3*X.sum(1)) + sn*rng.randn(n)
    Xs = np.random.RandomState(seed_test).rand(m, d) if m else None
    return X, y, Xs
