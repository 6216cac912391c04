"""
TEST INFRASTRUCTURE ONLY -- a numpy model of the DEVICE formulation of FITC
(pygp_b200/csrc/fitc.cu), not of the reference.  Same operand layout (rows =
data points, "t" suffix = transposed w.r.t. pygp/inference/fitc.py), same
regrouping of the gradient into three elementwise traces.  It shows that the
regrouped algebra reaches the parity tolerance against the oracle
(tests/test_oracle.py::test_fitc_device_model) before any CUDA runs, and gives
stage-by-stage intermediates to diff against device dumps.

Reference algebra: pygp/inference/fitc.py:66-100 (_update), :122-142
(posterior), :167-232 (loglikelihood).  With Lc = L^T (lower), Al = chol(A)^T:

  Kxu = k(X, U)                   (n, p)
  Vt  = Kxu Lc^-T                 (n, p)   = V^T
  ell = sqrt(kxx + sn2 - rowsum(Vt^2));  Vs = Vt/ell;  rs = r/ell
  A   = I + Vs^T Vs;  a = (Kxu/ell)^T rs;  Rl = Lc Al (= R^T);  b = Rl^-1 a
  beta = Al^-1 (Vs^T rs);  alpha = (rs - Vs Al^-T beta)/ell
  Bt  = (Vs ell) Lc^-1 ;  Wt = (Vs/ell) Al^-T ;  w = Bt^T alpha
  q   = alpha^2 + rowsum(Wt^2);  P = Bt^T Wt;  T2 = Wt P^T
  Cuu = Bt^T diag(q) Bt - P P^T - w w^T
  Cxu = 2 (alpha w^T - diag(q) Bt + T2)
  dlZ[h] = 1/2 [ dk_h(0) sum(q - 1/ell^2) + sum(dKuu_h o Cuu) + sum(dKxu_h o Cxu) ]
"""

import numpy as np
import scipy.linalg as sla


def fitc_update(kernel, sn2, mean, U, X, y):
    su2 = sn2 / 1e6
    p = U.shape[0]
    Lc = sla.cholesky(kernel.get(U) + su2*np.eye(p), lower=True)
    Kxu = kernel.get(X, U)
    kxx = kernel.dget(X)
    r = y - mean
    Vt = sla.solve_triangular(Lc, Kxu.T, lower=True).T
    ell = np.sqrt(kxx + sn2 - np.sum(Vt**2, axis=1))
    Vs = Vt / ell[:, None]
    rs = r / ell
    A = np.eye(p) + Vs.T.dot(Vs)
    a = (Kxu / ell[:, None]).T.dot(rs)
    Al = sla.cholesky(A, lower=True)
    Rl = Lc.dot(Al)
    b = sla.solve_triangular(Rl, a, lower=True)
    return dict(Lc=Lc, Al=Al, Rl=Rl, b=b, Vs=Vs, ell=ell, rs=rs, sn2=sn2, su2=su2, mean=mean)


def fitc_predict(kernel, st, U, Xs):
    Ksu = kernel.get(Xs, U)
    LKt = sla.solve_triangular(st['Lc'], Ksu.T, lower=True).T
    RKt = sla.solve_triangular(st['Rl'], Ksu.T, lower=True).T
    mu = st['mean'] + RKt.dot(st['b'])
    s2 = kernel.dget(Xs) + (np.sum(RKt**2, axis=1) - np.sum(LKt**2, axis=1))
    return mu, s2


def fitc_loglike(kernel, st, U, X, grad=False):
    Lc, Al, Vs, ell, rs = st['Lc'], st['Al'], st['Vs'], st['ell'], st['rs']
    sn2, su2 = st['sn2'], st['su2']
    n = X.shape[0]
    beta = sla.solve_triangular(Al, Vs.T.dot(rs), lower=True)
    t = sla.solve_triangular(Al, beta, lower=True, trans=1)          # Al^-T beta
    alpha = (rs - Vs.dot(t)) / ell
    lZ = -np.sum(np.log(np.diag(Al))) - np.sum(np.log(ell))
    lZ -= 0.5*(rs.dot(rs) - beta.dot(beta))
    lZ -= 0.5*n*np.log(2*np.pi)
    if not grad:
        return lZ

    G = sla.solve_triangular(Lc, np.eye(len(Lc)), lower=True).T      # Lc^-T, upper
    Bt = (Vs*ell[:, None]).dot(G.T)                                   # NT GEMM with G
    Wt = sla.solve_triangular(Al, (Vs/ell[:, None]).T, lower=True).T  # (Vs/ell) Al^-T
    w = Bt.T.dot(alpha)
    cw = np.sum(Wt**2, axis=1)
    bb = np.sum(Bt**2, axis=1)
    q = alpha**2 + cw
    P = Bt.T.dot(Wt)
    T2 = Wt.dot(P.T)
    Cuu = Bt.T.dot(Bt*q[:, None]) - P.dot(P.T) - np.outer(w, w)
    Cxu = 2*(np.outer(alpha, w) - Bt*q[:, None] + T2)

    nk = kernel.nhyper
    dlZ = np.zeros(nk + 2)
    v = 2*su2*bb
    dlZ[0] = (-sn2*(np.sum(1/ell**2) - np.sum(cw) - alpha.dot(alpha))
              - su2*(w.dot(w) + np.sum(P**2))
              + 0.5*np.sum(v*q))
    sq = np.sum(q - 1/ell**2)
    dKuu = kernel.grad(U)
    dKxu = kernel.grad(X, U)
    dk0 = kernel.dgrad(X[:1])
    for h, (duu, dxu, d0) in enumerate(zip(dKuu, dKxu, dk0), 1):
        dlZ[h] = 0.5*(d0[0]*sq + np.sum(duu*Cuu) + np.sum(dxu*Cxu))
    dlZ[-1] = np.sum(alpha)
    return lZ, dlZ


# -- DTC (pygp/inference/dtc.py:54-199) in the same device formulation ----------------
#   Luu = chol(Kuu + su2 I) (lower);  Lux = chol(Kuu + su2 I + Kxu^T Kxu / sn2),  a = Lux^-1 Kxu^T r
#   Vs = Kxu Luu^-T / ell, ell = sqrt(sn2);  rs = r / ell;  Al = chol(I + Vs^T Vs);  beta = Al^-1 Vs^T rs
#   alpha = rs - Vs Al^-T beta;  Bt = Vs Luu^-1;  Wt = Vs Al^-T;  w = Bt^T alpha;  v = Vs^T alpha
#   P = Bt^T Wt;  VW = Vs^T Wt;  C = Bt^T Bt - P P^T - w w^T;  T2 = Wt P^T
#   dlZ_h = 1/2 sum(dKuu_h o C) + (1 / (2 ell)) sum(dKxu_h o 2 (alpha w^T - Bt + T2))

def dtc_update(kernel, sn2, mean, U, X, y):
    su2 = sn2 * 1e-6
    p = U.shape[0]
    Kuu = kernel.get(U)
    Luu = sla.cholesky(Kuu + su2*np.eye(p), lower=True)
    Kxu = kernel.get(X, U)
    r = y - mean
    Lux = sla.cholesky(Kuu + su2*np.eye(p) + Kxu.T.dot(Kxu)/sn2, lower=True)
    a = sla.solve_triangular(Lux, Kxu.T.dot(r), lower=True)
    ell = np.sqrt(sn2)
    Vs = sla.solve_triangular(Luu, Kxu.T, lower=True).T / ell
    rs = r / ell
    Al = sla.cholesky(np.eye(p) + Vs.T.dot(Vs), lower=True)
    beta = sla.solve_triangular(Al, Vs.T.dot(rs), lower=True)
    return dict(Luu=Luu, Lux=Lux, a=a, Vs=Vs, rs=rs, Al=Al, beta=beta, sn2=sn2, su2=su2, ell=ell, mean=mean)


def dtc_predict(kernel, st, U, Xs):
    Ksu = kernel.get(Xs, U)
    b = sla.solve_triangular(st['Luu'], Ksu.T, lower=True).T
    c = sla.solve_triangular(st['Lux'], Ksu.T, lower=True).T
    mu = st['mean'] + c.dot(st['a'])/st['sn2']
    s2 = kernel.dget(Xs) + (np.sum(c**2, axis=1) - np.sum(b**2, axis=1))
    return mu, s2


def dtc_loglike(kernel, st, U, X, grad=False):
    Luu, Al, Vs, rs, beta = st['Luu'], st['Al'], st['Vs'], st['rs'], st['beta']
    su2, ell = st['su2'], st['ell']
    n = X.shape[0]
    lZ = -np.sum(np.log(np.diag(Al))) - n*np.log(ell) - 0.5*(rs.dot(rs) - beta.dot(beta)) - 0.5*n*np.log(2*np.pi)
    if not grad:
        return lZ
    alpha = rs - Vs.dot(sla.solve_triangular(Al, beta, lower=True, trans=1))
    Bt = sla.solve_triangular(Luu, Vs.T, lower=True, trans=1).T          # Vs Luu^-1
    Wt = sla.solve_triangular(Al, Vs.T, lower=True).T                    # Vs Al^-T
    w, v = Bt.T.dot(alpha), Vs.T.dot(alpha)
    P, VW = Bt.T.dot(Wt), Vs.T.dot(Wt)
    C = Bt.T.dot(Bt) - P.dot(P.T) - np.outer(w, w)
    Cxu = 2*(np.outer(alpha, w) - Bt + Wt.dot(P.T))
    nk = kernel.nhyper
    dlZ = np.zeros(nk + 2)
    dlZ[0] = -(-rs.dot(rs) + beta.dot(beta) + v.dot(v) + su2*w.dot(w) + n - np.sum(Vs**2) + np.sum(VW**2)
               - su2*(np.sum(Bt**2) - np.sum(P**2)))
    for h, (duu, dxu) in enumerate(zip(kernel.grad(U), kernel.grad(X, U)), 1):
        dlZ[h] = 0.5*np.sum(duu*C) + 0.5/ell*np.sum(dxu*Cxu)
    dlZ[-1] = np.sum(alpha)/ell
    return lZ, dlZ
