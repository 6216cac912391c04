"""
TEST INFRASTRUCTURE ONLY -- a numpy model of the DEVICE algorithm (not of the
reference): the same recursion, the same block sizes and the same operand
views that pygp_b200/csrc/chol.cu launches, with every kernel replaced by a
numpy expression.  It exists to (a) show that the device formulation reaches
the parity tolerances before any CUDA is written and (b) give stage-by-stage
intermediates to diff against device dumps when debugging.

Device formulation:
  * F is a row-major (N+1) x ld buffer; lower triangle of rows 0..N-1 holds
    K + sn2 I, row N holds r = y - mean.  After chol_rec the lower triangle is
    L (K~ = L L^T, L = R^T of the reference) and row N is a = L^-1 r.
  * chol_rec(j0, n): factor columns [j0, j0+n) for ALL rows below.
  * trsm_rec(B, j0, n): B[:, j0:j0+n] <- B L^-T restricted to those columns.
  * V = L^-T (upper) bottom-up: diagonal blocks, then pairs of blocks of
    doubling size (V12 = -(V11 L21^T) V22); K~^-1 = V V^T.
"""

import numpy as np

NB = 64          # base block (potrf_base / trsm_base)


def potrf_base(F, j0, n):
    A = F[j0:j0+n, j0:j0+n]
    for k in range(n):                       # right-looking, unblocked
        d = A[k, k]
        if not d > 0:
            return j0 + k + 1                # LAPACK-style info
        d = np.sqrt(d)
        A[k, k] = d
        A[k+1:, k] /= d
        for j in range(k+1, n):
            A[j:, j] -= A[j:, k] * A[j, k]
    return 0


def trsm_base(B, L, j0, n):
    """B[:, j0:j0+n] <- B[:, j0:j0+n] L[j0:j0+n, j0:j0+n]^-T by substitution."""
    T = L[j0:j0+n, j0:j0+n]
    X = B[:, j0:j0+n]
    for k in range(n):
        X[:, k] /= T[k, k]
        if k + 1 < n:
            X[:, k+1:] -= np.outer(X[:, k], T[k+1:, k])


def split(n):
    """Split point: a multiple of NB, roughly half."""
    h = (n // 2 + NB - 1) // NB * NB
    return h if h < n else n - NB if n > NB else n


def chol_rec(F, j0, n, M):
    """Factor columns [j0, j0+n) of the M-row buffer F (rows j0..M-1)."""
    if n <= NB:
        info = potrf_base(F, j0, n)
        if info:
            return info
        trsm_base(F[j0+n:M], F, j0, n)
        return 0
    n1 = split(n)
    info = chol_rec(F, j0, n1, M)
    if info:
        return info
    c0 = j0 + n1
    # trapezoid update: rows c0..M-1, cols c0..j0+n-1, tiles strictly above the
    # diagonal skipped (modelled with tril on the square part)
    P = F[c0:M, j0:c0]
    U = P @ F[c0:j0+n, j0:c0].T
    n2 = n - n1
    U[:n2, :n2] = np.tril(U[:n2, :n2])
    F[c0:M, c0:j0+n] -= U
    return chol_rec(F, c0, n2, M)


def trsm_rec(B, L, j0, n):
    if n <= NB:
        trsm_base(B, L, j0, n)
        return
    n1 = split(n)
    trsm_rec(B, L, j0, n1)
    B[:, j0+n1:j0+n] -= B[:, j0:j0+n1] @ L[j0+n1:j0+n, j0:j0+n1].T
    trsm_rec(B, L, j0+n1, n - n1)


def inv_upper(G, L, n):
    """G[0:n, 0:n] <- L^-T (upper), G pre-zeroed: bottom-up as chol.cu does.
    Diagonal 64-blocks by substitution on the identity, then for block sizes
    b = 64, 128, ... every pair of neighbouring blocks:
        V12 = -(V11 L21^T) V22     (V11, V22 upper triangular)"""
    for j0 in range(0, n, NB):
        m = min(NB, n - j0)
        G[j0:j0+m, j0:j0+m] = np.eye(m)
        trsm_base(G[j0:j0+m], L, j0, m)
    b = NB
    while b < n:
        for o in range(0, n, 2*b):
            n2 = min(n, o + 2*b) - (o + b)
            if n2 <= 0:
                continue
            T = -(G[o:o+b, o:o+b] @ L[o+b:o+b+n2, o:o+b].T)
            G[o:o+b, o+b:o+b+n2] = T @ G[o+b:o+b+n2, o+b:o+b+n2]
        b *= 2


def device_model(kernel, sn, mean, X, y, want_grad=True, Xs=None):
    """Run the device formulation in numpy; returns dict of results."""
    N = len(X)
    sn2 = np.exp(2*np.log(sn))
    F = np.zeros((N+1, N))
    F[:N] = np.tril(kernel.get(X) + sn2*np.eye(N))
    F[N] = y - mean
    info = chol_rec(F, 0, N, N+1)
    out = {'info': info, 'F': F}
    if info:
        return out
    L, a = F[:N], F[N]
    lZ = -0.5*np.dot(a, a) - 0.5*N*np.log(2*np.pi) - np.sum(np.log(np.diag(L)))
    out['lZ'] = lZ
    if want_grad:
        G = np.zeros((N, N))
        inv_upper(G, L, N)
        alpha = G @ a                         # alpha = L^-T a
        Pm = np.tril(G @ G.T)                 # K~^-1, lower (lauum: k >= max(i,j))
        Q = Pm - np.tril(np.outer(alpha, alpha))
        w = 2.0*np.ones((N, N))
        w[np.diag_indices(N)] = 1.0           # symmetric sum over the lower part
        g = [-sn2*np.trace(Q)]
        g += [-0.5*np.sum(w*Q*np.tril(dK)) for dK in kernel.grad(X)]
        g += [np.sum(alpha)]
        out['dlZ'] = np.array(g)
    if Xs is not None:
        B = kernel.get(Xs, X).copy()          # (m, N): rows = test points
        trsm_rec(B, L, 0, N)
        out['mu'] = mean + B @ a
        out['s2'] = kernel.dget(Xs) - np.sum(B*B, axis=1)
    return out
