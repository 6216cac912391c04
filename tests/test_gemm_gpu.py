"""The DMMA GEMM and the blocked factorisation as building blocks, on device
buffers held by torch (torch is only the allocator here), against float64
numpy results on the same inputs."""

import numpy as np
import numpy.testing as nt
import pytest

pytestmark = pytest.mark.gpu


def _ctx():
    from pygp_b200 import _lib
    return _lib, _lib.context(), _lib.lib()


@pytest.mark.parametrize('m,n,k', [(128, 128, 16), (128, 128, 64), (256, 384, 512), (130, 67, 33),
                                   (1, 200, 64), (300, 1, 1000), (257, 255, 130), (64, 64, 2)])
@pytest.mark.parametrize('alpha,beta', [(1.0, 0.0), (-1.0, 1.0)])
def test_gemm_nt(m, n, k, alpha, beta):
    import torch
    _lib, ctx, L = _ctx()
    rng = np.random.RandomState(m*7 + n*3 + k)
    lda, ldb, ldc = k + (k % 2) + 2, k + (k % 2), n + (n % 2) + 4
    A = np.zeros((m, lda)); A[:, :k] = rng.randn(m, k)
    B = np.zeros((n, ldb)); B[:, :k] = rng.randn(n, k)
    C0 = rng.randn(m, ldc)
    dA, dB, dC = (torch.tensor(x, device='cuda') for x in (A, B, C0))
    torch.cuda.synchronize()
    _lib.check(ctx, L.pgp_dev_gemm_nt(ctx.handle, m, n, k, alpha, dA.data_ptr(), lda, dB.data_ptr(), ldb,
                                      beta, dC.data_ptr(), ldc, 0))
    ctx.sync()
    C = dC.cpu().numpy()
    ref = C0.copy()
    ref[:, :n] = beta*C0[:, :n] + alpha*(A[:, :k] @ B[:, :k].T)
    nt.assert_allclose(C, ref, rtol=1e-12, atol=1e-12*np.sqrt(k))     # padding columns untouched too


def test_gemm_tri_skips_upper_tiles():
    import torch
    _lib, ctx, L = _ctx()
    rng = np.random.RandomState(0)
    n, k = 400, 96
    A = rng.randn(n, k)
    C0 = rng.randn(n, n)
    dA, dC = torch.tensor(A, device='cuda'), torch.tensor(C0, device='cuda')
    torch.cuda.synchronize()
    _lib.check(ctx, L.pgp_dev_gemm_nt(ctx.handle, n, n, k, -1.0, dA.data_ptr(), k, dA.data_ptr(), k,
                                      1.0, dC.data_ptr(), n, 1))
    ctx.sync()
    C = dC.cpu().numpy()
    ref = C0 - A @ A.T
    low = np.tril_indices(n)
    nt.assert_allclose(C[low], ref[low], rtol=1e-12, atol=1e-11)
    # tiles strictly above the diagonal (128-wide) were not touched
    nt.assert_array_equal(C[:128, 128:], C0[:128, 128:])
    nt.assert_array_equal(C[128:256, 256:], C0[128:256, 256:])


@pytest.mark.parametrize('n,extra', [(1, 0), (63, 1), (64, 1), (65, 2), (128, 1), (200, 0), (513, 3), (1000, 1),
                                     (1536, 1), (2100, 2), (3333, 1)])       # >= 1536: lookahead path
def test_potrf_with_extra_rows(n, extra):
    import torch
    _lib, ctx, L = _ctx()
    rng = np.random.RandomState(n)
    G = rng.randn(n, n + 8)
    K = G @ G.T / (n + 8) + 0.5*np.eye(n)
    Rhs = rng.randn(extra, n)
    ld = (n + 15)//16*16
    F = np.full((n + extra, ld), np.nan)
    F[:n, :n] = np.tril(K) + np.triu(np.full((n, n), 7.0), 1)    # junk above the diagonal is ignored
    F[n:, :n] = Rhs
    dF = torch.tensor(F, device='cuda')
    torch.cuda.synchronize()
    info = L.pgp_dev_potrf(ctx.handle, dF.data_ptr(), n, ld, extra)
    assert info == 0
    out = dF.cpu().numpy()
    Lref = np.linalg.cholesky(K)
    nt.assert_allclose(np.tril(out[:n, :n]), Lref, rtol=1e-11, atol=1e-12)
    if extra:
        import scipy.linalg as sla
        nt.assert_allclose(out[n:, :n], sla.solve_triangular(Lref, Rhs.T, lower=True).T, rtol=1e-9, atol=1e-10)


def test_potrf_reports_failing_minor():
    import torch
    _lib, ctx, L = _ctx()
    n = 150
    K = np.eye(n)
    K[100, 100] = -1.0
    dF = torch.tensor(K, device='cuda')
    torch.cuda.synchronize()
    assert L.pgp_dev_potrf(ctx.handle, dF.data_ptr(), n, n, 0) == 101      # LAPACK-style info


@pytest.mark.parametrize('tA,tB', [(0, 1), (1, 1), (1, 0)])
@pytest.mark.parametrize('m,n,k,splitk', [(128, 128, 64, 1), (130, 67, 333, 1), (33, 33, 5000, 0), (65, 96, 9000, 7),
                                          (1, 200, 64, 1), (257, 1, 1000, 3), (8, 8, 40, 1), (2048, 40, 40, 1)])
def test_gemm_general_forms(tA, tB, m, n, k, splitk):
    """TN / NN / TT-style operand layouts and the split contraction (FITC)."""
    import torch
    _lib, ctx, L = _ctx()
    rng = np.random.RandomState(m + 3*n + 7*k + tA + 2*tB)
    ev = lambda x: x + (x % 2)
    A = rng.randn(m, k)
    B = rng.randn(n, k)
    As = np.zeros((k, ev(m) + 2)) if tA else np.zeros((m, ev(k) + 2))
    Bs = np.zeros((k, ev(n))) if tB else np.zeros((n, ev(k)))
    if tA: As[:, :m] = A.T
    else: As[:, :k] = A
    if tB: Bs[:, :n] = B.T
    else: Bs[:, :k] = B
    ldc = ev(n) + 4
    C0 = rng.randn(m, ldc)
    dA, dB, dC = (torch.tensor(x, device='cuda') for x in (As, Bs, C0))
    torch.cuda.synchronize()
    _lib.check(ctx, L.pgp_dev_gemm(ctx.handle, tA, tB, m, n, k, -0.5, dA.data_ptr(), As.shape[1], dB.data_ptr(),
                                   Bs.shape[1], 1.0, dC.data_ptr(), ldc, 0, splitk))
    ctx.sync()
    C = dC.cpu().numpy()
    ref = C0.copy()
    ref[:, :n] = C0[:, :n] - 0.5*(A @ B.T)
    nt.assert_allclose(C, ref, rtol=1e-12, atol=1e-12*np.sqrt(k))


def test_gemm_tn_tri_splitk():
    """A = I + V^T V on the lower tiles with the contraction split (fitc.cu)."""
    import torch
    _lib, ctx, L = _ctx()
    rng = np.random.RandomState(5)
    n, p = 20000, 200
    V = rng.randn(n, p)/np.sqrt(n)
    C0 = np.eye(p)
    dV, dC = torch.tensor(V, device='cuda'), torch.tensor(C0, device='cuda')
    torch.cuda.synchronize()
    _lib.check(ctx, L.pgp_dev_gemm(ctx.handle, 1, 1, p, p, n, 1.0, dV.data_ptr(), p, dV.data_ptr(), p, 1.0,
                                   dC.data_ptr(), p, 1, 0))
    ctx.sync()
    C = dC.cpu().numpy()
    ref = np.eye(p) + V.T @ V
    low = np.tril_indices(p)
    nt.assert_allclose(C[low], ref[low], rtol=1e-12, atol=1e-13)
    up = np.triu_indices(p, 1)
    nt.assert_array_equal(C[up], 0.0)


@pytest.mark.parametrize('n,rows', [(1, 3), (64, 1), (65, 130), (200, 257), (513, 1), (1000, 300)])
@pytest.mark.parametrize('notrans', [0, 1])
def test_trsm_forms(n, rows, notrans):
    import torch
    import scipy.linalg as sla
    _lib, ctx, L = _ctx()
    rng = np.random.RandomState(n + rows)
    T = np.tril(rng.randn(n, n))/np.sqrt(n) + 2*np.eye(n)
    B = rng.randn(rows, n)
    ld = n + (n % 2) + 2
    Tb = rng.randn(n, ld)            # garbage above the diagonal must be ignored
    Tb[:, :n] = np.where(np.tril(np.ones((n, n))) > 0, T, Tb[:, :n])
    Bb = np.zeros((rows, ld)); Bb[:, :n] = B
    dT, dB = torch.tensor(Tb, device='cuda'), torch.tensor(Bb, device='cuda')
    torch.cuda.synchronize()
    _lib.check(ctx, L.pgp_dev_trsm(ctx.handle, dB.data_ptr(), rows, ld, dT.data_ptr(), n, ld, notrans))
    ctx.sync()
    X = dB.cpu().numpy()[:, :n]
    # notrans: X T = B  ->  X = (T^-T B^T)^T ; else X T^T = B -> X = (T^-1 B^T)^T
    ref = sla.solve_triangular(T, B.T, lower=True, trans=1 if notrans else 0).T
    nt.assert_allclose(X, ref, rtol=1e-10, atol=1e-10)


@pytest.mark.parametrize('tA,tB', [(0, 0), (0, 1), (1, 1), (1, 0)])
@pytest.mark.parametrize('m,n,k', [(2048, 1920, 300), (1601, 1537, 67), (4000, 700, 129)])
def test_gemm_large_tile_forms(tA, tB, m, n, k):
    """Shapes with more 128 x 128 tiles than SMs, so the large-tile kernel runs
    (the small shapes above all take the 64 x 64 variant): interior fast loop,
    ragged edges and a k tail, in every operand layout."""
    import torch
    _lib, ctx, L = _ctx()
    rng = np.random.RandomState(m + n + k + tA + 2*tB)
    ev = lambda x: x + (x % 2)
    A, B = rng.randn(m, k), rng.randn(n, k)
    As = np.zeros((k, ev(m))) if tA else np.zeros((m, ev(k) + 2))
    Bs = np.zeros((k, ev(n) + 2)) if tB else np.zeros((n, ev(k)))
    if tA: As[:, :m] = A.T
    else: As[:, :k] = A
    if tB: Bs[:, :n] = B.T
    else: Bs[:, :k] = B
    ldc = ev(n)
    C0 = rng.randn(m, ldc)
    dA, dB, dC = (torch.tensor(x, device='cuda') for x in (As, Bs, C0))
    torch.cuda.synchronize()
    _lib.check(ctx, L.pgp_dev_gemm(ctx.handle, tA, tB, m, n, k, 1.5, dA.data_ptr(), As.shape[1], dB.data_ptr(),
                                   Bs.shape[1], -1.0, dC.data_ptr(), ldc, 0, 1))
    ctx.sync()
    ref = C0.copy()
    ref[:, :n] = -C0[:, :n] + 1.5*(A @ B.T)
    nt.assert_allclose(dC.cpu().numpy(), ref, rtol=1e-12, atol=1e-12*np.sqrt(k))


def test_potrf_reports_failing_minor_lookahead():
    """same on the two-stream lookahead path (n >= 1536), failing inside a later panel"""
    import torch
    _lib, ctx, L = _ctx()
    n = 2000
    K = np.eye(n)
    K[1700, 1700] = -1.0
    dF = torch.tensor(K, device='cuda')
    torch.cuda.synchronize()
    assert L.pgp_dev_potrf(ctx.handle, dF.data_ptr(), n, n, 0) == 1701
