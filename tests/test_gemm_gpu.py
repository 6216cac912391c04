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


@pytest.mark.parametrize('n,extra', [(1, 0), (63, 1), (64, 1), (65, 2), (128, 1), (200, 0), (513, 3), (1000, 1)])
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
