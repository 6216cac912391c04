"""The block-column distributed Cholesky schedule (pygp_b200/distchol.py) on
CPU: world size 1 in-process and world size 2 over gloo, with a numpy stand-in
for the per-rank device arithmetic (the schedule, ownership, lookahead and
broadcast plumbing are under test; DeviceBackend itself is covered by
tests/test_exact_gpu.py::test_distributed_update_single_rank and tools/dist_check.py)."""

import os
import socket

import numpy as np
import numpy.testing as nt
import pytest
import scipy.linalg as sla


class NumpyBackend(object):
    """TEST stand-in for distchol.DeviceBackend: same interface, numpy arithmetic."""

    def __init__(self, K, r, nb):
        self.K, self.r, self.n, self.nb = K, r, len(K), nb
        self.F = np.full((self.n + 1, self.n), np.nan)
        self._recv = {}

    def build_panel(self, j0, w):
        p = np.zeros((self.n - j0 + 1, self.nb))
        p[:-1, :w] = self.K[j0:, j0:j0 + w]
        p[-1, :w] = self.r[j0:j0 + w]
        return p

    def recv_buffer(self, rows, w, slot):
        if slot not in self._recv:
            self._recv[slot] = np.zeros((self.n + 1, self.nb))
        return self._recv[slot][:rows]

    def factor_panel(self, panel, w):
        try:
            L = np.linalg.cholesky(panel[:w, :w])
        except np.linalg.LinAlgError:
            return 1
        panel[:w, :w] = L
        panel[w:, :w] = sla.solve_triangular(L, panel[w:, :w].T, lower=True).T
        return 0

    def update_from_factor(self, pj, j0, wj, c_lo, c_hi):
        A = self.F[j0:, c_lo:c_hi]
        assert not np.isnan(A).any()          # only stored (factored) panels may be read
        pj[:, :wj] -= A @ self.F[j0:j0 + wj, c_lo:c_hi].T

    def store_panel(self, pk, j0, w):
        self.F[j0:, j0:j0 + w] = pk[:, :w]

    def sync(self):
        pass


def _problem(n, seed=0):
    rng = np.random.RandomState(seed)
    A = rng.randn(n, n + 8)
    return A @ A.T/n + np.eye(n), rng.randn(n)


@pytest.mark.parametrize('group', [1, 2, 3])
@pytest.mark.parametrize('n,nb', [(64, 64), (100, 64), (257, 64), (400, 128), (700, 64)])
def test_schedule_single_rank(n, nb, group):
    from pygp_b200.distchol import distributed_factor, block_columns
    assert sum(w for _, w in block_columns(n, nb)) == n
    K, r = _problem(n)
    be = NumpyBackend(K, r, nb)
    assert distributed_factor(be, n, nb, 0, 1, None, group) == 0
    L = np.linalg.cholesky(K)
    low = np.tril_indices(n)
    nt.assert_allclose(be.F[:n][low], L[low], rtol=1e-11, atol=1e-12)
    nt.assert_allclose(be.F[n], sla.solve_triangular(L, r, lower=True), rtol=1e-10, atol=1e-12)


def test_not_positive_definite_info():
    from pygp_b200.distchol import distributed_factor
    K, r = _problem(200)
    K[150, 150] = -1.0
    be = NumpyBackend(K, r, 64)
    assert distributed_factor(be, 200, 64, 0, 1, None) > 128      # reported inside the third block column


@pytest.mark.parametrize('n,nb,size', [(200, 64, 1), (200, 64, 2), (333, 64, 3), (300, 128, 2), (64, 64, 2)])
def test_gradient_partition_model(n, nb, size):
    """The block-column partition of the gradient (dist.cu: two solves per owned block column
    + trace over i >= c, one all-reduce) == exact.py:128-141 computed densely."""
    from pygp_b200.distchol import gradient_partition_model
    K, r = _problem(n, seed=n + 1)
    rng = np.random.RandomState(n)
    dK = []
    for _ in range(3):
        A = rng.randn(n, n)
        dK.append(A + A.T)
    sn2 = 0.3
    L = np.linalg.cholesky(K)
    a = sla.solve_triangular(L, r, lower=True)
    S = sum(gradient_partition_model(L, a, dK, sn2, nb, rank, size) for rank in range(size))
    iK = np.linalg.inv(K)
    alpha = iK @ r
    Q = iK - np.outer(alpha, alpha)
    want = np.r_[np.trace(Q), [np.sum(Q*d) for d in dK]]
    nt.assert_allclose(S, want, rtol=1e-9, atol=1e-9*np.abs(want).max())


@pytest.mark.parametrize('n,nb,size', [(300, 64, 1), (300, 64, 2), (500, 128, 3), (130, 64, 4), (257, 64, 2), (640, 128, 8)])
def test_staircase_solve_model(n, nb, size):
    """The staircase recursions of the distributed gradient (row counts per column range, contraction starts) give,
    on every rank, the owned columns of K~^-1 at and below their diagonal and alpha -- including ranks that own
    nothing (130 / 64 on 4 ranks) and ragged last blocks -- with n^3 / (3 size) flops per solve up to block effects."""
    from pygp_b200.distchol import staircase_solve_model, block_columns
    K, r = _problem(n, seed=3*n)
    L = np.linalg.cholesky(K)
    a = sla.solve_triangular(L, r, lower=True)
    iK = np.linalg.inv(K)
    alpha = iK @ r
    total = 0.0
    for rank in range(size):
        B, fl = staircase_solve_model(L, nb, rank, size, a)
        total += fl
        nt.assert_allclose(B[0], alpha, rtol=1e-9, atol=1e-9*np.abs(alpha).max())
        row = 1
        for j, (j0, w) in enumerate(block_columns(n, nb)):
            if j % size != rank:
                continue
            for i in range(w):
                c = j0 + i
                nt.assert_allclose(B[row, c:], iK[c:, c], rtol=1e-8, atol=1e-9*np.abs(iK).max())
                row += 1
        assert row == len(B)
    # flops of the GEMM part over all ranks: two solves of ~n^3/3 each, never the rows x n^2 of a dense solve
    assert total <= 2*(n**3/3.0)*1.0 + 4.0*n*n*nb


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, size, port, q):
    import torch
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=size)
    try:
        from pygp_b200.distchol import distributed_factor

        class H(object):
            def __init__(self, work, t, buf):
                self.work, self.t, self.buf = work, t, buf

            def wait(self):
                self.work.wait()
                self.buf[...] = self.t.numpy()

        def bcast(buf, src):
            t = torch.from_numpy(np.ascontiguousarray(buf))
            return H(dist.broadcast(t, src=src, async_op=True), t, buf)

        for n, nb, group in [(300, 64, 1), (257, 128, 2), (64, 64, 2), (900, 64, 3)]:
            K, r = _problem(n, seed=n)
            be = NumpyBackend(K, r, nb)
            info = distributed_factor(be, n, nb, rank, size, bcast, group)
            assert info == 0
            L = np.linalg.cholesky(K)
            low = np.tril_indices(n)
            # every rank ends with the complete factor (the broadcasts are the all-gather)
            nt.assert_allclose(be.F[:n][low], L[low], rtol=1e-11, atol=1e-12)
            nt.assert_allclose(be.F[n], sla.solve_triangular(L, r, lower=True), rtol=1e-10, atol=1e-12)

        # not positive definite: `info` is found by the owner of the failing panel only; with the
        # MIN all-reduce every rank reports the same first failing minor (distchol.py:90)
        def allreduce_min(v):
            t = torch.tensor([v], dtype=torch.int64)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            return int(t.item())
        K, r = _problem(300, seed=5)
        K[150, 150] = -1.0                      # block column 2 (nb = 64) -> owned by rank 0 of 2
        with np.errstate(all='ignore'):
            info = distributed_factor(NumpyBackend(K, r, 64), 300, 64, rank, size, bcast, 1, allreduce_min)
        assert 128 < info <= 192, info
        K[150, 150] = K[149, 149]
        K[100, 100] = -1.0                      # block column 1 -> owned by rank 1
        with np.errstate(all='ignore'):
            info = distributed_factor(NumpyBackend(K, r, 64), 300, 64, rank, size, bcast, 1, allreduce_min)
        assert 64 < info <= 128, info
        q.put((rank, 'ok'))
    except Exception:      # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_schedule_gloo_world_size_2():
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=150) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    for rank, msg in res:
        assert msg == 'ok', 'rank %d: %s' % (rank, msg)
