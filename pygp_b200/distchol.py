"""
1-D block-column distributed Cholesky for the largest N (SURVEY.md 8e, config
C5): the factorisation of `ExactGP._update` (pygp/inference/exact.py:50-55)
spread over the GPUs of one node, one process per GPU.

Layout.  Block column j (width nb) of the lower triangle -- rows j nb .. N,
plus the residual row r = y - mean riding along as row N -- is a dense
((N - j nb) + 1, nb) panel owned by rank j mod G ("1-D block-cyclic").

Schedule (right-looking, one-panel lookahead):

    for k in 0 .. nblk-1:
        owner(k) has panel k up to date:  potrf of its top nb x nb block + the
            right-solve of the rows below (pgp_dev_potrf with `extra` rows)
        broadcast panel k over NVLink (NCCL), asynchronously
        every rank copies panel k into its replica of the full factor F
        every rank updates the panels it owns:  P_j -= P_k[rows >= j nb] P_k[block j]^T
            (pgp_dev_gemm_nt; DMMA) -- the owner of k+1 does panel k+1 FIRST,
            factors it and starts its broadcast before touching the rest, so the
            next panel travels while this step's GEMMs run.

The only collective on the data path is that panel broadcast (total ~4 N^2
bytes per factorisation); lZ needs nothing further because every rank ends up
with the complete factor: the broadcasts double as the all-gather that the
sharded predict (replicated L, test points split across ranks) needs anyway.

PRODUCT PATH: `distributed_update` / `distributed_loglikelihood` call
`pgp_dist_exact_update` / `pgp_dist_exact_loglike` (pygp_b200/csrc/dist.cu): the
whole schedule -- panel Gram, potrf, pack, `ncclBroadcast` on a communication
stream, DMMA updates, and for the gradient two "staircase" triangular solves, a
trace and one `ncclAllReduce` -- is enqueued from C++ with events, the host never
waits inside the loop.  (Round 1 paced it from Python, one ctypes call and one
torch.distributed broadcast per panel.)  torch.distributed is used ONCE, to hand
NCCL's unique id from rank 0 to the other ranks.

`distributed_factor` below is the schedule's MODEL: the same ownership, lookahead
and broadcast order behind a small backend interface, exercised on CPU with gloo
and a numpy backend (tests/test_distchol.py); `gradient_partition_model` is the
numpy model of the block-column gradient (tests/test_distchol.py pins both to
dense LAPACK results).
"""

import numpy as np

from . import sharding

__all__ = ['block_columns', 'distributed_factor', 'gradient_partition_model', 'staircase_solve_model', 'Communicator',
           'communicator', 'distributed_update', 'distributed_loglikelihood']


def block_columns(n, nb):
    """[(j0, width)] of the block columns of an n x n matrix."""
    return [(j0, min(nb, n - j0)) for j0 in range(0, n, nb)]


def distributed_factor(be, n, nb, rank, size, bcast, group=1, allreduce_min=None):
    """Run the schedule.  `be` provides the arithmetic on this rank's buffers:

        be.build_panel(j0, w)          -> panel ((n - j0) + 1, w): K[j0:, j0:j0+w] + noise, last row r[j0:j0+w]
        be.recv_buffer(rows, w, slot)  -> buffer (rows, .) a broadcast panel is received into (two slots)
        be.factor_panel(panel, w)      -> LAPACK-style info (0 ok): potrf of the top w x w block in place,
                                          rows below <- rows L^-T
        be.store_panel(pk, j0, w)      copy a factored panel into the replicated factor F (rows j0.., cols j0..j0+w)
        be.update_from_factor(pj, j0, w, c_lo, c_hi)
                                       pj -= F[j0:, c_lo:c_hi] @ F[j0:j0+w, c_lo:c_hi].T   (stored panels only)
        be.sync()                      make the panel's producer visible before it is broadcast

    `bcast(buf, src) -> handle` starts the (asynchronous) broadcast of `buf` from
    rank `src`; `handle.wait()` completes it (None for world size 1).

    `group` > 1: trailing panels are brought up to date with `group` factored panels
    at a time (one GEMM with K = group * nb read from the replicated factor; the
    panel after next is always kept fully up to date).  Measured on 8 GPUs at
    N = 65536, nb = 512 this is SLOWER (0.58 s for 2, 0.60 s for 4, against 0.51 s
    for 1): the per-rank GEMM work is already hidden behind the panel chain
    (factor + broadcast), and deferring it only makes it arrive in bursts.
    `allreduce_min(int) -> int` makes the potrf `info` GLOBAL: it is only known on the
    rank that owns the failing panel (receivers see a NaN-filled panel, not an info), and
    ranks that disagree about success diverge in the caller's error handling and deadlock
    in the next collective.  Returns info (> 0: order of the first non positive-definite
    leading minor), the same value on every rank."""
    cols = block_columns(n, nb)
    nblk = len(cols)
    mine = {j: be.build_panel(*cols[j]) for j in range(nblk) if j % size == rank}
    applied = dict((j, 0) for j in mine)      # panels 0 .. applied[j] - 1 are already applied to panel j
    info = 0

    def catch_up(j, upto):
        if applied[j] < upto:
            c_lo = cols[applied[j]][0]
            c_hi = cols[upto - 1][0] + cols[upto - 1][1]
            be.update_from_factor(mine[j], cols[j][0], cols[j][1], c_lo, c_hi)
            applied[j] = upto

    def produce(k):
        """owner: factor panel k and start its broadcast; others: post the receive."""
        j0, w = cols[k]
        owner = k % size
        if owner == rank:
            buf = mine[k]
            bad = be.factor_panel(buf, w)
            be.sync()
        else:
            buf, bad = be.recv_buffer(n - j0 + 1, w, k % 2), 0
        return buf, (bcast(buf, owner) if size > 1 else None), (j0 + bad if bad else 0)

    cur, handle, bad = produce(0)
    info = info or bad
    for k in range(nblk):
        j0, w = cols[k]
        if handle is not None:
            handle.wait()
        be.store_panel(cur, j0, w)
        nxt = None
        # lookahead: the next panel first, then its factorisation and broadcast
        if k + 1 < nblk:
            if (k + 1) % size == rank:
                catch_up(k + 1, k + 1)
            nxt = produce(k + 1)
            info = info or nxt[2]
        for j in range(k + 2, nblk):
            if j % size == rank and (j == k + 2 or k + 1 - applied[j] >= group):
                catch_up(j, k + 1)
        if nxt is not None:
            cur, handle = nxt[0], nxt[1]
    be.sync()
    if allreduce_min is not None and size > 1:
        big = 1 << 62
        info = allreduce_min(info if info else big)
        info = 0 if info == big else info
    return info


def gradient_partition_model(L, a, dK, sn2, nb, rank, size):
    """numpy MODEL of one rank's share of `ExactGP.loglikelihood(True)`'s gradient
    (exact.py:128-141) under the block-column partition of dist.cu:

        for every owned block column J:  W = L^-1 E_J (rows >= J nb),  H_J = L^-T W (rows >= J nb)
        alpha = L^-T a
        S_0  = sum_{c in J} Q_cc,   S_h = sum_{c in J} (Q_cc dK_h,cc + 2 sum_{i > c} Q_ic dK_h,ic)
        with Q = K~^-1 - alpha alpha^T.

    Summed over the ranks, dlZ = [-sn2 S_0, -1/2 S_h ..., sum(alpha)].  `dK` is the list of dense
    dK_h matrices (test sizes only).  Returns the vector (S_0, S_1, ...)."""
    import scipy.linalg as sla
    n = len(L)
    alpha = sla.solve_triangular(L, a, lower=True, trans=1)
    S = np.zeros(1 + len(dK))
    for j, (j0, w) in enumerate(block_columns(n, nb)):
        if j % size != rank:
            continue
        Lt = L[j0:, j0:]
        E = np.zeros((n - j0, w))
        E[:w] = np.eye(w)
        W = sla.solve_triangular(Lt, E, lower=True)               # (L^-1 E_J)[j0:]
        H = sla.solve_triangular(Lt, W, lower=True, trans=1)      # K~^-1[j0:, J]
        Q = H - np.outer(alpha[j0:], alpha[j0:j0 + w])
        wgt = 2.0*np.tril(np.ones((n - j0, w)), -1) + np.eye(n - j0, w)
        S[0] += np.trace(Q[:w])
        for h, dk in enumerate(dK):
            S[1 + h] += np.sum(wgt*Q*dk[j0:, j0:j0 + w])
    return S


def staircase_solve_model(L, nb, rank, size, a=None, leaf=64):
    """numpy MODEL of the two staircase right-solves of `pgp_dist_exact_loglike` (csrc/chol.cu:
    trsm_rec_stair / trsm_nt_rec_stair with csrc/chol.cuh: Stair), same recursion, same row counts.

    The rows of B are (optionally) one dense front row `a` followed by this rank's rows of the identity:
    block q belongs to global block column J = rank + q size and starts at column J nb.  Returns
    (B, flops): after  B <- B L^-T  on the identity rows and  B <- B L^-1  on all rows, row r of block J
    holds K~^-1[i, c(r)] at the columns i >= J nb, and the front row holds alpha^T = a^T L^-1.
    `flops` counts the multiply-adds x 2 the recursion performs with every GEMM restricted to the rows
    (and, for the first solve, the contraction range) the staircase allows."""
    import scipy.linalg as sla
    n = len(L)
    blocks = [j for j in range(-(-n // nb)) if j % size == rank]
    starts = [j*nb for j in blocks for _ in range(min(nb, n - j*nb))]
    front = 0 if a is None else 1
    B = np.zeros((front + len(starts), n))
    if front:
        B[0] = a
    for r, j in enumerate([j*nb + i for j in blocks for i in range(min(nb, n - j*nb))]):
        B[front + r, j] = 1.0
    flops = [0.0]

    def rows(c_end, fr, total):
        nblk = -(-c_end // nb) if c_end > 0 else 0
        mine = -(-(nblk - rank) // size) if nblk > rank else 0
        return min(fr + mine*nb, total)

    def split(m):
        h = -(-(m // 2) // leaf)*leaf
        h = leaf if h <= 0 else h
        return m - leaf if h >= m else h

    def lt(Bv, fr, j0, m):                      # Bv[:, j0:j0+m] <- Bv[:, j0:j0+m] L[j0.., j0..]^-T with earlier columns eliminated
        total = len(Bv)
        if m <= leaf:
            r = rows(j0 + m, fr, total)
            if r:
                T = L[j0:j0 + m, j0:j0 + m]
                Bv[:r, j0:j0 + m] = sla.solve_triangular(T, Bv[:r, j0:j0 + m].T, lower=True).T
            return
        m1 = split(m)
        c0 = j0 + m1
        lt(Bv, fr, j0, m1)
        r = rows(c0, fr, total)
        if r:
            Bv[:r, c0:j0 + m] -= Bv[:r, j0:c0] @ L[c0:j0 + m, j0:c0].T
            # contraction of row block q starts at its first column (GemmArgs::stair bit 1)
            st = np.array([0]*fr + starts)[:r]
            flops[0] += 2.0*np.sum(np.maximum(c0 - np.maximum(st, j0), 0))*(m - m1)
        lt(Bv, fr, c0, m - m1)

    def ln(Bv, fr, j0, m):                      # Bv[:, j0:j0+m] <- (. L^-1)[:, j0:j0+m], later columns first
        total = len(Bv)
        if m <= leaf:
            r = rows(j0 + m, fr, total)
            if r:
                T = L[j0:j0 + m, j0:j0 + m]
                Bv[:r, j0:j0 + m] = sla.solve_triangular(T, Bv[:r, j0:j0 + m].T, lower=True, trans=1).T
            return
        m1 = split(m)
        c0 = j0 + m1
        ln(Bv, fr, c0, m - m1)
        r = rows(c0, fr, total)
        if r:
            Bv[:r, j0:c0] -= Bv[:r, c0:j0 + m] @ L[c0:j0 + m, j0:c0]
            st = np.array([0]*fr + starts)[:r]
            # output columns left of a row block's first column are skipped in 64-wide tiles (stair bit 2)
            flops[0] += 2.0*np.sum(np.maximum(c0 - np.maximum(st, j0), 0))*(m - m1)
        ln(Bv, fr, j0, m1)

    lt(B[front:], 0, 0, n)
    ln(B, front, 0, n)
    return B, flops[0]


class Communicator(object):
    """Owner of a `pgp_dist*` (NCCL communicator + communication stream + staging buffers)."""

    def __init__(self, ctx, handle, rank, size):
        self.ctx, self.handle, self.rank, self.size = ctx, handle, rank, size

    def allreduce(self, x, op='sum'):
        from . import _lib
        x = _lib.as_f64(x).copy()
        _lib.check(self.ctx, _lib.lib().pgp_dist_allreduce(self.handle, _lib.ptr(x), x.size, {'sum': 0, 'max': 1, 'min': 2}[op]))
        return x

    def __del__(self):
        try:
            if self.handle:
                from . import _lib
                _lib.lib().pgp_dist_destroy(self.handle)
                self.handle = None
        except Exception:       # interpreter shutdown
            pass


_comms = {}


def communicator(group=None):
    """The library's communicator over the ranks of `group` (default: the world), created
    on first use.  torch.distributed only carries NCCL's unique id from rank 0 to the others."""
    import ctypes as C
    from . import _lib
    key = id(group) if group is not None else None
    if key in _comms:
        return _comms[key]
    rank, size = sharding.world(group)
    ctx, L = _lib.context(), _lib.lib()
    uid = np.zeros(128, dtype=np.uint8)
    if size > 1:
        import torch
        import torch.distributed as dist
        if rank == 0:
            _lib.check(ctx, L.pgp_dist_unique_id(ctx.handle, uid.ctypes.data_as(C.c_void_p)))
        t = torch.from_numpy(uid).to(sharding._device_for(group))
        src = dist.get_global_rank(group, 0) if group is not None else 0
        dist.broadcast(t, src=src, group=group)
        uid = t.cpu().numpy()
    h = C.c_void_p()
    _lib.check(ctx, L.pgp_dist_init(ctx.handle, size, rank, uid.ctypes.data_as(C.c_void_p), C.byref(h)))
    _comms[key] = Communicator(ctx, h, rank, size)
    return _comms[key]


def distributed_update(gp, nb=512, group=None, group_size=0, chunks=0):
    """`ExactGP._update` with the factorisation spread over the ranks of `group`.
    Every rank must hold the same model (data and hypers) and call this together.
    Afterwards the model on every rank is factored exactly as after `pgp_exact_update`;
    a matrix that is not positive definite raises the same `LinAlgError` on every rank."""
    from . import _lib
    if gp._dev is None:
        raise RuntimeError('distributed_update needs a model with data (add_data first)')
    comm = communicator(group)
    hyp = _lib.as_f64(gp.get_hyper())
    _lib.check(comm.ctx, _lib.lib().pgp_dist_set_group(comm.handle, int(group_size)))   # 0: the library's default
    _lib.check(comm.ctx, _lib.lib().pgp_dist_set_chunks(comm.handle, int(chunks)))
    _lib.check(comm.ctx, _lib.lib().pgp_dist_exact_update(comm.handle, gp._dev.handle, _lib.ptr(hyp), int(nb)))
    return gp


def distributed_loglikelihood(gp, grad=False, nb=512, group=None):
    """`ExactGP.loglikelihood(grad)` after `distributed_update`, the gradient partitioned by
    block column across the ranks (collective when grad=True; identical result on every rank)."""
    import ctypes as C
    from . import _lib
    comm = communicator(group)
    lZ = C.c_double()
    dlZ = np.empty(gp.nhyper) if grad else None
    _lib.check(comm.ctx, _lib.lib().pgp_dist_exact_loglike(
        comm.handle, gp._dev.handle, int(nb), int(bool(grad)), C.byref(lZ), None if dlZ is None else _lib.ptr(dlZ)))
    return (lZ.value, dlZ) if grad else lZ.value
