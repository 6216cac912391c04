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

The arithmetic is behind a small backend interface so that the schedule can be
exercised on CPU with gloo (tests/test_distchol.py, numpy stand-in); the
product backend is `DeviceBackend` (C ABI on torch-allocated device buffers --
torch is allocator and NCCL plumbing only).
"""

import numpy as np

from . import sharding

__all__ = ['block_columns', 'distributed_factor', 'DeviceBackend', 'distributed_update']


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


class DeviceBackend(object):
    """The arithmetic of `distributed_factor` on this rank's GPU through the C
    ABI; panels are torch-allocated device buffers (row-major, ld = w)."""

    def __init__(self, gp, hyp, nb):
        import ctypes as C
        import torch
        from . import _lib
        self._lib, self._torch, self._C = _lib, torch, C
        self.gp, self.hyp = gp, np.ascontiguousarray(hyp, dtype=np.float64)
        self.ctx, self.L = gp._dev.ctx, _lib.lib()
        self.dev = torch.device('cuda', self.ctx.device)
        self.stream = torch.cuda.ExternalStream(self.ctx.stream, device=self.dev)
        self.n, self.nb = gp.ndata, int(nb)
        if self.nb < 64 or self.nb % 64:
            raise ValueError('block width must be a multiple of 64')
        nk = gp._kernel.nhyper
        self.sn2 = float(np.exp(2*self.hyp[0]))
        self.khyp = _lib.as_f64(self.hyp[1:1 + nk])
        self.mean = float(self.hyp[-1])
        self.spec = gp._kernel._spec()
        with torch.cuda.stream(self.stream):
            self.X = torch.from_numpy(np.ascontiguousarray(gp._X)).to(self.dev)
            self.r = torch.from_numpy(np.ascontiguousarray(gp._y - self.mean)).to(self.dev)
        p, ld = C.c_void_p(), C.c_int64()
        _lib.check(self.ctx, self.L.pgp_exact_factor_buffer(gp._dev.handle, C.byref(p), C.byref(ld)))
        self.F_ptr, self.ld = p.value, ld.value
        self._recv = {}
        self.info = 0

    def build_panel(self, j0, w):
        """((n - j0) + 1, nb) device panel (logical width w <= nb; row pitch nb keeps
        every operand 16-byte aligned with an even leading dimension)."""
        torch = self._torch
        rows, nb = self.n - j0, self.nb
        with torch.cuda.stream(self.stream):
            panel = torch.empty((rows + 1, nb), dtype=torch.float64, device=self.dev)
            tgt = panel if w == nb else torch.empty((rows, w), dtype=torch.float64, device=self.dev)
            self._lib.check(self.ctx, self.L.pgp_gram_dev(
                self.ctx.handle, self.spec, self._lib.ptr(self.khyp), self.X[j0:].data_ptr(), rows,
                self.X[j0:j0 + w].data_ptr(), w, tgt.data_ptr()))
            if w != nb:                                  # ragged last block column
                panel.zero_()
                panel[:rows, :w] = tgt
            panel[:w, :w].diagonal().add_(self.sn2)     # + sn2 I   (exact.py:51-53)
            panel[rows, :w] = self.r[j0:j0 + w]
        return panel

    def recv_buffer(self, rows, w, slot):
        torch = self._torch
        if slot not in self._recv:
            with torch.cuda.stream(self.stream):
                self._recv[slot] = torch.empty((self.n + 1, self.nb), dtype=torch.float64, device=self.dev)
        return self._recv[slot][:rows]

    def factor_panel(self, panel, w):
        rc = self.L.pgp_dev_potrf(self.ctx.handle, panel.data_ptr(), w, self.nb, panel.shape[0] - w)
        if rc < 0:
            self._lib.check(self.ctx, rc)
        return rc

    def update_from_factor(self, pj, j0, wj, c_lo, c_hi):
        # operands straight from the replicated factor: rows j0.. of the stored block columns [c_lo, c_hi)
        a = self.F_ptr + (j0*self.ld + c_lo)*8
        self._lib.check(self.ctx, self.L.pgp_dev_gemm_nt(
            self.ctx.handle, pj.shape[0], wj, c_hi - c_lo, -1.0, a, self.ld, a, self.ld,
            1.0, pj.data_ptr(), self.nb, 0))

    def store_panel(self, pk, j0, w):
        # strided device copy into the model's factor buffer (rows j0 .. n, columns j0 .. j0 + w)
        dst = self.F_ptr + (j0*self.ld + j0)*8
        self._lib.check(self.ctx, self.L.pgp_dev_copy2d(self.ctx.handle, dst, self.ld*8, pk.data_ptr(), self.nb*8,
                                                        w*8, pk.shape[0]))

    def sync(self):
        self.ctx.sync()


def distributed_update(gp, nb=512, group=None, panels_per_update=1):
    """`ExactGP._update` with the factorisation spread over the ranks of `group`.
    Every rank must hold the same model (data and hypers).  Afterwards the model
    on every rank is factored exactly as after `pgp_exact_update`."""
    import torch
    import torch.distributed as dist
    from . import _lib
    rank, size = sharding.world(group)
    hyp = _lib.as_f64(gp.get_hyper())
    if gp._dev is None:
        raise RuntimeError('distributed_update needs a model with data (add_data first)')
    be = DeviceBackend(gp, hyp, nb)

    class _Handle(object):
        def __init__(self, work):
            self.work = work

        def wait(self):
            self.work.wait()

    def bcast(buf, src):
        with torch.cuda.stream(be.stream):
            src_global = dist.get_global_rank(group, src) if group is not None else src
            return _Handle(dist.broadcast(buf, src=src_global, group=group, async_op=True))

    def allreduce_min(v):
        t = torch.tensor([v], dtype=torch.int64, device=be.dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
        return int(t.item())

    with torch.cuda.stream(be.stream):
        info = distributed_factor(be, gp.ndata, nb, rank, size, bcast, panels_per_update, allreduce_min)
    if info:
        raise np.linalg.LinAlgError('%d-th leading minor of the array is not positive definite' % info)
    _lib.check(be.ctx, _lib.lib().pgp_exact_adopt_factor(gp._dev.handle, _lib.ptr(hyp)))
    return gp
