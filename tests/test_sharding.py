"""Host-side multi-GPU logic (pygp_b200/sharding.py) on CPU: world_size-2 gloo
process groups, with the oracle standing in for the per-rank device call (the
sharding / gather / moment logic is what is under test here; the device call
itself is covered by the -m gpu tests)."""

import os
import socket

import numpy as np
import numpy.testing as nt
import pytest


def test_shard_range_partitions():
    from pygp_b200.sharding import shard_range
    for total in (0, 1, 7, 8, 16, 1000003):
        for size in (1, 2, 3, 8):
            r = [shard_range(total, k, size) for k in range(size)]
            assert r[0][0] == 0 and r[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1


def test_world_size_one_needs_no_process_group():
    from pygp_b200 import sharding
    from oracle.pygp_oracle import make_kernel, OExactGP, synthetic_problem
    X, y, Xs = synthetic_problem(60, 2, 11)
    gp = OExactGP(0.1, make_kernel(('se', 1.0, [0.5, 0.6])), 0.0)
    gp.add_data(X, y)
    mu, s2 = sharding.sharded_posterior(gp, Xs)
    mu0, s20 = gp.posterior(Xs)
    nt.assert_array_equal(mu, mu0)
    nt.assert_array_equal(s2, s20)


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, size, port, q):
    import copy
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=size)
    try:
        from pygp_b200 import sharding
        from oracle.pygp_oracle import make_kernel, OExactGP, synthetic_problem
        X, y, Xs = synthetic_problem(80, 3, 13)          # 13 points over 2 ranks: ragged
        gp = OExactGP(0.1, make_kernel(('matern', 1.0, [0.5, 0.6, 0.7], 5)), 0.1)
        gp.add_data(X, y)
        mu, s2 = sharding.sharded_posterior(gp, Xs)
        mu0, s20 = gp.posterior(Xs)
        nt.assert_allclose(mu, mu0, rtol=1e-13, atol=1e-15)   # BLAS blocking differs with the slice shape
        nt.assert_allclose(s2, s20, rtol=1e-12, atol=1e-15)

        rng = np.random.RandomState(2)
        H = gp.get_hyper() + rng.uniform(-0.5, 0.5, size=(7, gp.nhyper))

        def lz(hs):
            out = []
            for h in hs:
                g = copy.deepcopy(gp)
                g.set_hyper(h)
                out.append(g.loglikelihood())
            return np.array(out)

        def pred(hs, x):
            parts = []
            for h in hs:
                g = copy.deepcopy(gp)
                g.set_hyper(h)
                parts.append(g.posterior(x))
            return np.array([p[0] for p in parts]), np.array([p[1] for p in parts])

        got = sharding.sharded_batched_loglike(gp, H, local_fn=lz)
        nt.assert_array_equal(got, lz(H))
        mu, s2 = sharding.sharded_mixture_posterior(gp, H, Xs, local_fn=pred)
        mu_, s2_ = pred(H, Xs)                              # mcmc.py:84-93 on one rank
        mu0 = np.mean(mu_, axis=0)
        s20 = np.mean(s2_ + (mu_ - mu0)**2, axis=0)
        nt.assert_allclose(mu, mu0, rtol=1e-14)
        nt.assert_allclose(s2, s20, rtol=1e-12)
        # a rank with no work (1 sample, 2 ranks) still takes part in the collectives
        got = sharding.sharded_batched_loglike(gp, H[:1], local_fn=lz)
        nt.assert_array_equal(got, lz(H[:1]))

        # values and gradients in one sharded call (multi-restart objective, optimization.py:54-62)
        def lzg(hs):
            out = []
            for h in hs:
                g = copy.deepcopy(gp)
                g.set_hyper(h)
                l, d = g.loglikelihood(True)
                out.append(np.r_[l, d])
            return np.array(out).reshape(len(hs), 1 + gp.nhyper)
        l2, d2 = sharding.sharded_batched_loglike(gp, H, local_fn=lzg, grad=True)
        ref = lzg(H)
        nt.assert_array_equal(l2, ref[:, 0])
        nt.assert_array_equal(d2, ref[:, 1:])

        # a failure on ONE rank (non-PD hyper slice) must surface on EVERY rank before the
        # all-reduces, as the same exception class, instead of leaving the others waiting
        def pred_bad(hs, x):
            if rank == 1:
                raise np.linalg.LinAlgError('not positive definite (rank 1 only)')
            return pred(hs, x)
        try:
            sharding.sharded_mixture_posterior(gp, H, Xs, local_fn=pred_bad)
            raise AssertionError('expected LinAlgError on every rank')
        except np.linalg.LinAlgError:
            pass
        # ... and the group is still usable afterwards
        got = sharding.sharded_batched_loglike(gp, H, local_fn=lz)
        nt.assert_array_equal(got, lz(H))
        q.put((rank, 'ok'))
    except Exception as e:      # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_gloo_world_size_2():
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
