"""Multi-GPU check + timing (run under torchrun on the GPU box, one rank per GPU):
  torchrun --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py [N] [nb] [kernel]
* block-column distributed Cholesky and block-column gradient (pygp_b200/csrc/dist.cu through
  pygp_b200.distchol) == single-GPU update / loglikelihood(True), and their speed-up
* sharded predict (test points) and sharded batched loglike / mixture posterior == single rank.
kernel: 'se8' (SE-ARD d=8, default) or 'c5' (SE + Periodic d=1 on sorted inputs: BASELINE configs[4]).
Rank 0 prints JSON lines; any mismatch beyond the parity tolerances exits non-zero on every rank."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    nb = int(sys.argv[2]) if len(sys.argv) > 2 else 512
    kern = sys.argv[3] if len(sys.argv) > 3 else 'se8'
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    os.environ['PYGP_B200_DEVICE'] = str(local)
    if os.environ.get('PGP_DIST_PROF_DUMP'):
        os.environ['PGP_PROF_DUMP'] = '%s.rank%d' % (os.environ['PGP_DIST_PROF_DUMP'], rank)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    import pygp_b200 as pygp
    from pygp_b200 import sharding, distchol, _lib
    ctx = _lib.context(local)

    def say(**kw):
        if rank == 0:
            print(json.dumps(kw), flush=True)

    rng = np.random.RandomState(0)
    if kern == 'c5':
        d = 1
        X = np.sort(rng.rand(n, 1), axis=0)*64
        mk = lambda: pygp.inference.ExactGP(pygp.likelihoods.Gaussian(0.1),
                                            pygp.kernels.SE(1.0, 0.5, 1) + pygp.kernels.Periodic(0.5, 1.0, 0.25), 0.0)
    else:
        d = 8
        X = rng.rand(n, d)
        mk = lambda: pygp.inference.ExactGP(pygp.likelihoods.Gaussian(0.1), pygp.kernels.SE(1.0, [0.5*np.sqrt(d)]*d), 0.0)
    y = np.sin(3*X.sum(1)) + 0.1*rng.randn(n)
    Xs = np.random.RandomState(1).rand(4096, d)*(64 if kern == 'c5' else 1)
    if os.environ.get('PGP_DIST_SWEEP'):
        # timing sweep only: "nb:group,nb:group,..." -- distributed evaluation per setting, no one-GPU reference
        g2 = mk()
        g2.add_data(X, y)
        comm = distchol.communicator()
        for item in os.environ['PGP_DIST_SWEEP'].split(','):
            vals = [int(v) for v in item.split(':')]
            nb_s, grp, chk = vals[0], vals[1], (vals[2] if len(vals) > 2 else 0)
            os.environ['PGP_DIST_GROUP_NOW'] = str(grp)
            for rep in range(2):
                dist.barrier()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                distchol.distributed_update(g2, nb=nb_s, group_size=grp, chunks=chk)
                ctx.sync()
                t_u = comm.allreduce([time.perf_counter() - t0], 'max')[0]
                t0 = time.perf_counter()
                lZs, dlZs = distchol.distributed_loglikelihood(g2, True, nb=nb_s)
                t_g = comm.allreduce([time.perf_counter() - t0], 'max')[0]
            say(check='dist_sweep', kernel=kern, n=n, world=world, nb=nb_s, group=grp, chunks=chk, update_s=t_u, grad_s=t_g, lZ=lZs)
        dist.destroy_process_group()
        return
    gp = mk()
    gp.add_data(X, y)                      # single-GPU factorisation on every rank (reference)
    ctx.sync()
    t0 = time.perf_counter()
    gp.set_hyper(gp.get_hyper())
    ctx.sync()
    t_single = time.perf_counter() - t0
    t0 = time.perf_counter()
    lZ0, dlZ0 = gp.loglikelihood(True)
    t_single_grad = time.perf_counter() - t0
    mu0, s20 = gp.posterior(Xs[:256])
    del gp                                 # its gradient buffers (2 N^2) go back to the pool

    g2 = mk()
    g2.add_data(X, y)
    comm = distchol.communicator()
    for rep in range(2):
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        distchol.distributed_update(g2, nb=nb)
        ctx.sync()
        t_dist = comm.allreduce([time.perf_counter() - t0], 'max')[0]
        t0 = time.perf_counter()
        lZ1, dlZ1 = distchol.distributed_loglikelihood(g2, True, nb=nb)
        t_dist_grad = comm.allreduce([time.perf_counter() - t0], 'max')[0]
    if os.environ.get('PGP_DIST_PROF'):
        # per-class device time of one distributed evaluation on this rank (CUDA events per launch)
        names = ['gemm', 'gram', 'trace', 'potrf_base', 'trsm_base', 'other']
        for what in ('update', 'grad'):
            ctx.profile(True)
            t0 = time.perf_counter()
            if what == 'update':
                distchol.distributed_update(g2, nb=nb)
            else:
                distchol.distributed_loglikelihood(g2, True, nb=nb)
            ctx.sync()
            wall = time.perf_counter() - t0
            prof = {nm: ctx.profile_read(i) for i, nm in enumerate(names)}
            ctx.profile(False)
            say(check='dist_profile_rank0', what=what, wall_s=wall,
                classes={nm: {'launches': v[0], 'ms': round(v[1], 2), 'tflops_or_GBps': round(v[2]/max(v[1], 1e-9)/1e9, 2)}
                         for nm, v in prof.items()})
    mu1, s21 = g2.posterior(Xs[:256])
    gscale = float(np.abs(dlZ0).max())
    say(check='distributed_eval', kernel=kern, n=n, nb=nb, world=world, lZ_single=lZ0, lZ_dist=lZ1,
        lZ_rel_err=abs(lZ1 - lZ0)/abs(lZ0), dlZ_rel_err=float(np.abs(dlZ1 - dlZ0).max()/gscale),
        mu_err=float(np.abs(mu1 - mu0).max()), s2_err=float(np.abs(s21 - s20).max()),
        update_single_s=t_single, update_dist_s=t_dist, update_speedup=t_single/t_dist,
        grad_single_s=t_single_grad, grad_dist_s=t_dist_grad, grad_speedup=t_single_grad/t_dist_grad,
        eval_single_s=t_single + t_single_grad, eval_dist_s=t_dist + t_dist_grad,
        eval_speedup=(t_single + t_single_grad)/(t_dist + t_dist_grad),
        update_tflops_single=n**3/3/t_single/1e12, update_tflops_dist_aggregate=n**3/3/t_dist/1e12,
        grad_tflops_dist_aggregate=2*n**3/3/t_dist_grad/1e12)
    assert abs(lZ1 - lZ0) <= 1e-10*abs(lZ0), (lZ0, lZ1)
    assert np.abs(dlZ1 - dlZ0).max() <= 1e-8*gscale, (dlZ0, dlZ1)

    gp = g2
    # sharded predict
    dist.barrier()
    t0 = time.perf_counter()
    mu_s, s2_s = sharding.sharded_posterior(gp, Xs)
    dist.barrier()
    t_sh = time.perf_counter() - t0
    t0 = time.perf_counter()
    mu_f, s2_f = gp.posterior(Xs)
    t_one = time.perf_counter() - t0
    say(check='sharded_predict', points=len(Xs), max_err=float(max(np.abs(mu_s - mu_f).max(), np.abs(s2_s - s2_f).max())),
        t_sharded_s=t_sh, t_single_rank_s=t_one)
    assert np.allclose(mu_s, mu_f, rtol=1e-12, atol=1e-13) and np.allclose(s2_s, s2_f, rtol=1e-12, atol=1e-13)

    # sharded batched hypers (C4 shape, scaled): N=2048, 64 samples per rank
    Xb, yb = X[:2048], y[:2048]
    gb = mk()
    gb.add_data(Xb, yb)
    H = gb.get_hyper() + np.random.RandomState(2).uniform(-0.5, 0.5, size=(64*world, gb.nhyper))
    dist.barrier()
    t0 = time.perf_counter()
    lz_sh = sharding.sharded_batched_loglike(gb, H)
    dist.barrier()
    t_sh = time.perf_counter() - t0
    ref = []
    for h in H[:4]:
        gb.set_hyper(h)
        ref.append(gb.loglikelihood())
    mu_m, s2_m = sharding.sharded_mixture_posterior(gb, H[:8*world], Xs[:64])
    say(check='sharded_batched_loglike', B=len(H), t_s=t_sh, samples_per_s=len(H)/t_sh,
        rel_err_first4=float(np.max(np.abs(lz_sh[:4] - np.array(ref))/np.abs(ref))),
        mixture_finite=bool(np.all(np.isfinite(mu_m)) and np.all(s2_m > 0)))
    assert np.allclose(lz_sh[:4], ref, rtol=1e-10)
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
