"""Development probe (run on the GPU box): FP64 pipe peaks, cuBLAS DGEMM, our
DMMA GEMM, blocked Cholesky vs cuSOLVER (through torch), and a per-class time
breakdown of one loglike+grad evaluation.  Prints JSON lines."""

import json
import os
import subprocess
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pygp_b200 import _lib  # noqa: E402
import pygp_b200 as pygp     # noqa: E402


def ev_time(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [8192]
    exe = os.path.join(ROOT, 'tools', 'bin', 'fp64_peak')
    if os.path.exists(exe):
        print(subprocess.check_output([exe]).decode().strip(), flush=True)
    ctx, L = _lib.context(), _lib.lib()
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device='cuda')
    b = torch.randn(n, n, dtype=torch.float64, device='cuda')
    c = torch.zeros(n, n, dtype=torch.float64, device='cuda')
    t = ev_time(lambda: torch.matmul(a, b.T, out=c), reps=5)
    print(json.dumps({'cublas_dgemm_tflops': 2*n**3/t/1e12, 'n': n}), flush=True)

    def ours():
        _lib.check(ctx, L.pgp_dev_gemm_nt(ctx.handle, n, n, n, 1.0, a.data_ptr(), n, b.data_ptr(), n, 0.0,
                                          c.data_ptr(), n, 0))
        ctx.sync()
    t = ev_time(ours, reps=5)
    ref = torch.matmul(a, b.T)
    err = float((c - ref).abs().max() / ref.abs().max())
    print(json.dumps({'dmma_gemm_tflops': 2*n**3/t/1e12, 'n': n, 'rel_err_vs_cublas': err}), flush=True)
    del a, b, c, ref

    for n in sizes:
        g = torch.randn(n, n + 64, dtype=torch.float64, device='cuda')
        K = g @ g.T / (n + 64) + torch.eye(n, dtype=torch.float64, device='cuda')
        del g
        F = K.clone()
        t_cus = ev_time(lambda: torch.linalg.cholesky(K), reps=2)

        def potrf():
            F.copy_(K)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            info = L.pgp_dev_potrf(ctx.handle, F.data_ptr(), n, n, 0)
            assert info == 0
            return time.perf_counter() - t0
        potrf()
        t_ours = min(potrf() for _ in range(2))
        Lr = torch.linalg.cholesky(K)
        err = float((torch.tril(F) - Lr).abs().max())
        print(json.dumps({'n': n, 'potrf_ours_s': t_ours, 'potrf_ours_tflops': n**3/3/t_ours/1e12,
                          'potrf_cusolver_s': t_cus, 'potrf_cusolver_tflops': n**3/3/t_cus/1e12,
                          'max_abs_err': err}), flush=True)
        del K, F, Lr
        torch.cuda.empty_cache()

        d = 16
        rng = np.random.RandomState(0)
        X = rng.rand(n, d)
        y = np.sin(3*X.sum(1)) + 0.1*rng.randn(n)
        gp = pygp.inference.ExactGP(pygp.likelihoods.Gaussian(0.1), pygp.kernels.Matern(1.0, [2.0]*d, 5), 0.0)
        t0 = time.perf_counter()
        gp.add_data(X, y)
        gp.loglikelihood(True)
        t_first = time.perf_counter() - t0
        h = gp.get_hyper()
        ctx.profile(True)
        t0 = time.perf_counter()
        gp.set_hyper(h + 0.01)
        t_upd = time.perf_counter() - t0
        lZ, dlZ = gp.loglikelihood(True)
        t_all = time.perf_counter() - t0
        names = ['gemm', 'gram', 'trace', 'potrf_base', 'trsm_base', 'other']
        prof = {}
        for i, nm in enumerate(names):
            cnt, ms, work = ctx.profile_read(i)
            prof[nm] = {'launches': cnt, 'ms': round(ms, 3), 'work': work}
        ctx.profile(False)
        t0 = time.perf_counter()
        gp.set_hyper(h)
        gp.loglikelihood(True)
        t_noprof = time.perf_counter() - t0
        print(json.dumps({'n': n, 'd': d, 'first_eval_s': t_first, 'update_s': t_upd, 'update+grad_s': t_all,
                          'update+grad_noprof_s': t_noprof, 'eff_tflops_n3': n**3/t_noprof/1e12,
                          'gemm_tflops': prof['gemm']['work']/max(prof['gemm']['ms'], 1e-9)/1e9,
                          'lZ': lZ, 'prof': prof}), flush=True)
        Xs = np.random.RandomState(1).rand(4096, d)
        t0 = time.perf_counter()
        gp.posterior(Xs)
        t_pred = time.perf_counter() - t0
        print(json.dumps({'n': n, 'predict_pts': 4096, 'predict_s': t_pred, 'pts_per_s': 4096/t_pred,
                          'predict_tflops': 4096*n*n/t_pred/1e12}), flush=True)
        del gp


if __name__ == '__main__':
    main()
