"""Development probe 2 (run on the GPU box): per-class time breakdown of
predict, the batched small-N path (config C4 shape) and the Gram build for a
few kernels.  Prints JSON lines."""

import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pygp_b200 import _lib  # noqa: E402
import pygp_b200 as pygp     # noqa: E402

NAMES = ['gemm', 'gram', 'trace', 'potrf_base', 'trsm_base', 'other']


def prof_dict(ctx):
    out = {}
    for i, nm in enumerate(NAMES):
        cnt, ms, work = ctx.profile_read(i)
        out[nm] = {'launches': cnt, 'ms': round(ms, 3), 'work': work}
    return out


def problem(n, d, seed=0):
    rng = np.random.RandomState(seed)
    X = rng.rand(n, d)
    y = np.sin(3*X.sum(1)) + 0.1*rng.randn(n)
    return X, y


def predict_probe(ctx, L, n, d, m):
    X, y = problem(n, d)
    ell = [0.5*np.sqrt(d)]*d
    gp = pygp.inference.ExactGP(pygp.likelihoods.Gaussian(0.1), pygp.kernels.Matern(1.0, ell, 5), 0.0)
    gp.add_data(X, y)
    Xs = torch.rand(m, d, dtype=torch.float64, device='cuda')
    mu = torch.empty(m, dtype=torch.float64, device='cuda')
    s2 = torch.empty(m, dtype=torch.float64, device='cuda')

    def run():
        _lib.check(ctx, L.pgp_exact_predict_dev(gp._dev.handle, Xs.data_ptr(), m, mu.data_ptr(), s2.data_ptr()))
    run()
    t0 = time.perf_counter()
    run()
    t_plain = time.perf_counter() - t0
    ctx.profile(True)
    t0 = time.perf_counter()
    run()
    t_prof = time.perf_counter() - t0
    p = prof_dict(ctx)
    ctx.profile(False)
    print(json.dumps({'probe': 'predict', 'n': n, 'd': d, 'm': m, 'wall_s': t_plain, 'wall_prof_s': t_prof,
                      'pts_per_s': m/t_plain, 'tflops': m*float(n)*n/t_plain/1e12, 'prof': p}), flush=True)
    del gp


def batched_probe(ctx, L, n, d, B):
    X, y = problem(n, d)
    kern = pygp.kernels.SE(1.0, [0.5*np.sqrt(d)]*d)
    spec = kern._spec()
    rng = np.random.RandomState(2)
    base = np.r_[np.log(0.1), 0.0, np.log(0.5*np.sqrt(d))*np.ones(d), 0.0]
    hyps = np.ascontiguousarray(base + rng.uniform(-0.5, 0.5, size=(B, d + 3)))
    lZ = np.empty(B)
    info = np.zeros(B, dtype=np.int32)
    Xc, yc = _lib.as_f64(X, 2), _lib.as_f64(y)

    def run():
        _lib.check(ctx, L.pgp_batched_loglike(ctx.handle, spec, _lib.ptr(Xc), _lib.ptr(yc), n, _lib.ptr(hyps), B,
                                              _lib.ptr(lZ), None, info.ctypes.data_as(_lib._ip)))
    run()
    t0 = time.perf_counter()
    run()
    t_plain = time.perf_counter() - t0
    ctx.profile(True)
    run()
    p = prof_dict(ctx)
    ctx.profile(False)
    print(json.dumps({'probe': 'batched_loglike', 'n': n, 'd': d, 'B': B, 'wall_s': t_plain,
                      'samples_per_s': B/t_plain, 'tflops_potrf': B*float(n)**3/3/t_plain/1e12,
                      'info_nonzero': int((info != 0).sum()), 'prof': p}), flush=True)


def gram_probe(ctx, L, kern, n, label):
    d = kern.ndim
    X, _ = problem(n, d)
    Xd = torch.tensor(X, device='cuda')
    out = torch.empty(n, n, dtype=torch.float64, device='cuda')
    hyp = _lib.as_f64(kern.get_hyper())
    spec = kern._spec()

    def run():
        _lib.check(ctx, L.pgp_gram_dev(ctx.handle, spec, _lib.ptr(hyp), Xd.data_ptr(), n, None, n, out.data_ptr()))
    run()
    ctx.profile(True)
    run()
    cnt, ms, work = ctx.profile_read(1)
    ctx.profile(False)
    print(json.dumps({'probe': 'gram', 'kernel': label, 'n': n, 'd': d, 'ms': ms, 'GBps': work/ms/1e6}), flush=True)
    del Xd, out
    torch.cuda.empty_cache()


def fitc_probe(ctx, L, n, p, d):
    X, y = problem(n, d)
    U = np.random.RandomState(3).rand(p, d)
    gp = pygp.inference.FITC(pygp.likelihoods.Gaussian(0.1), pygp.kernels.SE(1.0, [0.6]*d), 0.0, U)
    t0 = time.perf_counter()
    gp.add_data(X, y)
    t_first = time.perf_counter() - t0
    h = gp.get_hyper()
    gp.set_hyper(h + 0.01)
    gp.loglikelihood(True)
    ctx.profile(True)
    t0 = time.perf_counter()
    gp.set_hyper(h)
    t_upd = time.perf_counter() - t0
    lZ, dlZ = gp.loglikelihood(True)
    t_all = time.perf_counter() - t0
    pr = prof_dict(ctx)
    ctx.profile(False)
    Xs = np.random.RandomState(1).rand(65536, d)
    gp.posterior(Xs[:1024])
    t0 = time.perf_counter()
    gp.posterior(Xs)
    t_pred = time.perf_counter() - t0
    flops = 8.0*p*p*n      # update: trsm p^2 n + syrk p^2 n; grad: 2 trsm + 3 tn-gemm + 1 nt-gemm (2 p^2 n) ~ 8 p^2 n
    print(json.dumps({'probe': 'fitc', 'n': n, 'p': p, 'd': d, 'first_s': t_first, 'update_s': t_upd,
                      'update+grad_s': t_all, 'eff_tflops_8p2n': flops/t_all/1e12, 'lZ': lZ,
                      'predict_pts_per_s': len(Xs)/t_pred, 'prof': pr}), flush=True)
    del gp


def main():
    what = sys.argv[1:] or ['predict', 'batched', 'gram', 'fitc']
    ctx, L = _lib.context(), _lib.lib()
    if 'predict' in what:
        predict_probe(ctx, L, 8192, 16, 16384)
        predict_probe(ctx, L, 32768, 16, 16384)
    if 'batched' in what:
        batched_probe(ctx, L, 2048, 8, 256)
        batched_probe(ctx, L, 512, 8, 1024)
    if 'fitc' in what:
        fitc_probe(ctx, L, 65536, 512, 8)
        fitc_probe(ctx, L, 262144, 2048, 8)
    if 'gram' in what:
        pk = pygp.kernels
        n = 32768
        gram_probe(ctx, L, pk.SE(1.0, 0.5, 1), n, 'se_iso_d1')
        gram_probe(ctx, L, pk.SE(1.0, [1.4]*8), n, 'se_ard_d8')
        gram_probe(ctx, L, pk.Matern(1.0, [2.0]*16, 5), n, 'matern5_ard_d16')
        gram_probe(ctx, L, pk.Matern(1.0, [2.0]*16, 1), n, 'matern1_ard_d16')
        gram_probe(ctx, L, pk.RQ(1.0, [1.4]*8, 0.5), n, 'rq_ard_d8')
        gram_probe(ctx, L, pk.SE(1.0, 0.5, 1) + pk.Periodic(0.5, 1.0, 0.25), n, 'se_plus_periodic_d1')


if __name__ == '__main__':
    main()
