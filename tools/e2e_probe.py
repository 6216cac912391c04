"""Where does the e2e step (new model + add_data + loglikelihood(True)) spend its wall time?"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pygp_b200 as pygp
from pygp_b200 import _lib
n, d = 32768, 16
rng = np.random.RandomState(0)
X = rng.rand(n, d); y = np.sin(3*X.sum(1)) + 0.1*rng.randn(n)
ell = list(0.5*np.sqrt(d)*np.ones(d))
ctx = _lib.context()
for rep in range(4):
    t0 = time.perf_counter()
    gp = pygp.inference.ExactGP(pygp.likelihoods.Gaussian(0.1), pygp.kernels.Matern(1.0, ell, 5), 0.0)
    t1 = time.perf_counter()
    gp.add_data(X, y)
    t2 = time.perf_counter()
    lZ, dlZ = gp.loglikelihood(True)
    t3 = time.perf_counter()
    del gp
    ctx.sync()
    t4 = time.perf_counter()
    print('rep %d: ctor %.4f add_data %.4f loglike(grad) %.4f del %.4f total %.4f' % (rep, t1-t0, t2-t1, t3-t2, t4-t3, t4-t0), flush=True)
