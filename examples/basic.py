"""The reference's basic demo (pygp/demos/basic.py) on the B200 path, without plotting:
fit a GP to the demo data set by type-II maximum likelihood, predict, and
marginalise the hyper-parameters by MCMC.  Run on a machine with a B200:

    python examples/basic.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pygp_b200 as pygp            # same names as `import pygp`   # noqa: E402


def main():
    data = np.load(os.path.join(ROOT, 'tests', 'golden', 'xy.npz'))      # pygp/demos/xy.npz
    X, y = data['X'], data['y']

    gp = pygp.BasicGP(sn=.1, sf=1, ell=.1, mu=0)
    gp.add_data(X, y)
    print('initial   lZ = %.4f' % gp.loglikelihood())
    pygp.optimize(gp, {'sn': None})                 # hold the noise fixed, as the demo does
    print('optimised lZ = %.4f  hypers = %s' % (gp.loglikelihood(), np.round(np.exp(gp.get_hyper()[:3]), 4)))

    x = np.linspace(X.min(), X.max(), 500)[:, None]
    mu, s2 = gp.posterior(x)
    print('posterior on 500 points: mean in [%.3f, %.3f], max std %.3f' % (mu.min(), mu.max(), np.sqrt(s2.max())))
    mu, s2, dmu, ds2 = gp.posterior(x[:3], grad=True)
    print('d mu / dx at the first points:', np.round(dmu.ravel(), 4))

    priors = {'sn': pygp.priors.Uniform(0.01, 1.0), 'sf': pygp.priors.Uniform(0.01, 5.0),
              'ell': pygp.priors.Uniform(0.01, 1.0), 'mu': pygp.priors.Uniform(-2, 2)}
    mcmc = pygp.meta.MCMC(gp, priors, n=50, burn=20, rng=0)
    mu_m, s2_m = mcmc.posterior(x)
    print('MCMC mixture over 50 hyper samples: max std %.3f' % np.sqrt(s2_m.max()))

    f = gp.sample(x[::10], 3, rng=1)
    print('3 joint posterior draws at 50 points:', f.shape)


if __name__ == '__main__':
    main()
