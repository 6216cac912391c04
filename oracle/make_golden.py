"""
TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.npz from the UNMODIFIED
reference (imported from /root/reference, build container only) and assert
that the numpy restatement in oracle/pygp_oracle.py reproduces it.

    python -m oracle.make_golden            # writes tests/golden/

Each fixture stores the reference's own outputs, so the committed files carry
the reference's arithmetic to the GPU box where /root/reference is absent.
"""

import os
import sys

import numpy as np

from . import ref_loader
from .cases import (KERNEL_CASES, GP_CASES, DTC_CASES, GP_SN, GP_MEAN, kernel_inputs,
                    gp_inputs)
from .pygp_oracle import make_kernel, OExactGP, OFITC, ODTC

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(os.path.dirname(HERE), 'tests', 'golden')


def _close(a, b, what, rtol=1e-12, atol=1e-13):
    a, b = np.asarray(a), np.asarray(b)
    if not np.allclose(a, b, rtol=rtol, atol=atol):
        err = np.max(np.abs(a-b) / (atol/rtol + np.abs(b)))
        raise AssertionError('oracle != reference for %s (scaled err %.3g)' % (what, err))


def kernel_golden(pygp):
    out = {}
    for name, spec in KERNEL_CASES.items():
        rk = ref_loader.ref_kernel(pygp, spec)
        ok = make_kernel(spec)
        x1, x2 = kernel_inputs(rk.ndim)
        _close(ok.get_hyper(), rk.get_hyper(), name + '.hyper', 0, 0)
        rec = {
            'hyper': rk.get_hyper(),
            'get12': rk.get(x1, x2),
            'get11': rk.get(x1),
            'grad12': np.array(list(rk.grad(x1, x2))),
            'grad11': np.array(list(rk.grad(x1))),
            'dget': rk.dget(x1),
            'dgrad': np.array(list(rk.dgrad(x1))),
            'gradx12': rk.gradx(x1, x2),
            'grady12': rk.grady(x1, x2),
            'gradx11': rk.gradx(x1),
        }
        _close(ok.gradx(x1, x2), rec['gradx12'], name + '.gradx12')
        _close(ok.grady(x1, x2), rec['grady12'], name + '.grady12')
        _close(ok.gradx(x1), rec['gradx11'], name + '.gradx11')
        try:                                   # SE and its composites only (tests/test_kernels.py:130-146)
            rec['gradxy12'] = rk.gradxy(x1, x2)
            _close(ok.gradxy(x1, x2), rec['gradxy12'], name + '.gradxy12', 1e-11, 1e-13)
        except NotImplementedError:
            try:
                ok.gradxy(x1, x2)
                raise AssertionError('oracle defines gradxy where the reference does not: ' + name)
            except NotImplementedError:
                pass
        _close(ok.get(x1, x2), rec['get12'], name + '.get12')
        _close(ok.get(x1), rec['get11'], name + '.get11')
        _close(np.array(ok.grad(x1, x2)), rec['grad12'], name + '.grad12')
        _close(np.array(ok.grad(x1)), rec['grad11'], name + '.grad11')
        _close(ok.dget(x1), rec['dget'], name + '.dget')
        _close(np.array(ok.dgrad(x1)), rec['dgrad'], name + '.dgrad')
        # a second hyper vector, through set_hyper (exercises the log-space layout)
        h2 = rk.get_hyper() + 0.1*np.random.RandomState(7).randn(rk.nhyper)
        rk2 = rk.copy(h2)
        ok2 = ok.copy_with(h2)
        rec['hyper2'] = h2
        rec['get12_h2'] = rk2.get(x1, x2)
        rec['grad12_h2'] = np.array(list(rk2.grad(x1, x2)))
        _close(ok2.get(x1, x2), rec['get12_h2'], name + '.get12_h2')
        _close(np.array(ok2.grad(x1, x2)), rec['grad12_h2'], name + '.grad12_h2')
        for k, v in rec.items():
            out['%s/%s' % (name, k)] = v
    return out


def gp_golden(pygp):
    out = {}
    for name, (spec, N, d, fitc) in GP_CASES.items():
        X, y, Xs, U = gp_inputs(N, d, fitc)
        like = pygp.likelihoods.Gaussian(GP_SN)
        rk = ref_loader.ref_kernel(pygp, spec)
        if fitc:
            rgp = pygp.inference.FITC(like, rk, GP_MEAN, U)
            ogp = OFITC(GP_SN, make_kernel(spec), GP_MEAN, U)
        else:
            rgp = pygp.inference.ExactGP(like, rk, GP_MEAN)
            ogp = OExactGP(GP_SN, make_kernel(spec), GP_MEAN)
        rgp.add_data(X, y)
        ogp.add_data(X, y)
        lZ, dlZ = rgp.loglikelihood(True)
        mu, s2, dmu, ds2 = rgp.posterior(Xs, grad=True)
        olZ, odlZ = ogp.loglikelihood(True)
        omu, os2, odmu, ods2 = ogp.posterior(Xs, grad=True)
        _close(odmu, dmu, name + '.dmu', 1e-10, 1e-12)
        _close(ods2, ds2, name + '.ds2', 1e-10, 1e-12)
        _close(olZ, lZ, name + '.lZ')
        _close(odlZ, dlZ, name + '.dlZ', 1e-10, 1e-11)
        _close(omu, mu, name + '.mu')
        _close(os2, s2, name + '.s2', 1e-11, 1e-13)
        rec = {'hyper': rgp.get_hyper(), 'lZ': lZ, 'dlZ': dlZ, 'mu': mu, 's2': s2, 'dmu': dmu, 'ds2': ds2}
        # joint posterior and draws (exact.py:64-79 / fitc.py:102-120, _base.py:143-177)
        Xj = np.random.RandomState(13).rand(6, d)
        fmu, fS = rgp._full_posterior(Xj)
        omu_f, oS = ogp.full_posterior(Xj)
        _close(omu_f, fmu, name + '.full_mu')
        _close(oS, fS, name + '.full_Sigma', 1e-10, 1e-13)
        smp = rgp.sample(Xj, 3, latent=False, rng=5)
        _close(ogp.sample(Xj, 3, latent=False, rng=5), smp, name + '.sample', 1e-8, 1e-9)
        rec.update({'Xj': Xj, 'full_mu': fmu, 'full_Sigma': fS, 'sample': smp})
        if not fitc:
            rec['R'] = rgp._R if N <= 64 else rgp._R[:8, :8]
            rec['a'] = rgp._a
        # second hyper vector through set_hyper (the optimiser's access path)
        h2 = rgp.get_hyper() + 0.05*np.random.RandomState(11).randn(rgp.nhyper)
        rgp.set_hyper(h2)
        ogp.set_hyper(h2)
        lZ2, dlZ2 = rgp.loglikelihood(True)
        mu2, s22 = rgp.posterior(Xs)
        olZ2, odlZ2 = ogp.loglikelihood(True)
        _close(olZ2, lZ2, name + '.lZ_h2')
        _close(odlZ2, dlZ2, name + '.dlZ_h2', 1e-10, 1e-11)
        rec.update({'hyper2': h2, 'lZ_h2': lZ2, 'dlZ_h2': dlZ2, 'mu_h2': mu2, 's2_h2': s22})
        for k, v in rec.items():
            out['%s/%s' % (name, k)] = np.asarray(v)
    return out


def dtc_golden(pygp):
    out = {}
    for name, (spec, N, d) in DTC_CASES.items():
        X, y, Xs, U = gp_inputs(N, d, True)
        rgp = pygp.inference.DTC(pygp.likelihoods.Gaussian(GP_SN), ref_loader.ref_kernel(pygp, spec), GP_MEAN, U)
        ogp = ODTC(GP_SN, make_kernel(spec), GP_MEAN, U)
        rgp.add_data(X, y)
        ogp.add_data(X, y)
        rec = {}
        for tag, h in (('', None), ('_h2', rgp.get_hyper() + 0.05*np.random.RandomState(11).randn(rgp.nhyper))):
            if h is not None:
                rgp.set_hyper(h)
                ogp.set_hyper(h)
                rec['hyper2'] = h
            lZ, dlZ = rgp.loglikelihood(True)
            mu, s2, dmu, ds2 = rgp.posterior(Xs, grad=True)
            fmu, fS = rgp._full_posterior(Xs)
            olZ, odlZ = ogp.loglikelihood(True)
            omu, os2, odmu, ods2 = ogp.posterior(Xs, grad=True)
            _close(olZ, lZ, name + '.lZ' + tag)
            _close(odlZ, dlZ, name + '.dlZ' + tag, 1e-9, 1e-10)
            _close(omu, mu, name + '.mu' + tag, 1e-10, 1e-12)
            _close(os2, s2, name + '.s2' + tag, 1e-9, 1e-12)
            _close(odmu, dmu, name + '.dmu' + tag, 1e-9, 1e-11)
            _close(ods2, ds2, name + '.ds2' + tag, 1e-8, 1e-11)
            _close(ogp.full_posterior(Xs)[1], fS, name + '.Sigma' + tag, 1e-8, 1e-11)
            rec.update({'lZ' + tag: lZ, 'dlZ' + tag: dlZ, 'mu' + tag: mu, 's2' + tag: s2, 'dmu' + tag: dmu,
                        'ds2' + tag: ds2, 'full_Sigma' + tag: fS})
        rec['hyper'] = ogp.get_hyper() if False else None
        rec.pop('hyper')
        for k, v in rec.items():
            out['%s/%s' % (name, k)] = np.asarray(v)
    return out


def main():
    pygp = ref_loader.load()
    os.makedirs(GOLDEN, exist_ok=True)
    kg = kernel_golden(pygp)
    np.savez_compressed(os.path.join(GOLDEN, 'kernels.npz'), **kg)
    gg = gp_golden(pygp)
    gg.update(dtc_golden(pygp))
    np.savez_compressed(os.path.join(GOLDEN, 'gp.npz'), **gg)
    # the reference's own demo data set (pygp/demos/xy.npz, X:(20,1) y:(20,)), used by
    # its tests/test_learning.py:26-30 -- a data fixture, copied verbatim
    import shutil
    shutil.copyfile(os.path.join(ref_loader.REFERENCE_ROOT, 'pygp', 'demos', 'xy.npz'),
                    os.path.join(GOLDEN, 'xy.npz'))
    print('wrote %d kernel arrays, %d gp arrays to %s' % (len(kg), len(gg), GOLDEN))


if __name__ == '__main__':
    sys.exit(main())
