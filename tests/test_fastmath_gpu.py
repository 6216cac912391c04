"""Accuracy of the short FP64 elementary functions used in the covariance epilogues
(pygp_b200/csrc/fastmath.cuh) against numpy's libm, in units in the last place.  They
replace the library sqrt / exp inside Kernel.get / grad (pygp/kernels/se.py:53-66,
matern.py:44-90); the parity tolerance of those kernels is 1e-10 relative, the
functions are held to a few ulp here."""

import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run(which, x):
    from pygp_b200 import _lib
    ctx, L = _lib.context(), _lib.lib()
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty_like(x)
    _lib.check(ctx, L.pgp_dev_fastmath(ctx.handle, which, _lib.ptr(x), x.size, _lib.ptr(out)))
    return out


def _ulps(got, want):
    return np.abs(got - want)/np.spacing(np.abs(want))


def test_exp_tab():
    rng = np.random.RandomState(0)
    x = np.r_[rng.uniform(-700, 700, 400000), rng.uniform(-40, 5, 400000), rng.uniform(-1e-3, 1e-3, 100000),
              np.linspace(-745.2, -700, 5001), [0.0, -0.0, 1.0, -1.0, 709.7, -708.3, -708.5, -744.0, -746.0, -800.0, 710.0]]
    with np.errstate(over='ignore'):
        got, want = _run(0, x), np.exp(x)
    assert np.isinf(got[np.isinf(want)]).all()
    normal = (want > 2.3e-308) & np.isfinite(want)
    assert _ulps(got[normal], want[normal]).max() <= 2.0
    # below the normal range the kernel falls back to the library exp: denormals / zero as libm
    tiny = want <= 2.3e-308
    np.testing.assert_allclose(got[tiny], want[tiny], rtol=1e-12, atol=5e-324)
    assert np.isinf(_run(0, np.array([710.0, 1e4]))).all()
    assert np.isnan(_run(0, np.array([np.nan])))[0]
    # clamped variant: exact above -708, exp(-708) below, NaN kept
    c = _run(2, np.array([-1.5, -707.9, -708.1, -5000.0, np.nan]))
    assert _ulps(c[:2], np.exp([-1.5, -707.9])).max() <= 2.0
    np.testing.assert_allclose(c[2:4], np.exp(-708.0), rtol=1e-15)
    assert np.isnan(c[4])


def test_sqrt_pos():
    rng = np.random.RandomState(1)
    x = np.r_[rng.uniform(0, 4, 300000), 10.0**rng.uniform(-280, 300, 300000), [1.0, 4.0, 2.0, 1e-280, 5e5]]
    got, want = _run(1, x), np.sqrt(x)
    assert _ulps(got, want).max() <= 2.0
    # squared distances below 2^-943 (points closer than 1e-142) come back as ~0, not to 2 ulp: far under
    # every r < 1e-12 guard of the kernels
    assert _run(1, np.array([1e-300]))[0] < 1e-140
    assert _run(1, np.array([0.0]))[0] == 0.0          # exact zero at coincident points (Matern r = 0 guards)
    assert np.isnan(_run(1, np.array([np.nan])))[0]


def test_log_ge1_tab():
    """log for x >= 1 with ABSOLUTE accuracy (what exp(-alpha log E) of the RQ kernel needs, rq.py:56-63)."""
    rng = np.random.RandomState(2)
    x = np.r_[1 + rng.uniform(0, 1e-6, 100000), rng.uniform(1, 4, 300000), 10.0**rng.uniform(0, 300, 300000),
              [1.0, 2.0, np.e, 1 + 2.0**-52, 1e308]]
    got, want = _run(3, x), np.log(x)
    assert np.abs(got - want).max() <= 4e-16*np.maximum(1.0, np.abs(want)).max()
    assert (np.abs(got - want) <= 3e-16*np.maximum(1.0, np.abs(want))).all()
    assert _run(3, np.array([1.0]))[0] == 0.0
    assert np.isnan(_run(3, np.array([np.nan])))[0] and np.isinf(_run(3, np.array([np.inf])))[0]


def test_sin_unsigned_cw():
    """|sin x| by Cody-Waite reduction in pi/2 (the Periodic kernel squares it, periodic.py:53-59)."""
    rng = np.random.RandomState(3)
    x = np.r_[rng.uniform(0, 10, 300000), rng.uniform(0, 1e4, 300000), rng.uniform(0, 1.5e6, 100000),
              [0.0, np.pi/2, np.pi, 1e-300, 3e6, 1e12]]
    got, want = _run(4, x), np.abs(np.sin(x))
    assert np.abs(got - want).max() <= 3e-16
