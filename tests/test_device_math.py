"""Accuracy of the hand-written elementary functions of the layer loop (pistachio_b200/csrc/pst_math.cuh), run through
the host-emulation build of the SAME source and compared with mpmath at 60 digits."""
import ctypes as C

import mpmath as mp
import numpy as np
import pytest

from tests.test_host_emul import emul  # noqa: F401  (fixture: builds tests/host_emul/_build/libemul.so)

mp.mp.dps = 60


def ulp_err(got, exact):
    exact_f = float(exact)
    ulp = np.spacing(abs(exact_f)) if exact_f != 0 else np.spacing(1e-300)
    return abs(mp.mpf(got) - exact) / mp.mpf(ulp)


def test_sincos_fast_accuracy(emul):
    rng = np.random.default_rng(11)
    x = np.concatenate([
        rng.uniform(-np.pi, np.pi, 400), rng.uniform(-100, 100, 400), rng.uniform(-1e5, 1e5, 300),
        rng.uniform(-1e9, 1e9, 200),                         # still the fast path (|x| < 2^30)
        rng.uniform(2e9, 1e15, 50),                          # library slow path
        np.arange(-40, 41) * (np.pi / 2),                    # multiples of pi/2: worst cancellation in the reduction
        np.arange(1, 40) * (np.pi / 2) * (1 + 1e-15), [0.0, 1e-300, -1e-10, 2.0 ** 29.9],
    ])
    s, c = np.empty_like(x), np.empty_like(x)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    emul.emul_sincos(len(x), p(x), p(s), p(c))
    worst = 0.0
    for xi, si, ci in zip(x, s, c):
        xm = mp.mpf(float(xi))
        es, ec = mp.sin(xm), mp.cos(xm)
        # near a zero of sin/cos the absolute error is what matters (it enters the matrix additively)
        err_s = min(ulp_err(si, es), abs(mp.mpf(float(si)) - es) / mp.mpf(2.0 ** -53))
        err_c = min(ulp_err(ci, ec), abs(mp.mpf(float(ci)) - ec) / mp.mpf(2.0 ** -53))
        worst = max(worst, float(err_s), float(err_c))
    print(f"sincos_fast: worst error {worst:.3f} ulp (relative ulp, or 2^-53 absolute near zeros)")
    assert worst < 1.5


def test_cosh_sinh_scaled_accuracy_and_range(emul):
    rng = np.random.default_rng(12)
    y = np.concatenate([rng.uniform(-30, 30, 500), rng.uniform(-1e-6, 1e-6, 100), rng.uniform(-5000, 5000, 200),
                        [0.0, 1e-300, 709.0, -709.0, 745.2, 1e5, -1e5]])
    ch, sh = np.empty_like(y), np.empty_like(y)
    sc = np.empty(len(y), dtype=np.int32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    emul.emul_cosh_sinh(len(y), p(y), p(ch), p(sh), p(sc))
    assert np.isfinite(ch).all() and np.isfinite(sh).all()  # the power of two stays in the integer: no overflow
    assert (ch > 0.70).all() and (ch < 2.9).all()  # e^r (1 + f), |r| <= ln2/2
    worst = 0.0
    for yi, c_, s_, k in zip(y, ch, sh, sc):
        ym = mp.mpf(float(yi))
        ec, es = mp.cosh(ym) / mp.mpf(2) ** int(k), mp.sinh(ym) / mp.mpf(2) ** int(k)
        worst = max(worst, float(ulp_err(c_, ec)))
        worst = max(worst, float(min(ulp_err(s_, es), abs(mp.mpf(float(s_)) - es) / mp.mpf(2.0 ** -53))))
    print(f"cosh_sinh_scaled: worst error {worst:.3f} ulp")
    assert worst < 3.0  # Horner + even/odd recombination: ~2.5 ulp worst case (glibc exp used by the reference: < 1 ulp)
