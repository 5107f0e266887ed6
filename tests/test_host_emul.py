"""CPU-side check of the DEVICE arithmetic: tests/host_emul/emul.cu compiles the per-point source of the CUDA
kernels (pistachio_b200/csrc/pst_chain.cuh) with nvcc's host pass and loops over the grid on the CPU.  This is test
infrastructure (it exists to catch algebra mistakes before GPU time is spent); the product has no CPU path."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from tests import parity

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_emul", "emul.cu")
OUT = os.path.join(HERE, "host_emul", "_build", "libemul.so")
DEPS = [SRC] + [os.path.join(HERE, "..", "pistachio_b200", "csrc", f) for f in ("pst_chain.cuh", "pst_math.cuh", "pst_fit.cuh")]


def load_emul():
    """Build (when stale) and load tests/host_emul/_build/libemul.so; None without nvcc."""
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        return None
    if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in DEPS):
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        subprocess.run([nvcc, "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-Xcompiler", "-fPIC", "-shared", "-o", OUT, SRC],
                       check=True)
    lib = C.CDLL(OUT)
    lib.emul_eval_grid.restype = C.c_int
    return lib


@pytest.fixture(scope="module")
def emul():
    lib = load_emul()
    if lib is None:
        pytest.skip("nvcc not available")
    return lib


def run_emul(lib, wl, n_list, d_list, thetas, pol=2, ncol=1, W=1, flags=0):
    wl = np.ascontiguousarray(wl, dtype=np.float64)
    n = np.ascontiguousarray(np.array(n_list, dtype=np.complex128))
    nre, nim = np.ascontiguousarray(n.real), np.ascontiguousarray(n.imag)
    d = np.ascontiguousarray(d_list, dtype=np.float64)
    ct = np.ascontiguousarray(np.cos(np.atleast_1d(thetas)))
    npol = 2 if pol == 2 else 1
    shape = (npol, len(wl), len(ct))
    R, T, A = np.empty(shape), np.empty(shape), np.empty(shape)
    M = np.empty(shape + (2, 2), dtype=np.complex128)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib.emul_eval_grid(len(wl), len(ct), len(d), p(wl), p(ct), p(nre), p(nim), p(d), pol, ncol, W, flags, p(R), p(T),
                            p(A), p(M))
    assert rc == 0
    return R, T, A, M


@pytest.mark.parametrize("name", ["c1", "c2_sub", "c3_sub_d0", "c3_sub_d66666", "c4_sub"])
def test_emul_matches_reference_cases(emul, ref_cases, name):
    c = ref_cases[name]
    R, T, A, _ = run_emul(emul, c["wl"], list(c["n"]), list(c["d"]), c["theta"])
    rep = parity.compare_grid(name, T, R, c["wl"], list(c["n"]), list(c["d"]), c["theta"], c["T"], c["R"], got_A=A)
    rep.check(max_loose_fraction=0.01)


def test_emul_segmented_combine_matches_pointwise(emul, ref_cases):
    c = ref_cases["c4_sub"]
    R1, T1, _, M1 = run_emul(emul, c["wl"], list(c["n"]), list(c["d"]), c["theta"], ncol=2, W=1)
    for W in (2, 7, 16):
        R, T, A, M = run_emul(emul, c["wl"], list(c["n"]), list(c["d"]), c["theta"], ncol=2, W=W)
        np.testing.assert_allclose(R, R1, rtol=0, atol=2e-12)
        np.testing.assert_allclose(T, T1, rtol=0, atol=2e-12)
    c = ref_cases["c2_sub"]
    R1, T1, _, _ = run_emul(emul, c["wl"], list(c["n"]), list(c["d"]), c["theta"])
    R, T, _, _ = run_emul(emul, c["wl"], list(c["n"]), list(c["d"]), c["theta"], W=4)
    rep = parity.compare_grid("c2_sub_W4", T, R, c["wl"], list(c["n"]), list(c["d"]), c["theta"], c["T"], c["R"])
    rep.check(max_loose_fraction=0.01)


def test_emul_full_matrix_and_numeric_det(emul, ref_cases):
    c = ref_cases["c1"]
    every = int(c["m_every"])
    R, T, A, M = run_emul(emul, c["wl"], list(c["n"]), list(c["d"]), c["theta"], ncol=2, flags=4)
    ref_m = c["M"]
    scale = np.abs(ref_m).max(axis=(-1, -2), keepdims=True)
    assert np.all(np.abs(M[:, ::every] - ref_m) <= 1e-12 * scale)
    np.testing.assert_allclose(T, c["T"], rtol=0, atol=1e-12)


def test_emul_fuzz(emul, ref_cases):
    n = int(ref_cases["fuzz"]["count"])
    loose = total = 0
    for i in range(0, n, 2):
        c = ref_cases[f"fuzz{i}"]
        R, T, A, M = run_emul(emul, c["wl"], list(c["n"]), list(c["d"]), c["theta"], ncol=2)
        rep = parity.compare_grid(f"fuzz{i}", T, R, c["wl"], list(c["n"]), list(c["d"]), c["theta"], c["T"], c["R"], got_A=A)
        assert rep.n_fail == 0, (rep.summary(), rep.worst)
        loose += rep.n_loose
        total += rep.n
        ref_m = c["M"]
        ok = np.isfinite(ref_m).all(axis=(-1, -2))
        scale = np.abs(ref_m).max(axis=(-1, -2), keepdims=True)
        assert np.all(np.abs(M - ref_m)[ok] <= 1e-11 * scale[ok]), f"fuzz{i}: matrix mismatch"
    print(f"fuzz: {total} values, {loose} needed the conditioning clause")
    assert loose <= 0.05 * total


def test_emul_thick_absorber_and_deep_bragg_stay_finite(emul):
    wl = np.linspace(0.8, 1.25, 16)
    # 200 um of k = 10: the reference overflows to NaN; T must underflow to 0 cleanly
    n_list = [np.full(16, 1.0 + 0j), np.full(16, 2.0 + 10.0j), np.full(16, 1.5 + 0j)]
    R, T, A, _ = run_emul(emul, wl, n_list, [0.0, 200.0, 0.0], [0.0, 0.7])
    assert np.isfinite(R).all() and np.isfinite(T).all() and (T == 0).all() and (R > 0).all() and (R < 1).all()
    # 2000-layer high-contrast Bragg stack: reference returns NaN at 7 of 16 wavelengths
    L = 2000
    n_list = [np.full(16, 1.0 + 0j)] + [np.full(16, (2.32 if j % 2 == 0 else 1.36) + 0j) for j in range(L)] + [np.full(16, 1.5 + 0j)]
    d = [0.0] + [1.0 / (4 * (2.32 if j % 2 == 0 else 1.36)) for j in range(L)] + [0.0]
    for W in (1, 8):
        R, T, A, _ = run_emul(emul, wl, n_list, d, [0.0], W=W)
        assert np.isfinite(R).all() and np.isfinite(T).all()
        assert (R >= 0).all() and (R <= 1 + 1e-12).all() and (T >= 0).all()
        np.testing.assert_allclose(R + T, 1.0, atol=1e-9)


def test_emul_periodic_power_matches_reference_and_sequential(emul, ref_cases):
    """Closed-form power of a repeated lossless pair (binary powering of the period, apply_pair_power) vs the reference
    fixture and vs the layer-by-layer product, on the 2 x 20-pair DBR cavity and on a 1000-pair stop band (rescaled
    coefficients)."""
    c = ref_cases["c2_sub"]
    Rs, Ts, _, Ms = run_emul(emul, c["wl"], list(c["n"]), list(c["d"]), c["theta"], ncol=2)
    Rp, Tp, Ap, Mp = run_emul(emul, c["wl"], list(c["n"]), list(c["d"]), c["theta"], ncol=2, flags=256)
    rep = parity.compare_grid("c2_sub_power", Tp, Rp, c["wl"], list(c["n"]), list(c["d"]), c["theta"], c["T"], c["R"], got_A=Ap)
    rep.check(max_loose_fraction=0.01)
    assert np.abs(Rp - Rs).max() < 2e-13 and np.abs(Tp - Ts).max() < 2e-13
    scale = np.abs(Ms).max(axis=(-1, -2), keepdims=True)
    assert np.all(np.abs(Mp - Ms) <= 1e-12 * scale)
    # deep Bragg stack: the recurrence grows like 1.7^1000 and must be rescaled, the result stays finite and agrees
    wl = np.linspace(0.8, 1.25, 12)
    pairs = 1000
    n_list = [np.full(12, 1.0 + 0j)] + [np.full(12, v + 0j) for v in (2.32, 1.36)] * pairs + [np.full(12, 1.5 + 0j)]
    d = [0.0] + [1.0 / (4 * 2.32), 1.0 / (4 * 1.36)] * pairs + [0.0]
    Rs, Ts, _, _ = run_emul(emul, wl, n_list, d, [0.0, 0.6])
    Rp, Tp, _, _ = run_emul(emul, wl, n_list, d, [0.0, 0.6], flags=256)
    assert np.isfinite(Rp).all() and np.isfinite(Tp).all()
    np.testing.assert_allclose(Rp, Rs, rtol=0, atol=1e-12)
    np.testing.assert_allclose(Tp, Ts, rtol=0, atol=1e-12)
    np.testing.assert_allclose(Rp + Tp, 1.0, atol=1e-12)
    # the reference overflows here (no oracle): adjudicate against the 50-digit evaluation of the model
    from oracle import mp_truth

    Tt, Rt = mp_truth.truth_grid(wl, n_list, d, np.cos([0.0, 0.6]))
    err_p = max(np.abs(Rp - Rt).max(), np.abs(Tp - Tt).max())
    err_s = max(np.abs(Rs - Rt).max(), np.abs(Ts - Tt).max())
    assert err_p < 2e-12 and err_p < 2 * err_s + 1e-13, (err_p, err_s)


@pytest.mark.parametrize("pairs", [50, 100, 500, 1500])
def test_emul_low_contrast_periodic_stacks_against_oracle(emul, pairs):
    """Low-contrast quarter-wave stacks around their band edges, both polarisations, theta in {0, 0.4}: the closed-form
    period power (flags = 256: what the typed kernel does by default) must hold the 1e-12 rule against the ORACLE like
    the layer-by-layer product does -- not merely agree with the other path."""
    from tests import periodic_cases as pc

    n_wide, n_edge = pc.CASES[pairs]
    wl = pc.band_edge_wavelengths(1.46, 1.45, n_wide, n_edge)
    cols, lm, ld = pc.quarter_wave_stack(pairs, wl)
    n_list = pc.per_layer(cols, lm)
    T0, R0, M0 = parity.oracle_grid(wl, n_list, ld, pc.THETAS, want_m=True)
    assert np.isfinite(T0).all() and np.isfinite(R0).all()
    sT, sR = parity.sensitivity(wl, n_list, ld, pc.THETAS, T0, R0)
    for flags in (256, 0):
        R, T, A, _ = run_emul(emul, wl, n_list, ld, pc.THETAS, flags=flags)
        rep = parity.Report(f"qw{pairs} flags={flags}")
        rep.add("R", R, R0, 4 * sR)
        rep.add("T", T, T0, 4 * sT + parity.det_noise(M0))
        rep.check(max_loose_fraction=0.01)


def test_emul_period_power_against_mpmath_truth(emul):
    """n = 1.50/1.45 quarter-wave stack, 200 pairs, against a 50-digit evaluation of the model.  The
    closed-form power must be as close to the truth as the layer-by-layer product (round 1's recurrence: 4.6e-12)."""
    from oracle import mp_truth
    from tests import periodic_cases as pc

    wl = pc.band_edge_wavelengths(1.50, 1.45, 10, 3, lo=0.90, hi=1.10)
    cols, lm, ld = pc.quarter_wave_stack(200, wl, n_hi=1.50, n_lo=1.45)
    n_list = pc.per_layer(cols, lm)
    Tt, Rt = mp_truth.truth_grid(wl, n_list, ld, np.cos(pc.THETAS))
    Rp, Tp, _, _ = run_emul(emul, wl, n_list, ld, pc.THETAS, flags=256)
    Rs, Ts, _, _ = run_emul(emul, wl, n_list, ld, pc.THETAS, flags=0)
    err_p = max(np.abs(Rp - Rt).max(), np.abs(Tp - Tt).max())
    err_s = max(np.abs(Rs - Rt).max(), np.abs(Ts - Tt).max())
    print(f"200 pairs 1.50/1.45 vs mpmath: closed-form power {err_p:.2e}, layer by layer {err_s:.2e}")
    assert err_p < 1e-12 and err_p < 3 * err_s + 2e-14


def test_compile_time_pair_power_equals_the_bit_loop(emul):
    """pair_power_small (pair counts 1..kPowerFixedMax resolved at compile time: the cavity and typed kernels' default)
    performs the operations of the generic left-to-right bit loop in the same order: (t_k, s_k) must be equal bit for
    bit for every compiled count, in pass bands (|x| < 1), at band edges and inside stop bands; counts above the range
    are refused.  Both agree with the Chebyshev polynomials t_k = T_k(x), s_k = U_{k-1}(x) when s2 = 1 - x^2."""
    rng = np.random.default_rng(5)
    xs = np.concatenate([rng.uniform(-1.0, 1.0, 12), [-1.0, 1.0, 0.0, -0.999999, 1.000001], rng.uniform(1.0, 2.5, 6) * rng.choice([-1, 1], 6)])
    out_f, out_g = (C.c_double * 3)(), (C.c_double * 3)()
    emul.emul_pair_power.argtypes = [C.c_double, C.c_double, C.c_int, C.c_int, C.POINTER(C.c_double)]
    emul.emul_pair_power.restype = None
    for k in range(1, 41):
        for x in xs:
            for s2 in (1.0 - x * x, (1.0 - x * x) * (1.0 + 3e-13)):
                emul.emul_pair_power(float(x), float(s2), k, 1, out_f)
                emul.emul_pair_power(float(x), float(s2), k, 0, out_g)
                assert out_f[2] == 1.0
                assert out_f[0] == out_g[0] and out_f[1] == out_g[1], (k, x)
        x = 0.3
        emul.emul_pair_power(x, 1.0 - x * x, k, 1, out_f)
        th = np.arccos(x)
        np.testing.assert_allclose([out_f[0], out_f[1]], [np.cos(k * th), np.sin(k * th) / np.sin(th)], rtol=0, atol=1e-13 * k)
    emul.emul_pair_power(0.3, 0.91, 41, 1, out_f)
    assert out_f[2] == 0.0
