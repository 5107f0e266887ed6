"""Refraction mode (PST_FLAG_SNELL).  The reference has no counterpart (it keeps the incidence
angle in every layer, tm.py:270), so the pins are: the reference-pinned oracle at normal incidence, Fresnel/Airy closed
forms, an independently formulated textbook implementation and a 50-digit mpmath evaluation (oracle/snell_oracle.py).
CPU tests run the device source through tests/host_emul; GPU tests call the CUDA path through the C ABI."""
import numpy as np
import pytest

from oracle import snell_oracle as so
from oracle import tmm_oracle as to
from tests.test_host_emul import emul, run_emul  # noqa: F401  (fixture)

SNELL = 1 << 12
POLS = (so.S_WAVE, so.P_WAVE)


def random_stack(rng, L, lossless=False, n0=1.0, nf=1.5):
    ns = [complex(n0)]
    for _ in range(L - 2):
        k = 0.0 if (lossless or rng.random() < 0.5) else 10 ** rng.uniform(-4, 0.5)
        ns.append(complex(rng.uniform(1.0, 4.0), k))
    ns.append(complex(nf))
    d = [0.0] + list(rng.uniform(0.02, 1.5, L - 2)) + [0.0]
    return ns, d


def oracle_grid(wl, ns, d, thetas):
    R = np.empty((2, len(wl), len(thetas)))
    T = np.empty_like(R)
    for ip, pol in enumerate(POLS):
        for il, lam in enumerate(wl):
            for it, th in enumerate(thetas):
                T[ip, il, it], R[ip, il, it], _ = so.point(lam, ns, d, th, pol)
    return R, T


def close(got, ref, tol=1e-12, floor=1e-3):
    return np.abs(got - ref) <= tol * np.maximum(np.abs(ref), floor)


# ------------------------------------------------------------------------------------------ oracle pins (CPU)
def test_oracle_single_interface_closed_forms():
    n1, n2 = 1.0, 1.5
    for th in np.linspace(0.0, 1.5, 31):
        c1 = np.cos(th)
        c2 = np.sqrt(1 - (n1 * np.sin(th) / n2) ** 2)
        rs = (n1 * c1 - n2 * c2) / (n1 * c1 + n2 * c2)
        rp = (n2 * c1 - n1 * c2) / (n2 * c1 + n1 * c2)
        Ts, Rs, _ = so.point(1.0, [n1 + 0j, n2 + 0j], [0, 0], th, so.S_WAVE)
        Tp, Rp, _ = so.point(1.0, [n1 + 0j, n2 + 0j], [0, 0], th, so.P_WAVE)
        assert abs(Rs - rs ** 2) < 1e-14 and abs(Rp - rp ** 2) < 1e-14
        assert abs(Rs + Ts - 1) < 1e-14 and abs(Rp + Tp - 1) < 1e-14
    # Brewster angle and total internal reflection (glass -> air)
    assert so.point(1.0, [1.0 + 0j, 1.5 + 0j], [0, 0], np.arctan(1.5), so.P_WAVE)[1] < 1e-30
    for pol in POLS:
        T, R, _ = so.point(1.0, [1.5 + 0j, 1.0 + 0j], [0, 0], np.deg2rad(45.0), pol)
        assert abs(R - 1) < 1e-14 and abs(T) < 1e-14


def test_oracle_airy_slab():
    n0, n1, n2, d, lam = 1.0, 2.2, 1.5, 0.37, 0.8
    for th in (0.0, 0.4, 1.1):
        s = n0 * np.sin(th)
        kz = [np.sqrt(n * n - s * s) for n in (n0, n1, n2)]
        beta = 2 * np.pi / lam * kz[1] * d
        r01 = (kz[0] - kz[1]) / (kz[0] + kz[1])
        r12 = (kz[1] - kz[2]) / (kz[1] + kz[2])
        r = (r01 + r12 * np.exp(2j * beta)) / (1 + r01 * r12 * np.exp(2j * beta))
        T, R, _ = so.point(lam, [n0 + 0j, n1 + 0j, n2 + 0j], [0, d, 0], th, so.S_WAVE)
        assert abs(R - abs(r) ** 2) < 1e-14 and abs(R + T - 1) < 1e-14


def test_oracle_formulations_agree():
    rng = np.random.default_rng(5)
    for _ in range(40):
        ns, d = random_stack(rng, int(rng.integers(3, 10)))
        lam, th = rng.uniform(0.4, 2.0), rng.uniform(0, 1.5)
        for pol in POLS:
            T, R, _ = so.point(lam, ns, d, th, pol)
            Tb, Rb = so.point_textbook(lam, ns, d, th, pol)
            Tm, Rm = so.truth_point(lam, ns, d, th, pol)
            assert abs(R - Rb) < 1e-11 and abs(T - Tb) < 1e-11
            assert abs(R - Rm) < 1e-11 and abs(T - Tm) < 1e-11


def test_oracle_normal_incidence_is_the_reference_model():
    rng = np.random.default_rng(6)
    wl = np.linspace(0.5, 1.5, 7)
    ns, d = random_stack(rng, 6)
    n_layers = [np.full(wl.shape, n) for n in ns]
    R, T = oracle_grid(wl, ns, d, [0.0])
    for ip, pol in enumerate(POLS):
        Tr, Rr = to.spectrum(to.Stack(wl, n_layers, d), 0.0, pol)
        assert close(R[ip, :, 0], Rr).all() and close(T[ip, :, 0], Tr).all()


# ------------------------------------------------------------------------------------------ device source on the CPU
def test_emul_snell_matches_oracle(emul):  # noqa: F811
    rng = np.random.default_rng(7)
    wl = np.linspace(0.6, 1.4, 5)
    thetas = np.array([0.0, 0.3, 0.9, 1.3, 1.55])
    for trial in range(12):
        ns, d = random_stack(rng, int(rng.integers(2, 40)), n0=rng.choice([1.0, 1.5, 2.5]), nf=rng.choice([1.0, 1.5]))
        R, T, A, _ = run_emul(emul, wl, [np.full(wl.shape, n) for n in ns], d, thetas, flags=SNELL)
        Ro, To = oracle_grid(wl, ns, d, thetas)
        assert close(R, Ro, 1e-11).all(), trial
        assert close(T, To, 1e-11).all(), trial
        assert np.array_equal(A, (1.0 - R) - T)


def test_emul_snell_normal_incidence_equals_default_mode(emul):  # noqa: F811
    rng = np.random.default_rng(8)
    wl = np.linspace(0.6, 1.4, 9)
    ns, d = random_stack(rng, 12)
    n_layers = [np.full(wl.shape, n) for n in ns]
    R0, T0, _, _ = run_emul(emul, wl, n_layers, d, [0.0])
    R1, T1, _, _ = run_emul(emul, wl, n_layers, d, [0.0], flags=SNELL)
    assert close(R1, R0, 1e-13).all() and close(T1, T0, 1e-13).all()


def test_emul_snell_total_internal_reflection_and_tunnelling(emul):  # noqa: F811
    wl = np.array([0.633])
    th = np.deg2rad(np.array([30.0, 41.0, 42.0, 45.0, 60.0, 80.0]))
    # glass -> air: R = 1 beyond the critical angle (41.81 deg)
    R, T, _, _ = run_emul(emul, wl, [np.array([1.5 + 0j]), np.array([1.0 + 0j])], [0.0, 0.0], th, flags=SNELL)
    assert np.all(np.abs(R[:, 0, 2:] - 1) < 1e-14) and np.all(np.abs(T[:, 0, 2:]) < 1e-14)
    assert np.all(R[:, 0, :2] < 1)
    # frustrated TIR: a thin air gap between two prisms transmits by tunnelling, R + T = 1
    ns = [1.5 + 0j, 1.0 + 0j, 1.5 + 0j]
    R, T, _, _ = run_emul(emul, wl, [np.array([n]) for n in ns], [0.0, 0.1, 0.0], th, flags=SNELL)
    Ro, To = oracle_grid(wl, ns, [0.0, 0.1, 0.0], th)
    assert close(R, Ro, 1e-12).all() and close(T, To, 1e-12).all()
    assert np.all(np.abs(R + T - 1) < 1e-13) and np.all(T[:, 0, 3:] > 1e-4)


# ------------------------------------------------------------------------------------------ CUDA path
@pytest.mark.gpu
def test_gpu_snell_matches_oracle():
    from pistachio_b200 import engine

    eng = engine.get_engine()
    rng = np.random.default_rng(9)
    wl = np.linspace(0.6, 1.4, 6)
    thetas = np.array([0.0, 0.2, 0.7, 1.2, 1.5])
    for trial in range(16):
        ns, d = random_stack(rng, int(rng.integers(2, 70)), n0=rng.choice([1.0, 1.5, 2.5]), nf=rng.choice([1.0, 1.5]))
        mat_n = np.array([np.full(wl.shape, n.real) for n in ns])
        mat_k = np.array([np.full(wl.shape, n.imag) for n in ns])
        res = eng.eval_grid(wl, thetas, mat_n, mat_k, list(range(len(ns))), d, cos_theta=np.cos(thetas),
                            flags=engine.FLAG_SNELL).numpy()
        Ro, To = oracle_grid(wl, ns, d, thetas)
        assert close(res["R"][0], Ro, 1e-11).all(), trial
        assert close(res["T"][0], To, 1e-11).all(), trial
        assert np.array_equal(res["A"][0], (1.0 - res["R"][0]) - res["T"][0])


@pytest.mark.gpu
def test_gpu_snell_matrix_sweep_and_host_path():
    from pistachio_b200 import engine

    eng = engine.get_engine()
    rng = np.random.default_rng(10)
    wl = np.linspace(0.7, 1.1, 33)
    thetas = np.linspace(0.0, 1.4, 40)
    ns, d = random_stack(rng, 7, lossless=True)
    mat_n = np.array([np.full(wl.shape, n.real) for n in ns])
    mat_k = np.zeros_like(mat_n)
    lm = list(range(len(ns)))
    sweep = np.array([0.11, 0.47, 0.93])
    res = eng.eval_grid(wl, thetas, mat_n, mat_k, lm, d, cos_theta=np.cos(thetas), sweep_layer=3, sweep_d=sweep,
                        want_matrix=True, flags=engine.FLAG_SNELL).numpy()
    assert np.all(np.abs(res["R"] + res["T"] - 1) < 1e-11)  # lossless: energy conservation at every angle
    for i, ds in enumerate(sweep):
        d2 = list(d)
        d2[3] = ds
        for ip, pol in enumerate(POLS):
            for il in (0, 17):
                for it in (0, 13, 39):
                    T, R, M = so.point(wl[il], ns, d2, thetas[it], pol)
                    assert close(res["R"][i, ip, il, it], R, 1e-11) and close(res["T"][i, ip, il, it], T, 1e-11)
                    assert np.allclose(res["M"][i, ip, il, it], M, rtol=1e-10, atol=1e-12)
    host = eng.eval_grid_host(wl, thetas, mat_n, mat_k, lm, d, cos_theta=np.cos(thetas), sweep_layer=3, sweep_d=sweep,
                              flags=engine.FLAG_SNELL)
    for k in ("R", "T", "A"):
        assert np.array_equal(np.asarray(host[k]), res[k])
    # normal incidence: refraction mode equals the reference's model
    a = eng.eval_grid(wl, [0.0], mat_n, mat_k, lm, d).numpy()
    b = eng.eval_grid(wl, [0.0], mat_n, mat_k, lm, d, flags=engine.FLAG_SNELL).numpy()
    assert close(b["R"], a["R"], 1e-13).all() and close(b["T"], a["T"], 1e-13).all()


@pytest.mark.gpu
def test_gpu_structure_snell_attribute():
    from pistachio_b200 import transfer_matrix as tm

    s = tm.Structure()
    for mat, d in (("Air", 0.0), ("ZnS", 0.12), ("MgF2", 0.2), ("ZnS", 0.12), ("CaF2", 0.0)):
        layer = tm.Layer(material=mat, thickness=d)
        layer.get_data_from_csv(mat + ".csv") if mat != "Air" else layer.set_complex_refractive([1.0], [0.0])
        if mat == "Air":
            layer.wavelengths, layer.refractive_index, layer.extinction_coeff = [1.0], [1.0], [0.0]
        s.add_layer(layer)
    s.initialize_struct(0.0, 70.0, 8, 0.8, 1.2, 64, tm.S_WAVE)
    ref = s.spectrum_grid()
    s.snell = True
    out = s.spectrum_grid()
    assert np.allclose(out["R"][:, :, 0], ref["R"][:, :, 0], rtol=1e-12, atol=1e-15)  # theta = 0
    assert not np.allclose(out["R"][:, :, -1], ref["R"][:, :, -1], rtol=1e-3)           # 70 degrees: models differ
    n = [np.asarray(layer.complex_refractive) for layer in s.layers]
    d = [layer.thickness for layer in s.layers]
    for ip, pol in enumerate(POLS):
        for il in (0, 31, 63):
            T, R, _ = so.point(s.wavelengths[il], [x[il] for x in n], d, s.theta[-1], pol)
            assert close(out["R"][ip, il, -1], R, 1e-11) and close(out["T"][ip, il, -1], T, 1e-11)


# ------------------------------------------------------------------------------------------ per-layer field amplitudes
def test_oracle_field_amplitudes_properties():
    """Layer 0 holds (1, r), the exit medium (t, 0) with R = |r|^2 and T = Re(det M)|t|^2 of the spectrum oracle; in a
    lossless stack the s-wave energy flux Re(n cos)(|A|^2 - |B|^2) is the same in every layer."""
    rng = np.random.default_rng(11)
    for snell in (True, False):
        ns, d = random_stack(rng, 9, lossless=True)
        lam, th = 0.9, 0.7
        for pol in POLS:
            amp = so.field_amplitudes_point(lam, ns, d, th, pol, snell=snell)
            if snell:
                T, R, _ = so.point(lam, ns, d, th, pol)
            else:
                Tg, Rg = to.spectrum(to.Stack([lam], [np.array([n]) for n in ns], d), th, pol)
                T, R = Tg[0], Rg[0]
            assert abs(amp[0, 0] - 1) < 1e-15 and abs(abs(amp[0, 1]) ** 2 - R) < 1e-12 and amp[-1, 1] == 0
            cs = so.layer_cosines(ns, ns[0], th) if snell else [np.cos(th)] * len(ns)
            det = (ns[-1] * cs[-1]) / (ns[0] * cs[0])
            assert abs(det.real * abs(amp[-1, 0]) ** 2 - T) < 1e-12
        if snell:
            amp = so.field_amplitudes_point(lam, ns, d, th, so.S_WAVE)
            cs = so.layer_cosines(ns, ns[0], th)
            flux = [(n * c).real * (abs(a) ** 2 - abs(b) ** 2) for n, c, (a, b) in zip(ns, cs, amp)]
            assert np.allclose(flux, flux[0], rtol=1e-11)


@pytest.mark.gpu
def test_gpu_field_amplitudes_match_oracle():
    from pistachio_b200 import engine

    eng = engine.get_engine()
    rng = np.random.default_rng(12)
    wl = np.linspace(0.6, 1.4, 5)
    thetas = np.array([0.0, 0.5, 1.2, 1.5])
    for trial in range(8):
        ns, d = random_stack(rng, int(rng.integers(2, 30)), n0=rng.choice([1.0, 1.5]), nf=rng.choice([1.0, 1.5]))
        mat_n = np.array([np.full(wl.shape, n.real) for n in ns])
        mat_k = np.array([np.full(wl.shape, n.imag) for n in ns])
        for snell in (False, True):
            amp = eng.field_amplitudes(wl, thetas, mat_n, mat_k, list(range(len(ns))), d, cos_theta=np.cos(thetas),
                                       flags=engine.FLAG_SNELL if snell else 0).cpu().numpy()
            assert amp.shape == (2, len(wl), len(thetas), len(ns), 2)
            for ip, pol in enumerate(POLS):
                for il in (0, 4):
                    for it in range(len(thetas)):
                        ref = so.field_amplitudes_point(wl[il], ns, d, thetas[it], pol, snell=snell)
                        scale = np.maximum(np.abs(ref).max(axis=1, keepdims=True), 1e-6)
                        assert np.all(np.abs(amp[ip, il, it] - ref) <= 1e-10 * scale), (trial, snell, pol, il, it)
    # a deep stop band: 1500 quarter-wave pairs (|t| ~ 1e-348); the unnormalised chain leaves the fp64 range, the amplitudes do not
    wl = np.array([1.0])
    mat_n = np.array([[1.0], [2.32], [1.36], [1.5]])
    lm = [0] + [1, 2] * 1500 + [3]
    d = [0.0] + [1.0 / (4 * 2.32), 1.0 / (4 * 1.36)] * 1500 + [0.0]
    amp = eng.field_amplitudes(wl, [0.0], mat_n, np.zeros_like(mat_n), lm, d, ("s-wave",)).cpu().numpy()[0, 0, 0]
    assert np.isfinite(amp.view(np.float64)).all() and abs(abs(amp[0, 1]) - 1) < 1e-12
    assert abs(amp[-1, 0]) < 1e-300 and 1e-240 < abs(amp[2001, 0]) < 1e-220 and np.all(np.abs(amp[1:, 0]) <= 2.0)
