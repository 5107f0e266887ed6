"""Fused spectral reduction: peak search with scipy.signal.find_peaks semantics.
CPU: the oracle restatement against scipy itself.  GPU: the kernel against scipy, index for index."""
import numpy as np
import pytest
from scipy import signal

from oracle import peaks_oracle as po


def random_spectra(rng, n_spec, n_pts, plateaus=False):
    if plateaus:
        y = rng.integers(0, 6, (n_spec, n_pts)).astype(np.float64)
        y += np.linspace(0, 1e-3, n_pts)[None, :] * rng.integers(0, 2, (n_spec, 1))  # some rows without exact ties
        return y
    y = np.cumsum(rng.normal(size=(n_spec, n_pts + 6)), axis=1)
    y = sum(y[:, k:k + n_pts] for k in range(7)) / 7.0  # smoothed: a few dozen maxima per spectrum
    return y + 3 * np.sin(np.linspace(0, 40, n_pts))[None, :]


def test_oracle_matches_scipy():
    rng = np.random.default_rng(0)
    for t in range(200):
        n = int(rng.integers(3, 400))
        x = random_spectra(rng, 1, n, plateaus=bool(t % 2))[0]
        h = None if t % 3 == 0 else float(np.median(x))
        d = None if t % 4 == 0 else int(rng.integers(1, 30))
        assert np.array_equal(po.find_peaks(x, height=h, distance=d), signal.find_peaks(x, height=h, distance=d)[0])
    assert len(po.find_peaks([1.0, 2.0], None, None)) == 0 and len(po.find_peaks([], None, None)) == 0
    assert np.isnan(po.mean_spacing(np.arange(5.0), np.array([2])))
    assert po.mean_spacing(np.array([0.0, 1.0, 4.0, 9.0]), np.array([0, 2, 3])) == 4.5


@pytest.mark.gpu
@pytest.mark.parametrize("plateaus", [False, True])
def test_kernel_matches_scipy_index_for_index(plateaus):
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from pistachio_b200.engine import get_engine

    eng = get_engine(0)
    rng = np.random.default_rng(5 + plateaus)
    n_outer, n_pts, n_inner = 3, (257 if plateaus else 517), 37  # integer plateaus make ~90 maxima in 257 samples (cap 128)
    y = random_spectra(rng, n_outer * n_inner, n_pts, plateaus).reshape(n_outer, n_inner, n_pts).transpose(0, 2, 1).copy()
    if plateaus:  # equal heights make scipy's visiting order unspecified (unstable argsort): break ties
        y += rng.uniform(0, 1e-9, (n_outer, 1, n_inner)) * 0 + np.arange(n_pts)[None, :, None] * 1e-12 * (np.arange(n_inner)[None, None, :] % 2)
    y_d = torch.from_numpy(y).cuda()
    axis = np.linspace(5.0, 9.0, n_pts) ** 2
    for height, distance, reverse in ((None, None, False), (0.5, None, False), (None, 7, False), (1.0, 25, True), (None, 1, True)):
        idx, hts, cnt, spc = eng.find_peaks(y_d, height=height, distance=distance, max_peaks=128, axis=axis, reverse=reverse)
        idx, hts, cnt, spc = idx.cpu().numpy(), hts.cpu().numpy(), cnt.cpu().numpy(), spc.cpu().numpy()
        for o in range(n_outer):
            for i in range(n_inner):
                x = y[o, :, i][::-1] if reverse else y[o, :, i]
                if plateaus and distance is not None and len(np.unique(x[signal.find_peaks(x)[0]])) != len(signal.find_peaks(x)[0]):
                    continue  # tied heights: visiting order is implementation-defined in scipy
                ref = signal.find_peaks(x, height=height, distance=distance)[0]
                assert cnt[o, i] == len(ref), (height, distance, reverse, o, i)
                assert np.array_equal(idx[o, i, :cnt[o, i]], ref)
                assert np.array_equal(hts[o, i, :cnt[o, i]], x[ref])
                ax = axis[::-1] if reverse else axis
                exp = po.mean_spacing(ax, ref)
                assert (np.isnan(exp) and np.isnan(spc[o, i])) or abs(spc[o, i] - exp) <= 1e-12 * abs(exp)
    # overflow is reported, not silently truncated
    idx, hts, cnt, spc = eng.find_peaks(y_d, max_peaks=4)
    assert (cnt.cpu().numpy() == -1).any()


@pytest.mark.gpu
def test_rippled_spectra_with_a_large_distance_are_not_an_overflow():
    """Hundreds of local maxima per spectrum (more than the kernel's candidate buffer) thinned by a large `distance` to a
    handful: scipy applies the distance to all candidates, so max_peaks bounds only what is reported (the kernel's
    re-scan path); without a distance condition the same spectra do overflow."""
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from pistachio_b200.engine import get_engine

    eng = get_engine(0)
    rng = np.random.default_rng(11)
    n_outer, n_pts, n_inner = 2, 2311, 33
    y = rng.normal(size=(n_outer, n_pts, n_inner)) + 2.0 * np.sin(np.linspace(0, 23, n_pts))[None, :, None]
    y[1, :, ::2] = np.cumsum(rng.normal(size=(n_pts, (n_inner + 1) // 2)), axis=0) / 7.0  # smooth rows: the buffered path
    y_d = torch.from_numpy(y).cuda()
    for height, distance, max_peaks, reverse in ((None, 117, 64, False), (0.5, 60, 64, True), (None, 400, 8, False), (None, 9, 128, True)):
        idx, hts, cnt, _ = eng.find_peaks(y_d, height=height, distance=distance, max_peaks=max_peaks, reverse=reverse)
        idx, hts, cnt = idx.cpu().numpy(), hts.cpu().numpy(), cnt.cpu().numpy()
        n_slow = 0
        for o in range(n_outer):
            for i in range(n_inner):
                x = y[o, :, i][::-1] if reverse else y[o, :, i]
                ref = signal.find_peaks(x, height=height, distance=distance)[0]
                n_slow += len(signal.find_peaks(x, height=height)[0]) > 128
                if len(ref) > max_peaks:
                    assert cnt[o, i] == -1
                    continue
                assert cnt[o, i] == len(ref), (height, distance, reverse, o, i, cnt[o, i], len(ref))
                assert np.array_equal(idx[o, i, :cnt[o, i]], ref)
                assert np.array_equal(hts[o, i, :cnt[o, i]], x[ref])
        assert n_slow >= n_inner  # the case is exercised
    _, _, cnt, _ = eng.find_peaks(y_d, max_peaks=128)
    assert (cnt.cpu().numpy()[0] == -1).all()


@pytest.mark.gpu
def test_cavity_length_from_fused_peaks():
    """Fabry-Perot etalon with a 20 um air gap: free spectral range 1e4/(2 * 20) = 250 cm^-1, so the notebook's formula
    must return ~20 um; peak lists agree with scipy on the downloaded spectrum."""
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import importlib.resources as pkg_resources

    from pistachio_b200 import default
    from pistachio_b200 import transfer_matrix as tm

    s = tm.Structure()
    s.load_struct_from_config(str(pkg_resources.files(default) / "fabry-perot.yaml"))
    out = s.cavity_length_from_peaks(n_eff=1.000273, theta=[0.0], height=1.0, distance=50)
    T = s.spectrum_grid([0.0], ("s-wave",), want=("T",))["T"][0, :, 0]
    ref = signal.find_peaks(T[::-1] * 100.0, height=1.0, distance=50)[0]
    nu = (1.0e4 / s.wavelengths)[::-1]
    assert out["count"][0] == len(ref) and len(ref) > 5
    np.testing.assert_array_equal(out["peak_wavenumber"][0], nu[ref])
    assert abs(out["mean_spacing"][0] - 250.0) < 5.0
    assert abs(out["L_cav"][0] - 20.0) < 0.5
    # thickness sweep: the fitted length follows the swept gap
    sw = np.array([10.0, 15.0, 25.0])
    out = s.cavity_length_from_peaks(n_eff=1.000273, theta=[0.0], height=1.0, distance=20, sweep_layer=2, sweep_d=sw)
    np.testing.assert_allclose(out["L_cav"][:, 0], sw, rtol=0.03)


def test_three_point_lorentzian_is_exact_on_a_nonuniform_axis():
    x = 1.0e4 / np.linspace(3.5, 4.5, 400)[::-1]  # wavenumbers of a uniform wavelength grid: non-uniform
    x0, g, A = 2531.7, 11.3, 0.83
    y = A / (1.0 + ((x - x0) / g) ** 2)
    i = int(np.argmax(y))
    c, w, a = po.lorentzian_three_point(x, y, i)
    assert abs(c - x0) < 1e-9 * x0 and abs(w - g) < 1e-7 * g and abs(a - A) < 1e-9 * A
    # two well separated lines: centres within a small fraction of the sample spacing, order lower/upper
    y2 = y + 0.5 / (1.0 + ((x - 2700.0) / 8.0) ** 2)
    lo, hi, *_ = po.polariton_branches(x, y2, height=0.1)
    assert abs(lo - x0) < 0.05 and abs(hi - 2700.0) < 0.05
    assert all(np.isnan(v) for v in po.polariton_branches(x, y))  # a single peak: no branches


@pytest.mark.gpu
def test_polariton_branches_kernel_matches_oracle():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from pistachio_b200.engine import get_engine

    eng = get_engine(0)
    rng = np.random.default_rng(11)
    n_outer, n_pts, n_inner = 2, 700, 19
    wl = np.linspace(3.5, 4.5, n_pts)
    nu = 1.0e4 / wl
    y = np.empty((n_outer, n_pts, n_inner))
    for o in range(n_outer):
        for i in range(n_inner):
            c1, c2 = 2400 + 60 * rng.random(), 2600 + 60 * rng.random()
            y[o, :, i] = (0.2 + 0.6 * rng.random()) / (1 + ((nu - c1) / (5 + 10 * rng.random())) ** 2) \
                + (0.2 + 0.6 * rng.random()) / (1 + ((nu - c2) / (5 + 10 * rng.random())) ** 2) + 0.01
    y[1, :, 3] = 0.01 + 0.5 / (1 + ((nu - 2500.0) / 9.0) ** 2)  # one line only -> NaN
    y_d = torch.from_numpy(y).cuda()
    idx, hts, cnt, _ = eng.find_peaks(y_d, height=0.05, distance=5, max_peaks=16, axis=nu, reverse=True)
    got = eng.polariton_branches(y_d, idx, hts, cnt, nu, reverse=True).cpu().numpy()
    for o in range(n_outer):
        for i in range(n_inner):
            ref = np.array(po.polariton_branches(nu[::-1], y[o, ::-1, i], height=0.05, distance=5))
            if np.isnan(ref[0]):
                assert np.isnan(got[o, i]).all()
                continue
            np.testing.assert_allclose(got[o, i], ref, rtol=1e-9, atol=0)
            assert got[o, i, 0] < got[o, i, 1]


@pytest.mark.gpu
def test_polariton_dispersion_of_a_strongly_coupled_cavity():
    """Au / Lorentz absorber / Au cavity tuned to the 4 um line (the C3 stack): two transmission branches that anticross;
    the upper one moves up with angle, the splitting is positive and smallest near resonance."""
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from pistachio_b200 import transfer_matrix as tm
    from pistachio_b200 import workloads as wk

    w = wk.c3_polariton(n_d=1, n_wl=4096, n_theta=16)
    s = tm.Structure()
    for spec in w.layers:
        l = tm.Layer(material=spec.material, thickness=spec.thickness)
        l.wavelengths, l.refractive_index, l.extinction_coeff = np.array(spec.tab_wl), np.array(spec.tab_n), np.array(spec.tab_k)
        l.set_complex_refractive(l.refractive_index, l.extinction_coeff)
        s.add_layer(l)
    s.layers[2].thickness = 4.0 / (2 * 1.5) * 3  # third-order cavity mode on the 4 um line
    s.wavelengths = w.wl
    for l in s.layers:
        l.make_datapoints(w.wl)
    out = s.polariton_dispersion(theta=w.theta, height=0.002, distance=10, fit="three-point")
    ok = ~np.isnan(out["splitting"])
    assert ok.sum() >= 12
    lp, up, spl = out["LP"][ok], out["UP"][ok], out["splitting"][ok]
    assert (spl > 0).all()
    assert up[-1] > up[0] + 100.0   # the photon-like branch blue-shifts with angle ...
    assert lp[-1] > lp[0] and lp.max() < 2520.0  # ... the lower one saturates at the 2500 cm^-1 vibration
    k = int(np.argmin(spl))
    assert 0 < k < len(spl) - 1 and 20.0 < out["min_splitting"] < 80.0  # anticrossing at an interior angle
    # the same numbers from the CPU oracle's spectrum at three angles
    from oracle import tmm_oracle as orc

    n_list = [np.asarray(l.complex_refractive) for l in s.layers]
    stack = orc.Stack(w.wl, n_list, [l.thickness for l in s.layers])
    nu = 1.0e4 / w.wl
    for it in (0, 7, 15):
        T, _ = orc.spectrum(stack, float(w.theta[it]), orc.S_WAVE)
        ref = po.polariton_branches(nu[::-1], T[::-1], height=0.002, distance=10)
        np.testing.assert_allclose([out["LP"][it], out["UP"][it]], ref[:2], rtol=1e-9)
    # the default: the notebook's joint two-Lorentzian fit over a wavenumber window, stopped at lmfit's tolerance, against
    # scipy's curve_fit (MINPACK, the minimiser lmfit calls) on the oracle's spectrum with the same seeds, window and
    # tolerance.  These spectra are NOT two Lorentzians (Fano-like line shapes, a third feature at large angles): the
    # misfit surface has flat valleys and several minima, so the pin is the misfit reached (never worse than MINPACK's)
    # and, at angles where both lines sit well inside the window, the centres MINPACK stops at -- to what a 1.5e-8
    # stopping rule defines.  The tight pin of the fit itself is on spectra that are two Lorentzians (below).
    win = (2350.0, 2700.0)
    fit = s.polariton_dispersion(theta=w.theta, height=0.002, distance=10, window=win)
    assert (fit["evaluations"][ok] > 0).sum() >= 12 and np.isfinite(fit["LP"][ok]).all()
    order = nu[::-1]
    lo, hi = np.searchsorted(order, win)
    for it in (0, 3, 7, 9, 11):
        T, _ = orc.spectrum(stack, float(w.theta[it]), orc.S_WAVE)
        seed = po.polariton_branches(order, T[::-1], height=0.002, distance=10)
        ref = po.fit_two_lorentzians(order[lo:hi], T[::-1][lo:hi], seed, tol=1.5e-8)
        assert fit["evaluations"][it] > 0
        np.testing.assert_allclose([fit["LP"][it], fit["UP"][it]], ref[:2], rtol=5e-5)
        np.testing.assert_allclose([fit["hwhm_LP"][it], fit["hwhm_UP"][it]], ref[2:4], rtol=2e-2)
        assert fit["chi2"][it] <= ref[6] * (1 + 1e-4)
    assert (fit["splitting"][[0, 3, 7, 9, 11]] > 0).all()


# --------------------------------------------------------------------------------------- joint two-Lorentzian fit
def two_peak_cases():
    """Synthetic two-branch spectra on a NON-uniform axis (wavenumbers of a uniform wavelength grid, ascending): well
    separated branches, overlapping branches near an anticrossing (separation ~ width: where a per-peak estimate is
    biased by the neighbour's tail), unequal heights, and a noisy one."""
    nu = (1.0e4 / np.linspace(5.5, 4.2, 700))  # ~1818 .. 2381 cm^-1, ascending
    rng = np.random.default_rng(3)
    cases = []
    for name, (a1, c1, s1, a2, c2, s2), noise in (
            ("separated", (40.0, 1950.0, 9.0, 55.0, 2230.0, 12.0), 0.0),
            ("overlapping", (30.0, 2080.0, 14.0, 30.0, 2112.0, 14.0), 0.0),
            ("shoulder", (60.0, 2090.0, 10.0, 18.0, 2118.0, 16.0), 0.0),
            ("noisy", (40.0, 2020.0, 8.0, 35.0, 2140.0, 10.0), 2e-3)):
        y = po.two_lorentzians(nu, a1, c1, s1, a2, c2, s2) + noise * rng.normal(size=nu.size)
        cases.append((name, nu, y, (a1, c1, s1, a2, c2, s2)))
    return cases


def seed_of(nu, y):
    """What the device pipeline feeds the fit: the three-point estimates of the two highest peaks."""
    row = po.polariton_branches(nu, y, height=None, distance=3)
    if not np.isfinite(row[0]):
        i = int(np.argmax(y))  # a shoulder is not a local maximum: seed both lines around the one peak
        c, g, h = po.lorentzian_three_point(nu, y, i)
        row = (c - g, c + g, g, g, h / 2, h / 2)
    return np.array(row, dtype=np.float64)


def test_fit_oracle_recovers_the_generating_parameters():
    for name, nu, y, truth in two_peak_cases():
        if name == "noisy":
            continue
        got = po.fit_two_lorentzians(nu, y, seed_of(nu, y))
        a1, c1, s1, a2, c2, s2 = truth
        np.testing.assert_allclose(got[:6], (c1, c2, s1, s2, a1, a2), rtol=1e-9, err_msg=name)
        assert got[6] < 1e-20


def test_device_fit_source_matches_scipy_curve_fit():
    """pst_fit.cuh through the host emulation against scipy's Levenberg-Marquardt on the same data and seeds."""
    import ctypes as C

    from tests import test_host_emul as he

    lib = he.load_emul()
    if lib is None:
        pytest.skip("host emulation unavailable (no nvcc)")
    for name, nu, y, truth in two_peak_cases():
        seed = seed_of(nu, y)
        ref = po.fit_two_lorentzians(nu, y, seed)
        out = np.empty(8)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        x = np.ascontiguousarray(nu)
        yy = np.ascontiguousarray(y)
        lib.emul_fit_two_lorentzians(len(x), p(x), p(yy), p(seed), C.c_double(5.0), C.c_double(1e-15), 400, p(out))
        assert out[7] > 0, (name, out)
        np.testing.assert_allclose(out[:4], ref[:4], rtol=1e-8, err_msg=name)
        np.testing.assert_allclose(out[4:6], ref[4:6], rtol=1e-7, err_msg=name)
        assert out[6] <= ref[6] * (1 + 1e-9) + 1e-25, (name, out[6], ref[6])
    # the joint fit against the per-peak estimate on overlapping branches: the three-point centres are pulled together by
    # the neighbour's tail, the joint fit returns the true splitting
    name, nu, y, truth = two_peak_cases()[1]
    seed = seed_of(nu, y)
    out = np.empty(8)
    lib.emul_fit_two_lorentzians(len(nu), p(np.ascontiguousarray(nu)), p(np.ascontiguousarray(y)), p(seed), C.c_double(5.0),
                                 C.c_double(1e-15), 400, p(out))
    true_split = truth[4] - truth[1]
    assert abs((out[1] - out[0]) - true_split) < 1e-7
    assert abs((seed[1] - seed[0]) - true_split) > 1.0  # cm^-1: the bias the joint fit removes


@pytest.mark.gpu
def test_fit_kernel_matches_scipy_curve_fit():
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from pistachio_b200.engine import get_engine

    eng = get_engine(0)
    cases = two_peak_cases()
    nu = cases[0][1]
    # spectra laid out like a grid result [outer][wavelength][theta] on a DEscending-wavenumber (ascending wavelength)
    # sample order, scanned in reverse: what Structure.polariton_dispersion passes
    Y = np.stack([c[2][::-1] for c in cases], axis=1)[None]  # (1, n_pts, 4)
    y_d = torch.from_numpy(np.ascontiguousarray(Y)).cuda()
    axis = nu[::-1].copy()
    idx, hts, cnt, _ = eng.find_peaks(y_d, distance=3, max_peaks=64, axis=axis, reverse=True)
    seeds = eng.polariton_branches(y_d, idx, hts, cnt, axis, reverse=True)
    seeds_h = seeds.cpu().numpy()
    for i, c in enumerate(cases):
        if not np.isfinite(seeds_h[0, i, 0]):
            seeds_h[0, i] = seed_of(c[1], c[2])
    seeds = torch.from_numpy(seeds_h).cuda()
    out = eng.fit_two_lorentzians(y_d, seeds, axis, reverse=True, max_iter=400, tol=1e-15).cpu().numpy()
    for i, (name, nu_i, y_i, truth) in enumerate(cases):
        ref = po.fit_two_lorentzians(nu_i, y_i, seeds_h[0, i])
        assert out[0, i, 7] > 0, (name, out[0, i])
        np.testing.assert_allclose(out[0, i, :4], ref[:4], rtol=1e-8, err_msg=name)
        np.testing.assert_allclose(out[0, i, 4:6], ref[4:6], rtol=1e-7, err_msg=name)
    # a window (the notebook's wavenumber mask): only the samples inside it enter the fit
    lo, hi = 150, 620
    outw = eng.fit_two_lorentzians(y_d, seeds, axis, reverse=True, window=(lo, hi), max_iter=400).cpu().numpy()
    name, nu_i, y_i, _ = cases[3]
    refw = po.fit_two_lorentzians(nu_i[lo:hi], y_i[lo:hi], seeds_h[0, 3])
    np.testing.assert_allclose(outw[0, 3, :4], refw[:4], rtol=1e-8)
