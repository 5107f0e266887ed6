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
    out = s.polariton_dispersion(theta=w.theta, height=0.002, distance=10)
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
