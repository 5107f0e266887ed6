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
