"""The reference's own tests (reference tests/test_layer.py, tests/test_structure.py) and its tutorial, re-run
against the drop-in module pistachio_b200.transfer_matrix on the GPU, plus parity of every API-level method."""
import importlib.resources as pkg_resources
import os

import numpy as np
import pytest
from numpy import testing

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import tmm_oracle as orc  # noqa: E402
from tests import parity  # noqa: E402


@pytest.fixture(scope="module")
def tm():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from pistachio_b200 import transfer_matrix

    return transfer_matrix


def default_yaml(name):
    from pistachio_b200 import default

    return str(pkg_resources.files(default) / name)


# ---------------------------------------------------------------- reference tests/test_layer.py
def test_make_datapoints(tm):
    layer = tm.Layer(material="Air")
    layer.refractive_index = [1.0]
    layer.extinction_coeff = [0.0]
    layer.make_datapoints(np.linspace(1.0, 10.0, 3))
    assert len(layer.wavelengths) == 3 and len(layer.refractive_index) == 3
    assert layer.complex_refractive.dtype == np.complex128


def test_make_datapoints_interpolation(tm, ref_cases):
    layer = tm.Layer()
    layer.get_data_from_csv("Au.csv")
    old_w, old_n, old_k = layer.wavelengths, layer.refractive_index, layer.extinction_coeff
    new_w = np.linspace(0.2, 15.0, 10000)
    layer.make_datapoints(new_w)
    testing.assert_almost_equal(old_w[0], 0.19077)
    testing.assert_almost_equal(old_w[264], 1.0164)
    for target, tol_n, tol_k, atol in ((0.2, 3, 2, 1e-8), (1.0, 2, 1, 1e-2), (2.0, 2, 1, 1e-2)):
        i_old = np.where(np.isclose(old_w, target))[0][0]
        i_new = np.where(np.isclose(new_w, target, atol=atol))[0][0]
        testing.assert_almost_equal(layer.refractive_index[i_new], old_n[i_old], decimal=tol_n)
        testing.assert_almost_equal(layer.extinction_coeff[i_new], old_k[i_old], decimal=tol_k)
    c = ref_cases["interp_au"]
    assert np.array_equal(layer.refractive_index, c["n"]) and np.array_equal(layer.extinction_coeff, c["k"])


def test_dynamical_matrix(tm, ref_cases):
    A = tm.Layer().dynamical_matrix(1.0 + 1.0j, 0.0, "s-wave")
    B = tm.Layer().dynamical_matrix(1.0 + 1.0j, 0.0, "p-wave")
    Cm = np.array([[1.0, 1.0], [1.0 + 1.0j, -1.0 - 1.0j]])
    testing.assert_allclose(A, B)
    testing.assert_allclose(A, Cm)
    assert np.array_equal(A, ref_cases["kat"]["D_s"])
    assert tm.Layer().dynamical_matrix(1.0, 0.0, "x-wave") is None  # prints the reference's message


def test_propagation_matrix(tm, ref_cases):
    kx = 2.0e5 + 0.0j
    d = np.pi / (2 * kx)
    P = tm.Layer().propagation_matrix(kx, d)
    testing.assert_allclose(P, np.array([[-1.0j, 0.0], [0.0, 1.0j]]), atol=1e-15)
    testing.assert_allclose(P, ref_cases["kat"]["P"], rtol=1e-15, atol=2e-16)


def test_transfer_matrix(tm, ref_cases):
    layer = tm.Layer(material="Air", thickness=3.0, num_points=1)
    out = layer.calculate_transfer_matrix(1.0 + 0.0j, 1.0, 0.0, "s-wave")
    assert len(out) == 4
    testing.assert_allclose(out[-1], np.identity(2), atol=1e-4)
    testing.assert_allclose(np.array(out), ref_cases["kat"]["M_air"], rtol=1e-14, atol=1e-15)
    lay2 = tm.Layer(material="x", thickness=0.37)
    for pol, key in (("s-wave", "M4_s"), ("p-wave", "M4_p")):
        got = np.array(lay2.calculate_transfer_matrix(1.7 + 0.02j, 0.83, 0.6, pol))
        testing.assert_allclose(got, ref_cases["kat"][key], rtol=2e-15, atol=1e-16)
    with pytest.raises(np.linalg.LinAlgError):
        lay2.calculate_transfer_matrix(1.0, 1.0, 0.0, "x-wave")


# ---------------------------------------------------------------- reference tests/test_structure.py
def test_load_struct_from_yaml(tm, ref_cases):
    s = tm.Structure()
    s.load_struct_from_config(default_yaml("Au_thin_layer.yaml"))
    assert len(s.layers) == 3 and len(s.wavelengths) == 100
    for layer in s.layers:
        assert len(layer.wavelengths) == 100
    testing.assert_array_less(s.theta, np.full(len(s.theta), np.pi / 2))
    c = ref_cases["yaml_au_thin"]
    assert np.array_equal(np.array([l.complex_refractive for l in s.layers]), c["n"])
    M = s.calculate_all_transfer_matrices(0.0, "s-wave")
    T, R = s.calculate_all_t_r(M)
    testing.assert_allclose(T, c["T"], rtol=1e-12)
    testing.assert_allclose(R, c["R"], rtol=1e-12)
    scale = np.abs(c["M"]).max(axis=(-1, -2), keepdims=True)
    assert np.all(np.abs(M - c["M"]) <= 1e-13 * scale)
    # the lazily built per-layer cache has the reference's shape and values
    cache = s.layers[1].transfer_matrices
    assert cache.shape == (100, 4, 2, 2)
    ref = orc.layer_matrices_all(c["n"][1], c["wl"], 0.0, "s-wave", c["d"][1])
    testing.assert_allclose(cache, ref, rtol=1e-13, atol=1e-15)
    kx, kz = orc.wavenumbers(c["n"][1], c["wl"], 0.0)
    assert np.array_equal(s.layers[1].kx, kx) and np.array_equal(s.layers[1].kz, kz)


def test_calculate_t_r(tm):
    T, R = tm.Structure().calculate_t_r(np.array([[1.0 + 0.0j, 0.0], [0.0, 1.0 + 0.0j]]))
    assert T == 1.0 and R == 0.0


def test_quarter_wave_yaml(tm, ref_cases):
    s = tm.Structure()
    s.load_struct_from_config(default_yaml("quarter-wave.yaml"))
    c = ref_cases["yaml_quarter_wave"]
    T, R = s.calculate_all_t_r(s.calculate_all_transfer_matrices(0.0, "s-wave"))
    testing.assert_allclose(T, c["T"], rtol=1e-12)
    testing.assert_allclose(R, c["R"], rtol=1e-12)


# ---------------------------------------------------------------- tutorial golden vectors through the drop-in API
def test_tutorial_example1(tm, tutorial_goldens):
    wavelength = np.linspace(3.5e-6, 4.5e-6, 1000)
    lor = 1.0e-14 / (wavelength ** 2 - (4.0e-6) ** 2 + 1j * wavelength * 1.5e-8)
    refractive = np.sqrt(1.5 ** 2 - lor)
    s = tm.Structure()
    caf2 = tm.Layer(material="CaF2", thickness=0.0)
    caf2.get_data_from_csv("CaF2.csv")
    s.add_layer(caf2)
    sample = tm.Layer(material="Sample")
    sample.wavelengths = wavelength
    sample.refractive_index = refractive.real
    sample.extinction_coeff = refractive.imag
    sample.set_complex_refractive(refractive.real, refractive.imag)
    s.add_layer(sample)
    s.layers[1].thickness = 2.0e-5
    caf2 = tm.Layer(material="CaF2", thickness=0.0)
    caf2.get_data_from_csv("CaF2.csv")
    s.add_layer(caf2)
    s.initialize_struct(0.0, 0.0, 1, 3.5e-6, 4.5e-6, 5000, "s-wave")
    M_all = s.calculate_all_transfer_matrices(0.0, "s-wave")
    T_all, R_all = s.calculate_all_t_r(M_all)
    testing.assert_allclose((1 / s.wavelengths) / 100, tutorial_goldens["tutorial1/wavenumber"], rtol=1e-15)
    testing.assert_allclose(T_all * 100, tutorial_goldens["tutorial1/T_percent"], rtol=1e-12)


def test_tutorial_fabry_perot(tm, tutorial_goldens, ref_cases):
    fp = tm.Structure()
    fp.load_struct_from_config(default_yaml("fabry-perot.yaml"))
    fp.layers[3].thickness = 2e-3  # the notebook's printed structure: 'Layer 3 : Au | d = 2 nm'
    T_all, R_all = fp.calculate_all_t_r(fp.calculate_all_transfer_matrices(0.0, "s-wave"))
    testing.assert_allclose((1 / fp.wavelengths) / 100, tutorial_goldens["fabry_perot/wavenumber"], rtol=1e-15)
    testing.assert_allclose(T_all * 100, tutorial_goldens["fabry_perot/T_percent"], rtol=1e-11)
    testing.assert_allclose(R_all * 100, tutorial_goldens["fabry_perot/R_percent"], rtol=1e-12)
    c = ref_cases["yaml_fabry_perot_2nm"]
    testing.assert_allclose(T_all, c["T"], rtol=1e-11)
    testing.assert_allclose(R_all, c["R"], rtol=1e-12)


# ---------------------------------------------------------------- cached (strict reference) dataflow and sweeps
def test_strict_reference_cache_mode(tm, ref_cases):
    c = ref_cases["c2_sub"]
    s = tm.Structure()
    for n, d in zip(c["n"], c["d"]):
        layer = tm.Layer(material="m", thickness=float(d))
        layer.wavelengths = c["wl"]
        layer.refractive_index, layer.extinction_coeff = n.real.copy(), n.imag.copy()
        layer.complex_refractive = n
        s.add_layer(layer)
    s.wavelengths = c["wl"]
    s.strict_reference_cache = True
    th = float(c["theta"][2])
    s.calculate_layer_matrices(th, "p-wave")
    M = s.calculate_all_transfer_matrices(0.0, "s-wave")  # arguments ignored, exactly like tm.py:556
    T, R = s.calculate_all_t_r(M)
    ref_m = c["M"][1, :, 2]
    scale = np.abs(ref_m).max(axis=(-1, -2), keepdims=True)
    assert np.all(np.abs(M - ref_m) <= 1e-11 * scale)
    sel = dict(wl=c["wl"], n_list=list(c["n"]), d_list=list(c["d"]), thetas=c["theta"][2:3], pols=("p-wave",))
    rep = parity.compare_grid("strict cache", T[None, :, None], R[None, :, None], sel["wl"], sel["n_list"], sel["d_list"], sel["thetas"],
                              c["T"][1:2, :, 2:3], c["R"][1:2, :, 2:3], pols=sel["pols"])
    rep.check(max_loose_fraction=0.01)
    # default mode honours the arguments
    s.strict_reference_cache = False
    T2, R2 = s.calculate_all_t_r(s.calculate_all_transfer_matrices(th, "p-wave"))
    rep = parity.compare_grid("honoured arguments", T2[None, :, None], R2[None, :, None], sel["wl"], sel["n_list"], sel["d_list"],
                              sel["thetas"], c["T"][1:2, :, 2:3], c["R"][1:2, :, 2:3], pols=sel["pols"])
    rep.check(max_loose_fraction=0.01)


def test_angle_resolved_spectra_and_grid(tm, ref_cases):
    c = ref_cases["c2_sub"]
    s = tm.Structure()
    for n, d in zip(c["n"], c["d"]):
        layer = tm.Layer(material="m", thickness=float(d))
        layer.wavelengths = c["wl"]
        layer.refractive_index, layer.extinction_coeff = n.real.copy(), n.imag.copy()
        layer.complex_refractive = n
        s.add_layer(layer)
    s.wavelengths = c["wl"]
    s.theta = c["theta"]
    out = s.angle_resolved_spectra("p-wave")
    assert len(out) == len(c["theta"])
    T = np.stack([t for t, _ in out], axis=1)[None]
    R = np.stack([r for _, r in out], axis=1)[None]
    rep = parity.compare_grid("angle_resolved_spectra", T, R, c["wl"], list(c["n"]), list(c["d"]), c["theta"], c["T"][1:2], c["R"][1:2],
                              pols=("p-wave",))
    rep.check(max_loose_fraction=0.01)
    g = s.spectrum_grid()
    assert g["R"].shape == (2, len(c["wl"]), len(c["theta"]))
    rep = parity.compare_grid("spectrum_grid", g["T"], g["R"], c["wl"], list(c["n"]), list(c["d"]), c["theta"], c["T"], c["R"], got_A=g["A"])
    rep.check(max_loose_fraction=0.01)


def test_fresnel_and_interface_matrix(tm):
    n1, n2 = 1.0 + 0.0j, 1.5 + 0.1j
    k1, k2 = 6.0 + 0.0j, 9.0 + 0.6j
    got = tm.fresnel(n1, n2, k1, k2)
    ref = orc.fresnel(n1, n2, k1, k2)
    testing.assert_allclose(np.array(got), np.array(ref), rtol=1e-14)
    D = tm.transmission_matrix(got[0], got[1])
    testing.assert_allclose(D, orc.transmission_matrix(ref[0], ref[1]), rtol=1e-14)
    # interface-matrix form equals D_i^-1 D_j of the dynamical matrices, s-wave, theta = 0
    Di = orc.dynamical(n1, 0.0, "s-wave")
    Dj = orc.dynamical(n2, 0.0, "s-wave")
    rs, ts, _, _ = tm.fresnel(n1, n2, n1, n2)
    testing.assert_allclose(tm.transmission_matrix(rs, ts), np.linalg.inv(Di) @ Dj, rtol=1e-13)


def test_print_structure_and_csv_export(tm, tmp_path, capsys):
    s = tm.Structure()
    s.add_layer(tm.Layer(material="CaF2", thickness=0.0))
    s.add_layer(tm.Layer(material="Au", thickness=10.0e-3))
    s.print_structure()
    out = capsys.readouterr().out
    assert "Structure Configuration" in out and "Layer 1 : Au | d = 10 nm" in out
    tm.write_tmm_results(0.5, [1.0, 2.0], [0.9, 0.8], [0.1, 0.2], str(tmp_path))
    text = (tmp_path / "0.5deg_.csv").read_text().splitlines()
    assert text[0] == "Wavelength,Transmittance,Reflectance" and text[1] == "1.0,0.9,0.1"


def test_command_line_writes_one_csv_per_angle(tm, tmp_path, capsys):
    """`python -m pistachio_b200 structure.yaml out -s` (tm.py:755-805): the YAML is loaded, every angle evaluated and one
    `Wavelength,Transmittance,Reflectance` file per angle written into make_sim_dir(out); the files read back to the
    values angle_resolved_spectra returns."""
    import csv

    import yaml
    import importlib.resources as pkg_resources
    from pistachio_b200 import default

    cfg = yaml.safe_load((pkg_resources.files(default) / "Au_thin_layer.yaml").read_text())
    cfg["wave"]["theta_f"], cfg["wave"]["num_angles"] = 40.0, 3
    path = tmp_path / "structure.yaml"
    path.write_text(yaml.safe_dump(cfg))
    out_dir = tm.main(tm.parse_arguments([str(path), str(tmp_path), "-s"]))
    printed = capsys.readouterr().out
    assert "Loading structure parameters from" in printed and "Calculation time:" in printed and "Wrote results to" in printed
    s = tm.Structure()
    s.load_struct_from_config(str(path))
    expect = s.angle_resolved_spectra("s-wave")
    assert os.path.isdir(out_dir) and out_dir == s.make_sim_dir(str(tmp_path))
    files = sorted(os.listdir(out_dir))
    assert files == sorted(str(th) + "deg_.csv" for th in s.theta)
    for th, (T, R) in zip(s.theta, expect):
        with open(os.path.join(out_dir, str(th) + "deg_.csv")) as fh:
            rows = list(csv.reader(fh))
        assert rows[0] == ["Wavelength", "Transmittance", "Reflectance"] and len(rows) == 1 + len(s.wavelengths)
        got = np.array(rows[1:], dtype=np.float64)
        assert np.array_equal(got[:, 0], s.wavelengths) and np.array_equal(got[:, 1], T) and np.array_equal(got[:, 2], R)
    assert not np.array_equal(expect[0][0], expect[2][0])  # the angles really differ (the reference's would not)
