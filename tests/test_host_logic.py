"""CPU-only tests: host logic of the drop-in classes, the C-ABI library (loads, exports every declared symbol,
struct layout matches the header), workloads, and loud failure without a GPU.  No compute call is made."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "pistachio_b200.h")


@pytest.fixture(scope="module")
def lib_path():
    from pistachio_b200 import build

    return build.build()


def test_library_exports_every_declared_symbol(lib_path):
    from pistachio_b200 import _cabi

    text = open(HEADER).read()
    declared = set(re.findall(r"^(?:int|int64_t|const char\*)\s+(pst_[a-z0-9_]+)\s*\(", text, re.M))
    assert declared == set(_cabi.SYMBOLS), declared ^ set(_cabi.SYMBOLS)
    dll = C.CDLL(lib_path)
    for name in declared:
        assert hasattr(dll, name), name
    dll.pst_abi_version.restype = C.c_int
    assert dll.pst_abi_version() == _cabi.ABI_VERSION


def test_grid_args_layout_matches_header(tmp_path):
    from pistachio_b200 import _cabi

    src = tmp_path / "layout.c"
    fields = [f for f, _ in _cabi.GridArgs._fields_]
    body = "".join(f'printf("{f} %zu\\n", offsetof(pst_grid_args, {f}));' for f in fields)
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "pistachio_b200.h"\n'
                   f'int main(void){{printf("size %zu\\n", sizeof(pst_grid_args));{body}return 0;}}')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)], check=True)
    out = dict(line.split() for line in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    assert int(out["size"]) == C.sizeof(_cabi.GridArgs)
    for f in fields:
        assert int(out[f]) == getattr(_cabi.GridArgs, f).offset, f


def test_no_gpu_fails_loudly(lib_path):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    from pistachio_b200 import _cabi, engine
    from pistachio_b200 import transfer_matrix as tm

    h = C.c_void_p()
    lib = _cabi.load_library()
    rc = lib.pst_create(0, C.byref(h))
    assert rc == -4 and b"no CPU fallback" in lib.pst_last_error()
    with pytest.raises(_cabi.PistachioError):
        engine.get_engine()
    with pytest.raises(_cabi.PistachioError):
        tm.Layer().dynamical_matrix(1.0 + 1.0j, 0.0, "s-wave")
    with pytest.raises(_cabi.PistachioError):
        tm.Structure().calculate_t_r(np.eye(2, dtype=complex))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "pistachio_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f
                assert "/root/reference" not in text, f


# ---------------------------------------------------------------- reference tests/test_structure.py list semantics
def make_structure(tm, num_layers, lmbda, thickness):
    s = tm.Structure()
    for i in range(num_layers):
        layer = tm.Layer(material="Material_" + str(i), thickness=thickness)
        layer.wavelengths = [lmbda]
        s.add_layer(layer)
    return s


def test_add_and_delete_layer():
    from pistachio_b200 import transfer_matrix as tm

    s = make_structure(tm, 3, 1.0e-5, 3.0e-5)
    new_layer = tm.Layer(material="New Material", thickness=1.0e-6)
    new_layer.wavelengths = 1.0e-5
    s.add_layer(new_layer)
    assert s.layers[-1].material == "New Material"
    s = make_structure(tm, 3, 1.0e-6, 1.0e-3)
    s.delete_layer(1)
    assert s.layers[1].material == "Material_2"


def test_layer_defaults_and_attribute_surface():
    from pistachio_b200 import transfer_matrix as tm

    l = tm.Layer()
    assert l.material == "Air" and l.thickness == 1.0e-3
    for attr in ("wavelengths", "refractive_index", "extinction_coeff", "kx", "kz", "transfer_matrices"):
        assert len(getattr(l, attr)) == 0
    assert l.complex_refractive.shape == (0,)
    l.set_complex_refractive([1.0, 2.0], [0.5, 0.0])
    assert l.complex_refractive.dtype == np.complex128 and l.complex_refractive[0] == 1.0 + 0.5j
    s = tm.Structure()
    for attr in ("layers", "theta", "wavelengths", "wavevectors", "transfer_matrices"):
        assert len(getattr(s, attr)) == 0
    for name in ("add_layer", "delete_layer", "load_yaml_config", "load_struct_from_config", "initialize_struct",
                 "calculate_t_r", "calculate_all_t_r", "calculate_layer_matrices", "calculate_all_transfer_matrices",
                 "one_job", "angle_resolved_spectra", "print_structure"):
        assert callable(getattr(s, name))
    for name in ("set_complex_refractive", "get_data_from_csv", "make_datapoints", "calculate_wavenumbers",
                 "dynamical_matrix", "propagation_matrix", "calculate_transfer_matrix", "calculate_all_transfer_matrices"):
        assert callable(getattr(l, name))
    assert callable(tm.fresnel) and callable(tm.transmission_matrix) and callable(tm.write_tmm_results)


def test_csv_loading_and_yaml_parsing_without_gpu(tmp_path, monkeypatch):
    from pistachio_b200 import default
    from pistachio_b200 import transfer_matrix as tm
    import importlib.resources as pkg_resources

    l = tm.Layer()
    l.get_data_from_csv("Au.csv")
    assert len(l.wavelengths) == 333 and abs(l.wavelengths[0] - 0.19077) < 1e-12 and l.extinction_coeff[0] == 1.352
    l.get_data_from_csv("CaF2.csv")
    assert len(l.wavelengths) == 101 and np.all(l.extinction_coeff == 0.0)
    with pytest.raises(FileNotFoundError):
        l.get_data_from_csv("Unobtainium.csv")
    p = tmp_path / "mine.csv"
    p.write_text('"Wavelength, um","n","k"\n1.0,1.5,0.1\n2.0,1.6\n')
    monkeypatch.setattr(tm, "CSV_SEARCH_PATH", [str(tmp_path)])
    l.get_data_from_csv("mine.csv")
    assert list(l.refractive_index) == [1.5, 1.6] and list(l.extinction_coeff) == [0.1, 0.0]
    cfg = tm.Structure().load_yaml_config(str(pkg_resources.files(default) / "fabry-perot.yaml"))
    assert cfg["num_points"] == 10000 and len(cfg["layers"]) == 5 and float(cfg["layers"]["layer1"]["thickness"]) == 0.01


def test_cli_arguments_and_sim_dir_naming(tmp_path):
    """The command line of tm.py:755-805 and a make_sim_dir that runs (the reference's raises UnboundLocalError at
    tm.py:669): name = materials joined by '_' + '<theta_i>--<theta_f>deg_<lambda_i>--<lambda_f>lambda/'."""
    from pistachio_b200 import transfer_matrix as tm

    a = tm.parse_arguments(["struct.yaml", "out", "-p"])
    assert a.structure == "struct.yaml" and a.output == "out" and a.pwave and not a.swave
    a = tm.parse_arguments(["struct.yaml", "out", "--swave"])
    assert a.swave and not a.pwave
    with pytest.raises(SystemExit):  # neither -p nor -s: the reference exits silently (tm.py:779-780)
        tm.main(tm.parse_arguments(["struct.yaml", "out"]))
    s = tm.Structure()
    for name in ("CaF2", "Au", "Air"):
        s.add_layer(tm.Layer(material=name))
    s.theta = np.deg2rad(np.linspace(0.0, 30.0, 4))
    s.wavelengths = np.linspace(2.0, 6.0, 5)
    assert s.make_sim_dir(str(tmp_path)) == os.path.join(str(tmp_path), "CaF2_Au_Air_0.0--30.0deg_2.0--6.0lambda/")
    assert s.make_sim_dir(str(tmp_path), sim_name="run7") == os.path.join(str(tmp_path), "run7")
    tm.write_tmm_results(np.deg2rad(12.5), [1.0], [0.5], [0.25], str(tmp_path), notebook_names=True)
    assert (tmp_path / "deg12.50_.csv").read_text().splitlines()[1] == "1.0,0.5,0.25"


def test_workload_shapes():
    from pistachio_b200 import workloads as wk

    w = wk.c2_dbr_cavity()
    assert len(w.layers) == 83 and w.interior_layers == 81 and w.points() == 16_777_216
    w = wk.c3_polariton()
    assert w.points(1) == 100_000 * 4096 * 64 and w.sweep_layer == 2
    w = wk.c4_chirped()
    assert w.interior_layers == 10_000 and w.points(1) == 16_384
    w = wk.c5_box_sweep()
    assert w.points() == 2 ** 30
    w = wk.c1_fabry_perot()
    assert w.points() == 4000 and w.interior_layers == 3
