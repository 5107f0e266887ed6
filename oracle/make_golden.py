"""TEST INFRASTRUCTURE ONLY.  Build-container script that produces the committed fixtures

    tests/golden/reference_cases.npz   outputs of the REAL reference on seeded inputs
    tests/golden/tutorial_goldens.npz  full-precision arrays stored in the reference tutorial's plots
    pistachio_b200/data/materials.npz  the (lambda, n, k) tables of the reference's packaged CSVs

Run:  python oracle/make_golden.py        (needs /root/reference; does not run on the GPU box)

Every case stores its inputs next to the reference's outputs, so the tests
never need the reference tree.  Keys are "<case>/<name>".
"""
from __future__ import annotations

import io
import contextlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle.ref_loader import load_reference, reference_default_yaml  # noqa: E402
from oracle import tmm_oracle as orc  # noqa: E402

NOTEBOOK = "/root/reference/tutorials/Tutorial 1 - Building a Structure.ipynb"


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def ref_structure(tm, wl, n_list, d_list):
    """Assemble a reference Structure whose layers already sit on the grid `wl`."""
    s = tm.Structure()
    for n, d in zip(n_list, d_list):
        layer = tm.Layer(material="m", thickness=float(d))
        layer.wavelengths = np.asarray(wl)
        layer.refractive_index = np.asarray(n).real.copy()
        layer.extinction_coeff = np.asarray(n).imag.copy()
        layer.complex_refractive = np.asarray(n, dtype=np.complex128)
        s.add_layer(layer)
    s.wavelengths = np.asarray(wl)
    return s


def ref_spectrum(tm, s, theta, pol):
    """BASELINE.md section 3 recipe: layer matrices must be rebuilt explicitly (tm.py:556 is commented out)."""
    s.calculate_layer_matrices(theta, pol)
    mats = s.calculate_all_transfer_matrices(theta, pol)
    T, R = s.calculate_all_t_r(mats)
    return T, R, np.array(mats)


def grid_case(tm, out, name, wl, n_list, d_list, thetas, keep_m_every=1):
    pols = ("s-wave", "p-wave")
    s = ref_structure(tm, wl, n_list, d_list)
    T = np.empty((2, len(wl), len(thetas)))
    R = np.empty_like(T)
    M = np.empty((2, len(wl[::keep_m_every]), len(thetas), 2, 2), dtype=np.complex128)
    for ip, pol in enumerate(pols):
        for it, th in enumerate(thetas):
            t, r, m = ref_spectrum(tm, s, float(th), pol)
            T[ip, :, it], R[ip, :, it] = t, r
            M[ip, :, it] = m[::keep_m_every]
    out[f"{name}/wl"] = np.asarray(wl)
    out[f"{name}/n"] = np.array(n_list, dtype=np.complex128)
    out[f"{name}/d"] = np.array(d_list, dtype=np.float64)
    out[f"{name}/theta"] = np.asarray(thetas, dtype=np.float64)
    out[f"{name}/T"] = T
    out[f"{name}/R"] = R
    out[f"{name}/M"] = M
    out[f"{name}/m_every"] = np.array(keep_m_every)


def yaml_case(tm, out, name, yaml_name, patch=None, keep_m_every=1):
    s = tm.Structure()
    quiet(s.load_struct_from_config, reference_default_yaml(yaml_name))
    if patch:
        patch(s)
        # the tutorial mutates thickness after loading; layer caches must be rebuilt
    s.calculate_layer_matrices(0.0, "s-wave")
    mats = s.calculate_all_transfer_matrices(0.0, "s-wave")
    T, R = s.calculate_all_t_r(mats)
    out[f"{name}/wl"] = np.asarray(s.wavelengths)
    out[f"{name}/theta_rad"] = np.asarray(s.theta)
    out[f"{name}/n"] = np.array([l.complex_refractive for l in s.layers])
    out[f"{name}/d"] = np.array([l.thickness for l in s.layers], dtype=np.float64)
    out[f"{name}/T"] = T
    out[f"{name}/R"] = R
    out[f"{name}/M"] = np.array(mats)[::keep_m_every]
    out[f"{name}/m_every"] = np.array(keep_m_every)


def resolved(w):
    """Workload -> dict with per-layer complex n on the grid (oracle resampling == reference make_datapoints, tested)."""
    n_list = []
    for l in w.layers:
        n, k = orc.resample_table(l.tab_wl, l.tab_n, l.tab_k, w.wl)
        n_list.append(n + 1j * k)
    return {"wl": w.wl, "theta": w.theta, "n": n_list, "d": [l.thickness for l in w.layers],
            "sweep_layer": w.sweep_layer, "sweep_d": w.sweep_d}


def package_materials(tm):
    """(lambda, n, k) tables of the reference's data files -> pistachio_b200/data/materials.npz.  The four tables next
    to transfer_matrix.py are read through the reference's own Layer.get_data_from_csv; the n / nk tables of the
    refractiveindex_info/ sub-folder (which that method cannot reach: importlib.resources takes no sub-paths) are parsed
    here with the same column convention (wavelength in micrometres, n, optional k; k = 0 when absent).  Spectra
    (transmittance_*/reflectance_* files) are not index tables and are left out."""
    import csv
    import glob

    mats = {}
    for fname in ("Au.csv", "CaF2.csv", "MgF2.csv", "ZnS.csv"):
        l = tm.Layer()
        l.get_data_from_csv(fname)
        key = fname[:-4]
        mats[f"{key}/wl"] = l.wavelengths
        mats[f"{key}/n"] = l.refractive_index
        mats[f"{key}/k"] = l.extinction_coeff
    folder = os.path.join(os.path.dirname(tm.__file__), "data", "refractive_index_data", "refractiveindex_info")
    for path in sorted(glob.glob(os.path.join(folder, "*.csv"))):
        stem = os.path.basename(path)[:-4]
        if stem.lower().startswith(("transmittance", "reflectance")):
            continue
        wl, n, k = [], [], []
        with open(path, "r", encoding="utf-8-sig") as fh:
            rows = csv.reader(fh)
            next(rows, None)
            for row in rows:
                if len(row) < 2 or row[0] == "":
                    continue
                wl.append(float(row[0]))
                n.append(float(row[1]))
                k.append(float(row[2]) if len(row) > 2 and row[2] != "" else 0.0)
        mats[f"{stem}/wl"], mats[f"{stem}/n"], mats[f"{stem}/k"] = np.array(wl), np.array(n), np.array(k)
    os.makedirs(os.path.join(ROOT, "pistachio_b200", "data"), exist_ok=True)
    np.savez_compressed(os.path.join(ROOT, "pistachio_b200", "data", "materials.npz"), **mats)


def main():
    tm = load_reference()
    out = {}

    package_materials(tm)
    if "--materials-only" in sys.argv:
        return

    # ---- known-answer tests of the reference (tests/test_layer.py:67-91, tests/test_structure.py:65-71)
    out["kat/D_s"] = tm.Layer().dynamical_matrix(1.0 + 1.0j, 0.0, "s-wave")
    out["kat/D_p"] = tm.Layer().dynamical_matrix(1.0 + 1.0j, 0.0, "p-wave")
    kx = 2.0e5 + 0.0j
    out["kat/P"] = tm.Layer().propagation_matrix(kx, np.pi / (2 * kx))
    lay = tm.Layer(material="Air", thickness=3.0, num_points=1)
    out["kat/M_air"] = np.array(lay.calculate_transfer_matrix(1.0 + 0.0j, 1.0, 0.0, "s-wave"))
    # a generic absorbing oblique layer, all four matrices, both polarisations
    lay2 = tm.Layer(material="x", thickness=0.37)
    out["kat/M4_s"] = np.array(lay2.calculate_transfer_matrix(1.7 + 0.02j, 0.83, 0.6, "s-wave"))
    out["kat/M4_p"] = np.array(lay2.calculate_transfer_matrix(1.7 + 0.02j, 0.83, 0.6, "p-wave"))

    # ---- interpolation (tests/test_layer.py:26-63 recipe) + extrapolation + unsorted table
    l = tm.Layer()
    l.get_data_from_csv("Au.csv")
    q = np.linspace(0.2, 15.0, 10000)
    l.make_datapoints(q)
    out["interp_au/wl"] = q
    out["interp_au/n"] = l.refractive_index
    out["interp_au/k"] = l.extinction_coeff
    l = tm.Layer()
    l.get_data_from_csv("CaF2.csv")
    q = np.linspace(3.5e-6, 4.5e-6, 257)  # the tutorial's metres-vs-micrometres extrapolation
    l.make_datapoints(q)
    out["interp_caf2_extrap/wl"] = q
    out["interp_caf2_extrap/n"] = l.refractive_index
    out["interp_caf2_extrap/k"] = l.extinction_coeff
    rng = np.random.default_rng(7)
    tw = rng.uniform(0.3, 3.0, 41)
    tn = rng.uniform(1.0, 3.0, 41)
    tk = rng.uniform(0.0, 1.0, 41)
    l = tm.Layer()
    l.wavelengths, l.refractive_index, l.extinction_coeff = tw, tn, tk
    q = np.concatenate([np.linspace(0.1, 3.5, 300), tw[:5]])
    l.make_datapoints(q)
    out["interp_unsorted/tab_wl"], out["interp_unsorted/tab_n"], out["interp_unsorted/tab_k"] = tw, tn, tk
    out["interp_unsorted/wl"] = q
    out["interp_unsorted/n"] = l.refractive_index
    out["interp_unsorted/k"] = l.extinction_coeff

    # ---- shipped YAML structures
    yaml_case(tm, out, "yaml_au_thin", "Au_thin_layer.yaml")
    yaml_case(tm, out, "yaml_quarter_wave", "quarter-wave.yaml", keep_m_every=10)
    yaml_case(tm, out, "yaml_fabry_perot", "fabry-perot.yaml", keep_m_every=100)

    def thin_mirror(s):
        s.layers[3].thickness = 2e-3

    yaml_case(tm, out, "yaml_fabry_perot_2nm", "fabry-perot.yaml", patch=thin_mirror, keep_m_every=100)

    # ---- BASELINE configs (C1 full, C2/C3/C4 sub-sampled) built by pistachio_b200/workloads.py
    from pistachio_b200 import workloads

    w = resolved(workloads.c1_fabry_perot())
    grid_case(tm, out, "c1", w["wl"], w["n"], w["d"], w["theta"], keep_m_every=20)
    w = resolved(workloads.c2_dbr_cavity(n_wl=8192, n_theta=1024))
    sel_l = np.arange(0, 8192, 128)
    sel_t = np.array([0, 341, 682, 1023])
    grid_case(tm, out, "c2_sub", w["wl"][sel_l], [n[sel_l] for n in w["n"]], w["d"], w["theta"][sel_t])
    out["c2_sub/sel_l"], out["c2_sub/sel_t"] = sel_l, sel_t
    w = resolved(workloads.c3_polariton(n_d=100000, n_wl=4096, n_theta=64))
    sel_l = np.arange(0, 4096, 128)
    sel_t = np.array([0, 21, 42, 63])
    sel_d = np.array([0, 33333, 66666, 99999])
    for k in sel_d:
        d = list(w["d"])
        d[w["sweep_layer"]] = float(w["sweep_d"][k])
        grid_case(tm, out, f"c3_sub_d{k}", w["wl"][sel_l], [n[sel_l] for n in w["n"]], d, w["theta"][sel_t])
    out["c3_sub/sel_l"], out["c3_sub/sel_t"], out["c3_sub/sel_d"] = sel_l, sel_t, sel_d
    w = resolved(workloads.c4_chirped(n_layers=10000, n_wl=16384))
    sel_l = np.arange(0, 16384, 4096)
    grid_case(tm, out, "c4_sub", w["wl"][sel_l], [n[sel_l] for n in w["n"]], w["d"], w["theta"])
    out["c4_sub/sel_l"] = sel_l

    # ---- seeded random stacks (SURVEY.md section 8d "fuzz parity", reduced count)
    rng = np.random.default_rng(0)
    n_stacks = 240
    out["fuzz/count"] = np.array(n_stacks)
    for i in range(n_stacks):
        L = int(rng.integers(2, 65)) if i >= 8 else 2 + i  # first few: the shallowest stacks incl. no interior layer
        wl = np.sort(rng.uniform(0.4, 2.0, 4))
        nr = rng.uniform(1.0, 4.0, (L, 1)) * np.ones((1, 4))
        lossy = rng.random(L) < 0.5
        kap = np.where(lossy, 10.0 ** rng.uniform(-4, 1, L), 0.0)[:, None] * np.ones((1, 4))
        n = nr + 1j * kap
        n[0] = n[0].real  # incidence medium lossless
        d = rng.uniform(0.0, 5.0, L) * wl.mean() / nr[:, 0]
        # keep the reference finite: bound the total attenuation exponent
        att = (2 * np.pi * kap[:, 0] * d / wl.min())
        d = np.where(att > 30.0, d * 30.0 / np.maximum(att, 1e-300), d)
        th = rng.uniform(0.0, 1.5, 2)
        grid_case(tm, out, f"fuzz{i}", wl, list(n), list(d), th)

    # ---- tutorial example 1 (Lorentz sample between CaF2 plates; metres grid vs micrometre table)
    num_points = 1000
    wavelength = np.linspace(3.5e-6, 4.5e-6, num_points)
    lor = 1.0e-14 / (wavelength ** 2 - (4.0e-6) ** 2 + 1j * wavelength * 1.5e-8)
    refractive = np.sqrt(1.5 ** 2 - lor)
    s = tm.Structure()
    a = tm.Layer(material="CaF2", thickness=0.0)
    a.get_data_from_csv("CaF2.csv")
    s.add_layer(a)
    smp = tm.Layer(material="Sample")
    smp.wavelengths = wavelength
    smp.refractive_index = refractive.real
    smp.extinction_coeff = refractive.imag
    smp.set_complex_refractive(refractive.real, refractive.imag)
    s.add_layer(smp)
    s.layers[1].thickness = 2.0e-5
    b = tm.Layer(material="CaF2", thickness=0.0)
    b.get_data_from_csv("CaF2.csv")
    s.add_layer(b)
    s.initialize_struct(0.0, 0.0, 1, 3.5e-6, 4.5e-6, 5000, "s-wave")
    mats = s.calculate_all_transfer_matrices(0.0, "s-wave")
    T, R = s.calculate_all_t_r(mats)
    out["tutorial1/table_wl"] = wavelength
    out["tutorial1/table_n"] = refractive
    out["tutorial1/wl"] = np.asarray(s.wavelengths)
    out["tutorial1/n"] = np.array([l.complex_refractive for l in s.layers])
    out["tutorial1/d"] = np.array([l.thickness for l in s.layers])
    out["tutorial1/T"] = T
    out["tutorial1/R"] = R

    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "reference_cases.npz"), **out)
    print("reference_cases.npz:", len(out), "arrays")

    # ---- arrays stored in the tutorial notebook's plot outputs
    nb = json.load(open(NOTEBOOK))
    figs = []
    for cell in nb["cells"]:
        for o in cell.get("outputs", []):
            pj = o.get("data", {}).get("application/vnd.plotly.v1+json")
            if pj:
                figs.append(pj)
    gold = {
        "lorentz/x_um": np.array(figs[0]["data"][0]["x"], dtype=np.float64),
        "lorentz/kappa": np.array(figs[0]["data"][0]["y"], dtype=np.float64),
        "tutorial1/wavenumber": np.array(figs[1]["data"][0]["x"], dtype=np.float64),
        "tutorial1/T_percent": np.array(figs[1]["data"][0]["y"], dtype=np.float64),
        "fabry_perot/wavenumber": np.array(figs[2]["data"][0]["x"], dtype=np.float64),
        "fabry_perot/T_percent": np.array(figs[2]["data"][0]["y"], dtype=np.float64),
        "fabry_perot/R_percent": np.array(figs[2]["data"][1]["y"], dtype=np.float64),
    }
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "tutorial_goldens.npz"), **gold)
    print("tutorial_goldens.npz:", {k: v.shape for k, v in gold.items()})


if __name__ == "__main__":
    main()
