"""Deterministic synthetic workloads of the shapes BASELINE.json names.

A workload is a plain description -- wavelength grid, angle grid and an ordered
list of layers, each carrying its *own* (lambda, n, k) table and a thickness --
exactly what a user of the reference would put into Layer objects
(reference: transfer_matrix.py:40-50, 404-428).  Resampling the tables onto the
grid is the consumer's job: the CUDA path does it on the device
(pst_resample_tables), the tests do it with the oracle.  Units are micrometres
throughout, as in the reference's shipped YAML/CSV files.  No RNG is involved.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field

import numpy as np

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "materials.npz")
_cache = {}


def load_material(name: str):
    """(wavelength_um, n, k) table of a packaged material ('Au', 'CaF2', 'MgF2', 'ZnS')."""
    if not _cache:
        with np.load(_DATA) as z:
            for key in z.files:
                _cache[key] = z[key]
    try:
        return _cache[f"{name}/wl"], _cache[f"{name}/n"], _cache[f"{name}/k"]
    except KeyError:
        raise KeyError(f"no packaged material table named {name!r}") from None


def packaged_materials():
    load_material("Au")
    return sorted({k.split("/")[0] for k in _cache})


def constant_table(n: float, k: float = 0.0):
    """One-entry table: broadcast to every wavelength (reference: transfer_matrix.py:158-162)."""
    return np.array([1.0]), np.array([float(n)]), np.array([float(k)])


def lorentz_table(lo, hi, rows, center, gamma, amp, n_background):
    """Lorentz-oscillator index n = sqrt(n_b^2 - amp/(w^2 - w0^2 + i w gamma)) tabulated on a
    uniform grid, the model of the reference tutorial (Tutorial 1, cells 2-5)."""
    w = np.linspace(lo, hi, rows)
    n = np.sqrt(n_background ** 2 - amp / (w ** 2 - center ** 2 + 1j * w * gamma))
    return w, n.real.copy(), n.imag.copy()


@dataclass
class LayerSpec:
    material: str
    thickness: float
    tab_wl: np.ndarray
    tab_n: np.ndarray
    tab_k: np.ndarray


@dataclass
class Workload:
    name: str
    wl: np.ndarray
    theta: np.ndarray  # radians
    layers: list = field(default_factory=list)
    sweep_layer: int = -1  # index of the layer whose thickness is swept (-1: none)
    sweep_d: np.ndarray | None = None

    @property
    def interior_layers(self) -> int:
        return max(len(self.layers) - 2, 0)

    def points(self, n_pol: int = 2) -> int:
        ns = 1 if self.sweep_d is None else len(self.sweep_d)
        return ns * n_pol * len(self.wl) * len(self.theta)


def _mat(name, d):
    w, n, k = load_material(name)
    return LayerSpec(name, float(d), w, n, k)


def _const(name, n, d, k=0.0):
    w, nn, kk = constant_table(n, k)
    return LayerSpec(name, float(d), w, nn, kk)


def c1_fabry_perot(n_wl: int = 2000) -> Workload:
    """configs[0]: CaF2 / Au 10 nm / air 20 um / Au 10 nm / CaF2 (reference default/fabry-perot.yaml)."""
    layers = [_mat("CaF2", 0.0), _mat("Au", 10.0e-3), _const("Air", 1.000273, 20.0), _mat("Au", 10.0e-3), _mat("CaF2", 0.0)]
    return Workload("c1_fabry_perot", np.linspace(2.0, 6.0, n_wl), np.array([0.0]), layers)


def c2_dbr_cavity(n_wl: int = 8192, n_theta: int = 1024, pairs: int = 20) -> Workload:
    """configs[1]: air / (ZnS|MgF2)x20 / Lorentz half-wave cavity / (MgF2|ZnS)x20 / n=1.5 substrate."""
    d_h, d_l = 107.7586207e-3, 183.8235294e-3  # quarter-wave at 1.0 um (reference default/quarter-wave.yaml:21,25)
    cav = LayerSpec("LorentzCavity", 1.0 / (2 * 1.5), *lorentz_table(0.75, 1.30, 1000, 1.0, 0.004, 6.0e-4, 1.5))
    layers = [_const("Air", 1.0, 0.0)]
    for _ in range(pairs):
        layers += [_mat("ZnS", d_h), _mat("MgF2", d_l)]
    layers.append(cav)
    for _ in range(pairs):
        layers += [_mat("MgF2", d_l), _mat("ZnS", d_h)]
    layers.append(_const("Substrate", 1.5, 0.0))
    theta = np.deg2rad(np.linspace(0.0, 60.0, n_theta))
    return Workload("c2_dbr_cavity", np.linspace(0.8, 1.25, n_wl), theta, layers)


def c3_polariton(n_d: int = 100000, n_wl: int = 4096, n_theta: int = 64) -> Workload:
    """configs[2]: CaF2 / Au / absorbing Lorentz layer of swept thickness / Au / CaF2."""
    mol = LayerSpec("Molecule", 15.0, *lorentz_table(3.4, 4.6, 1200, 4.0, 0.015, 1.0e-2, 1.5))
    layers = [_mat("CaF2", 0.0), _mat("Au", 10.0e-3), mol, _mat("Au", 10.0e-3), _mat("CaF2", 0.0)]
    theta = np.deg2rad(np.linspace(0.0, 30.0, n_theta))
    return Workload("c3_polariton", np.linspace(3.5, 4.5, n_wl), theta, layers, 2, np.linspace(5.0, 25.0, n_d))


def c4_chirped(n_layers: int = 10000, n_wl: int = 16384, n_hi: float = 1.46, n_lo: float = 1.45) -> Workload:
    """configs[3]: air / n_layers alternating lossless layers, linearly chirped quarter-wave / n=1.5."""
    layers = [_const("Air", 1.0, 0.0)]
    design = np.linspace(0.8, 1.25, n_layers)
    for j in range(n_layers):
        n = n_hi if j % 2 == 0 else n_lo
        layers.append(_const("H" if j % 2 == 0 else "L", n, design[j] / (4.0 * n)))
    layers.append(_const("Substrate", 1.5, 0.0))
    return Workload("c4_chirped", np.linspace(0.8, 1.25, n_wl), np.array([0.0]), layers)


def c5_box_sweep(n_wl: int = 8192, n_theta: int = 256, n_d: int = 256) -> Workload:
    """configs[4]: the C1 stack with the cavity spacing as a fourth axis (2^30 points with both polarisations)."""
    w = c1_fabry_perot(n_wl)
    w.name = "c5_box_sweep"
    w.theta = np.deg2rad(np.linspace(0.0, 60.0, n_theta))
    w.sweep_layer = 2
    w.sweep_d = np.linspace(15.0, 25.0, n_d)
    return w
