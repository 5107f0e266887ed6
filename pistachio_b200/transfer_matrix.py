"""Drop-in mirror of pistachio's `transfer_matrix` module on top of the B200 engine.

    from pistachio_b200 import transfer_matrix as tm   # instead of: from pistachio import transfer_matrix as tm

Same classes, constructor signatures, public attributes, method names, argument
order, string enums ("s-wave"/"p-wave") and return conventions as the reference
(garrekstemo/pistachio, src/pistachio/transfer_matrix.py; cited as tm.py:LINE).
Every numeric step runs on the GPU through libpistachio_b200.so; host code here
only marshals attributes.  There is no CPU fallback: without the CUDA library or
a GPU the compute methods raise.

Documented differences from the reference (DESIGN.md "Reference quirks"):
  * Structure.calculate_all_transfer_matrices(theta, pol) honours its arguments.
    The reference ignores them (tm.py:556 is commented out) and chains whatever
    the layers cached last.  Set `Structure.strict_reference_cache = True` to get
    the reference's cached behaviour (device kernel pst_chain_matrices).
  * Layer.transfer_matrices is computed lazily on first access.
  * initialize_struct keeps the reference's degrees-as-radians use of theta_i for
    the per-layer caches (tm.py:461-462); Structure.theta is converted correctly.
Additive API: Structure.spectrum_grid(...) evaluates whole (pol, wavelength, theta
[, thickness]) grids in one launch and returns R, T, A.  `Structure.snell = True`
(default False = the reference's model, which keeps the incidence angle in every
layer, tm.py:270) switches the Structure-level calls to refraction mode: theta is
the angle in the incidence medium and every layer propagates at its refracted
angle (PST_FLAG_SNELL).
"""
from __future__ import annotations

import argparse
import csv
import os
import sys
import time

import numpy as np

from . import engine as _engine
from .workloads import load_material, packaged_materials

S_WAVE = "s-wave"
P_WAVE = "p-wave"
C0 = 299792458.0

# Directories searched by Layer.get_data_from_csv before the packaged tables.
CSV_SEARCH_PATH = [os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "refractive_index_data")]
if os.environ.get("PISTACHIO_DATA_DIR"):
    CSV_SEARCH_PATH.insert(0, os.environ["PISTACHIO_DATA_DIR"])


def _eng():
    return _engine.get_engine()


def _to_numpy(t):
    return t.cpu().numpy()


class Layer:
    """One planar layer: material name, thickness and (wavelength, n, k) data (tm.py:25-50)."""

    def __init__(self, material="Air", thickness=1.0e-3, num_points=100):
        self.material = material
        self.thickness = thickness
        self.wavelengths = []
        self.refractive_index = []
        self.extinction_coeff = []
        self.set_complex_refractive(self.refractive_index, self.extinction_coeff)
        self.kx = []
        self.kz = []
        self._tm_request = None  # (wl, n, d, theta, pol) snapshot for the lazy cache
        self._tm_device = None   # torch complex128 (n_wl, 4, 2, 2)
        self._tm_host = []

    # ---- material data -------------------------------------------------------------------------
    def set_complex_refractive(self, n_real=None, n_imag=None):
        """complex_refractive = n + i k as complex128 (tm.py:52-70)."""
        n_real = np.asarray(n_real, dtype=np.float64).ravel()
        n_imag = np.asarray(n_imag, dtype=np.float64).ravel()
        if n_imag.size < n_real.size:
            raise IndexError("list index out of range")
        self.complex_refractive = (n_real + 1j * n_imag[: n_real.size]).astype(np.complex128)

    def get_data_from_csv(self, refractive_filename):
        """Load a refractiveindex.info style table: header row, then wavelength (kept in micrometres, as the
        reference does -- its conversion is disabled at tm.py:104-105), n and optionally k (0.0 when the column is
        missing).  Files are looked up in CSV_SEARCH_PATH; the four tables the reference ships (Au, CaF2, MgF2,
        ZnS) are packaged as arrays."""
        wl = n = k = None
        for folder in CSV_SEARCH_PATH:
            path = os.path.join(folder, refractive_filename)
            if os.path.isfile(path):
                w_, n_, k_ = [], [], []
                with open(path, "r", encoding="utf-8") as fh:
                    rows = csv.reader(fh)
                    next(rows, None)
                    for row in rows:
                        w_.append(float(row[0]))
                        n_.append(float(row[1]))
                        k_.append(float(row[2]) if len(row) > 2 and row[2] != "" else 0.0)
                wl, n, k = np.array(w_), np.array(n_), np.array(k_)
                break
        if wl is None:
            stem = refractive_filename[:-4] if refractive_filename.lower().endswith(".csv") else refractive_filename
            if stem not in packaged_materials():
                raise FileNotFoundError(f"no refractive index table {refractive_filename!r} in {CSV_SEARCH_PATH} "
                                        f"or among the packaged materials {packaged_materials()}")
            wl, n, k = (np.array(a) for a in load_material(stem))
        self.wavelengths = wl
        self.refractive_index = n
        self.extinction_coeff = k
        self.set_complex_refractive(n, k)

    def make_datapoints(self, wavelengths):
        """Resample this layer's table onto `wavelengths` (tm.py:117-170) with the device lookup kernel
        (pst_resample_tables): one-entry tables broadcast, otherwise scipy-interp1d linear inter/extrapolation,
        bit-identical to the reference.  Overwrites the table in place like the reference does."""
        wavelengths = np.asarray(wavelengths, dtype=np.float64)
        n_tab = np.asarray(self.refractive_index, dtype=np.float64).ravel()
        k_tab = np.asarray(self.extinction_coeff, dtype=np.float64).ravel()
        w_tab = np.asarray(self.wavelengths, dtype=np.float64).ravel() if len(n_tab) != 1 else np.zeros(1)
        mat_n, mat_k = _eng().resample_tables([(w_tab, n_tab, k_tab)], wavelengths.ravel())
        self.refractive_index = _to_numpy(mat_n[0])
        self.extinction_coeff = _to_numpy(mat_k[0])
        self.set_complex_refractive(self.refractive_index, self.extinction_coeff)
        self.wavelengths = wavelengths

    def calculate_wavenumbers(self, n_complex, theta=0.0):
        """kx = n w/c cos(theta), kz = n w/c sin(theta) attributes (tm.py:172-188)."""
        kx, kz = _eng().wavenumbers(n_complex, self.wavelengths, theta)
        self.kx = _to_numpy(kx)
        self.kz = _to_numpy(kz)

    # ---- single matrices ---------------------------------------------------------------------------
    def dynamical_matrix(self, n_complex, theta, wave_type="s-wave"):
        """2x2 dynamical matrix (tm.py:190-215).  An unknown wave_type prints the reference's message and
        returns None."""
        if wave_type not in (S_WAVE, P_WAVE):
            print("Must specify 's-wave' or 'p-wave'.")
            return None
        out = _eng().layer_matrices([n_complex], [1.0], theta, wave_type, 0.0)
        return _to_numpy(out[0, 0])

    def propagation_matrix(self, kx, d):
        """diag(exp(-i kx d), exp(+i kx d)) (tm.py:217-234)."""
        return _to_numpy(_eng().propagation_matrices([kx], d)[0])

    def calculate_transfer_matrix(self, n_complex, lmbda, theta, wave_type):
        """[D, P, D^-1, M] for one wavelength with this layer's thickness (tm.py:236-277)."""
        if wave_type not in (S_WAVE, P_WAVE):
            print("Must specify 's-wave' or 'p-wave'.")
            raise np.linalg.LinAlgError("0-dimensional array given. Array must be at least two-dimensional")
        out = _to_numpy(_eng().layer_matrices([n_complex], [lmbda], theta, wave_type, self.thickness))
        return [out[0, 0], out[0, 1], out[0, 2], out[0, 3]]

    def calculate_all_transfer_matrices(self, theta, wave_type):
        """Cache [D, P, D^-1, M] for every stored wavelength (tm.py:279-305).  The request is snapshotted now and
        evaluated on the device when `transfer_matrices` is first read or chained."""
        if wave_type not in (S_WAVE, P_WAVE):
            print("Must specify 's-wave' or 'p-wave'.")
            raise np.linalg.LinAlgError("0-dimensional array given. Array must be at least two-dimensional")
        wl = np.array(self.wavelengths, dtype=np.float64).ravel()
        n = np.array(self.complex_refractive, dtype=np.complex128).ravel()
        if n.size < wl.size:
            raise IndexError("index out of bounds: fewer refractive-index entries than wavelengths")
        self._tm_request = (wl, n[: wl.size], float(self.thickness), float(theta), wave_type)
        self._tm_device = None
        self._tm_host = None

    def _cache_device(self):
        if self._tm_device is None:
            if self._tm_request is None:
                return None
            wl, n, d, theta, pol = self._tm_request
            if wl.size == 0:
                return None
            self._tm_device = _eng().layer_matrices(n, wl, theta, pol, d)
        return self._tm_device

    @property
    def transfer_matrices(self):
        if self._tm_host is None:
            dev = self._cache_device()
            self._tm_host = [] if dev is None else _to_numpy(dev)
        return self._tm_host

    @transfer_matrices.setter
    def transfer_matrices(self, value):
        self._tm_request = None
        self._tm_device = None
        self._tm_host = value


class Structure:
    """Ordered stack of Layers: index 0 is the incidence medium, -1 the exit medium (tm.py:308-353)."""

    strict_reference_cache = False
    snell = False  # True: Snell refraction per layer (additive; not the reference's model)

    def _flags(self):
        return _engine.FLAG_SNELL if self.snell else 0

    def __init__(self):
        self.layers = []
        self.theta = []
        self.wavelengths = []
        self.wavevectors = []
        self.transfer_matrices = []

    def add_layer(self, layer):
        self.layers.append(layer)

    def delete_layer(self, layer_index):
        self.layers.pop(layer_index)

    # ---- configuration ---------------------------------------------------------------------------------
    def load_yaml_config(self, config_file):
        """Parse a structure YAML file into a dict (tm.py:355-376)."""
        import yaml

        with open(config_file, "r") as fh:
            return yaml.safe_load(fh)

    def load_struct_from_config(self, config_file):
        """Build the layers a YAML file describes and initialise the structure (tm.py:378-432)."""
        cfg = self.load_yaml_config(config_file)
        wave = cfg["wave"]
        layers = []
        for i in range(len(cfg["layers"])):
            spec = cfg["layers"]["layer" + str(i)]
            layer = Layer()
            layer.material = str(spec["material"])
            layer.thickness = float(spec["thickness"])
            if "refractive_filename" in spec:
                name = spec["refractive_filename"]
                if "csv" in name:
                    layer.get_data_from_csv(name)
                else:
                    print("Cannot get refractive index data. Check file type (must be .csv).")
            elif "refractive_index" in spec:
                layer.refractive_index = np.array([spec["refractive_index"]])
                layer.extinction_coeff = np.array([spec["extinction_coeff"]])
                layer.set_complex_refractive(layer.refractive_index, layer.extinction_coeff)
                layer.wavelengths = np.array([spec["wavelength"]])
            else:
                print("ERROR: Incorrect yaml config format. Reference default template.")
            layers.append(layer)
        self.layers = layers
        print("Initializing structure...")
        self.initialize_struct(float(wave["theta_i"]), float(wave["theta_f"]), int(wave["num_angles"]),
                               float(cfg["min_wavelength"]), float(cfg["max_wavelength"]), int(cfg["num_points"]),
                               str(wave["polarization"]))

    def initialize_struct(self, theta_i, theta_f, num_angles, min_wl, max_wl, num_wl, wave_type):
        """Common wavelength grid, angle grid (degrees -> radians) and per-layer resampling (tm.py:434-462).
        All layers are resampled in ONE device launch; their wavenumber attributes and matrix caches are requested
        at theta_i exactly as the reference does (theta_i is used un-converted there, tm.py:461-462)."""
        self.wavelengths = np.linspace(min_wl, max_wl, num_wl)
        self.theta = np.deg2rad(np.linspace(theta_i, theta_f, num_angles))
        if not self.layers:
            return
        tables = []
        for layer in self.layers:
            n_tab = np.asarray(layer.refractive_index, dtype=np.float64).ravel()
            k_tab = np.asarray(layer.extinction_coeff, dtype=np.float64).ravel()
            w_tab = np.asarray(layer.wavelengths, dtype=np.float64).ravel() if len(n_tab) != 1 else np.zeros(1)
            tables.append((w_tab, n_tab, k_tab))
        mat_n, mat_k = _eng().resample_tables(tables, self.wavelengths)
        mat_n, mat_k = _to_numpy(mat_n), _to_numpy(mat_k)
        for i, layer in enumerate(self.layers):
            layer.refractive_index = mat_n[i].copy()
            layer.extinction_coeff = mat_k[i].copy()
            layer.set_complex_refractive(layer.refractive_index, layer.extinction_coeff)
            layer.wavelengths = self.wavelengths
            layer.calculate_wavenumbers(layer.complex_refractive, theta_i)
            layer.calculate_all_transfer_matrices(theta_i, wave_type)

    # ---- T, R -------------------------------------------------------------------------------------------
    def calculate_t_r(self, M):
        """(T, R) of one total matrix: t = 1/M00, r = M10/M00, R = |r|^2, T = Re(det M |t|^2) (tm.py:464-488)."""
        T, R = _eng().t_r_from_matrices(np.asarray(M, dtype=np.complex128).reshape(1, 2, 2))
        return _to_numpy(T)[0], _to_numpy(R)[0]

    def calculate_all_t_r(self, transfer_matrices):
        """(T_array, R_array) for a sequence of total matrices (tm.py:490-514)."""
        import torch

        if not isinstance(transfer_matrices, torch.Tensor):
            transfer_matrices = np.asarray(transfer_matrices, dtype=np.complex128).reshape(-1, 2, 2)
        T, R = _eng().t_r_from_matrices(transfer_matrices)
        return _to_numpy(T), _to_numpy(R)

    # ---- matrices -----------------------------------------------------------------------------------------
    def calculate_layer_matrices(self, theta, wave_type):
        """Request every layer's matrix cache at (theta, wave_type) (tm.py:516-534)."""
        for layer in self.layers:
            layer.calculate_all_transfer_matrices(theta, wave_type)

    def _stack_on_grid(self):
        """(wl, mat_n, mat_k, layer_mat, layer_d) from the layers' current attributes, de-duplicating layers whose
        index arrays are identical."""
        wl = np.asarray(self.wavelengths, dtype=np.float64).ravel()
        seen, cols_n, cols_k, layer_mat = {}, [], [], []
        # a one-layer structure is its own incidence and exit medium in the reference (layers[0] is layers[-1])
        layers = self.layers if len(self.layers) != 1 else [self.layers[0], self.layers[0]]
        for layer in layers:
            n = np.asarray(layer.complex_refractive, dtype=np.complex128).ravel()
            if n.size != wl.size:
                raise ValueError(f"layer {layer.material!r} has {n.size} refractive-index points for {wl.size} "
                                 "wavelengths; call initialize_struct or make_datapoints first")
            key = n.tobytes()
            if key not in seen:
                seen[key] = len(cols_n)
                cols_n.append(n.real)
                cols_k.append(n.imag)
            layer_mat.append(seen[key])
        layer_d = [float(layer.thickness) for layer in layers]
        return wl, np.array(cols_n), np.array(cols_k), layer_mat, layer_d

    def calculate_all_transfer_matrices(self, theta, wave_type):
        """Total matrix M = D_0^-1 (M_1 ... M_{N-2}) D_{N-1} for every wavelength (tm.py:536-572).
        Returns an (n_wl, 2, 2) complex128 array (iterates like the reference's list of 2x2 arrays)."""
        if len(self.wavelengths) == 0:
            return np.empty((0, 2, 2), dtype=np.complex128)  # the reference's loop body never runs (tm.py:559)
        if len(self.layers) < 1:
            raise IndexError("list index out of range")
        if self.strict_reference_cache:
            return self._chain_cached()
        if wave_type not in (S_WAVE, P_WAVE):
            print("Must specify 's-wave' or 'p-wave'.")
            raise np.linalg.LinAlgError("0-dimensional array given. Array must be at least two-dimensional")
        wl, mat_n, mat_k, layer_mat, layer_d = self._stack_on_grid()
        res = _eng().eval_grid(wl, None, mat_n, mat_k, layer_mat, layer_d, (wave_type,),
                               cos_theta=np.array([np.cos(theta)]), want=(), want_matrix=True, flags=self._flags())
        return _to_numpy(res.M[0, 0, :, 0])

    def _chain_cached(self):
        """The reference's literal dataflow: chain whatever matrices the layers cached last (tm.py:559-568)."""
        import torch

        caches = [layer._cache_device() for layer in self.layers]
        if any(c is None for c in caches):
            raise IndexError("layer matrices have not been calculated (calculate_layer_matrices / initialize_struct)")
        n_wl = len(self.wavelengths)
        if any(c.shape[0] < n_wl for c in caches):
            raise IndexError("index out of bounds: a layer caches fewer wavelengths than the structure holds")
        first_inv = caches[0][:n_wl, 2].contiguous()
        last_d = caches[-1][:n_wl, 0].contiguous()
        interior = torch.stack([c[:n_wl, 3] for c in caches[1:-1]]) if len(caches) > 2 else None
        return _to_numpy(_eng().chain_matrices(first_inv, interior, last_d))

    # ---- sweeps --------------------------------------------------------------------------------------------
    def one_job(self, theta, wave_type):
        """(T, R) spectra for one angle (tm.py:574-595)."""
        return self.calculate_all_t_r(self.calculate_all_transfer_matrices(theta, wave_type))

    def angle_resolved_spectra(self, wave_type):
        """[(T_array, R_array) for each angle in self.theta] (tm.py:597-628).  The reference fans the angles out
        over a multiprocessing pool; here all angles are one kernel launch."""
        if self.strict_reference_cache:
            T, R = self.one_job(0.0, wave_type)
            return [(T.copy(), R.copy()) for _ in np.atleast_1d(self.theta)]
        out = self.spectrum_grid(np.atleast_1d(self.theta), (wave_type,))
        T, R = out["T"][0], out["R"][0]
        return [(np.ascontiguousarray(T[:, i]), np.ascontiguousarray(R[:, i])) for i in range(T.shape[1])]

    def spectrum_grid(self, theta=None, pols=(S_WAVE, P_WAVE), sweep_layer=-1, sweep_d=None, want=("R", "T", "A")):
        """Additive API: R, T, A over the whole (pol, wavelength, theta[, thickness]) grid in one launch.
        Returns numpy arrays shaped (n_pol, n_wl, n_theta), or (n_sweep, n_pol, n_wl, n_theta) with a sweep."""
        theta = np.atleast_1d(np.asarray(self.theta if theta is None else theta, dtype=np.float64))
        wl, mat_n, mat_k, layer_mat, layer_d = self._stack_on_grid()
        res = _eng().eval_grid(wl, theta, mat_n, mat_k, layer_mat, layer_d, pols, cos_theta=np.cos(theta),
                               sweep_layer=sweep_layer, sweep_d=sweep_d, want=want, flags=self._flags()).numpy()
        if sweep_d is None:
            res = {k: (v[0] if v is not None else None) for k, v in res.items()}
        return res

    def field_amplitudes(self, theta=None, pols=(S_WAVE, P_WAVE)):
        """Additive API (the reference's YAML carries unused A0/B0 keys): forward and backward
        amplitudes (A_j, B_j) at the left boundary of every layer for a unit incident amplitude -- layer 0 holds (1, r),
        the exit medium (t, 0).  Returns a complex128 array (n_pol, n_wl, n_theta, n_layers, 2)."""
        theta = np.atleast_1d(np.asarray(self.theta if theta is None else theta, dtype=np.float64))
        wl, mat_n, mat_k, layer_mat, layer_d = self._stack_on_grid()
        return _to_numpy(_eng().field_amplitudes(wl, theta, mat_n, mat_k, layer_mat, layer_d, pols,
                                                 cos_theta=np.cos(theta), flags=self._flags()))

    def cavity_length_from_peaks(self, n_eff=1.0, theta=None, pol=S_WAVE, height=None, distance=None, percent=True,
                                 sweep_layer=-1, sweep_d=None, max_peaks=64):
        """Additive API: the reference's "Cavity Length Determiner" notebook fused behind the
        spectrum.  Transmittance spectra are evaluated on the device, searched for peaks on the ascending-wavenumber
        axis nu = 1e4/lambda[um] with scipy.signal.find_peaks(T, height=, distance=) semantics (notebook cell 3, where
        T is in percent), and only the peak lists leave the GPU.  Returns a dict with `peak_wavenumber` (list per
        spectrum), `count`, `mean_spacing` and `L_cav = 1e4 / (2 n_eff mean_spacing)` (cells 5-6), each shaped
        (n_theta,) or (n_sweep, n_theta)."""
        theta = np.atleast_1d(np.asarray(self.theta if theta is None else theta, dtype=np.float64))
        wl, mat_n, mat_k, layer_mat, layer_d = self._stack_on_grid()
        eng = _eng()
        res = eng.eval_grid(wl, theta, mat_n, mat_k, layer_mat, layer_d, (pol,), cos_theta=np.cos(theta),
                            sweep_layer=sweep_layer, sweep_d=sweep_d, want=("T",), flags=self._flags())
        nu = 1.0e4 / wl
        h = None if height is None else (float(height) / 100.0 if percent else float(height))
        idx, hts, cnt, spc = eng.find_peaks(res.T[:, 0], height=h, distance=distance, max_peaks=max_peaks, axis=nu,
                                            reverse=bool(wl.size > 1 and wl[0] < wl[-1]))
        idx, cnt, spc = _to_numpy(idx), _to_numpy(cnt), _to_numpy(spc)
        if (cnt < 0).any():
            raise ValueError("more peaks than max_peaks in at least one spectrum")
        order = nu[::-1] if (wl.size > 1 and wl[0] < wl[-1]) else nu
        peaks = [[order[idx[s, t, :cnt[s, t]]] for t in range(idx.shape[1])] for s in range(idx.shape[0])]
        with np.errstate(divide="ignore", invalid="ignore"):
            out = {"peak_wavenumber": peaks, "count": cnt, "mean_spacing": spc, "L_cav": 1.0e4 / (2.0 * n_eff * spc)}
        if sweep_d is None:
            out = {k: v[0] for k, v in out.items()}
        return out

    def polariton_dispersion(self, theta=None, pol=S_WAVE, height=None, distance=None, sweep_layer=-1, sweep_d=None,
                             max_peaks=64, fit="joint", window=None, max_iter=400, tol=1.5e-8):
        """Additive API: the dispersion step of the reference's "Polariton Peak Analysis"
        notebook (cells 8-17: a two-Lorentzian fit per angle whose centres are the lower and upper polariton) fused
        behind the spectrum.  Transmittance spectra are evaluated on the device for every angle (and swept thickness),
        the two highest peaks on the wavenumber axis nu = 1e4/lambda[um] are located (find_peaks semantics, `height`
        as a fraction, `distance` in samples), estimated by the closed-form three-point Lorentzian and -- fit="joint",
        the default -- refined by the notebook's own fit: the sum of two lmfit-style Lorentzians least-squares fitted
        to the spectrum over `window` = (nu_lo, nu_hi) in cm^-1 (default: the whole grid; the notebook masks 500..8000),
        Levenberg-Marquardt on the device (pst_fit_two_lorentzians) stopped at lmfit's default tolerance `tol` (ftol =
        xtol = 1.5e-8; `evaluations` < 0 marks spectra that used all `max_iter` evaluations first -- typically a window
        that does not hold two lines).  fit="three-point" keeps the closed-form estimates.
        Only (LP, UP, widths, amplitudes) per spectrum leave the GPU.  Returns a dict of arrays shaped (n_theta,) or
        (n_sweep, n_theta): `theta`, `LP`, `UP`, `splitting` = UP - LP, `hwhm_LP`, `hwhm_UP`, `amp_LP`, `amp_UP`
        (peak heights for "three-point", lmfit amplitudes = areas for "joint", which also returns `chi2` and
        `evaluations`), and `min_splitting` (the Rabi-splitting estimate: the smallest UP - LP over the angles) with its
        `theta_at_min`."""
        if fit not in ("joint", "three-point"):
            raise ValueError("fit must be 'joint' or 'three-point'")
        theta = np.atleast_1d(np.asarray(self.theta if theta is None else theta, dtype=np.float64))
        wl, mat_n, mat_k, layer_mat, layer_d = self._stack_on_grid()
        eng = _eng()
        res = eng.eval_grid(wl, theta, mat_n, mat_k, layer_mat, layer_d, (pol,), cos_theta=np.cos(theta),
                            sweep_layer=sweep_layer, sweep_d=sweep_d, want=("T",), flags=self._flags())
        nu = 1.0e4 / wl
        rev = bool(wl.size > 1 and wl[0] < wl[-1])
        T = res.T[:, 0]
        idx, hts, cnt, _ = eng.find_peaks(T, height=height, distance=distance, max_peaks=max_peaks, axis=nu, reverse=rev)
        if (_to_numpy(cnt) < 0).any():
            raise ValueError("more peaks than max_peaks in at least one spectrum")
        seeds = eng.polariton_branches(T, idx, hts, cnt, nu, reverse=rev)
        br = _to_numpy(seeds)
        extra = {}
        if fit == "joint":
            win = None
            if window is not None:  # wavenumber bounds -> sample window of the scan order (ascending wavenumber)
                order = nu[::-1] if rev else nu
                lo, hi = np.searchsorted(order, [float(window[0]), float(window[1])], side="left")
                win = (int(lo), int(max(hi, lo + 6)))
            br = _to_numpy(eng.fit_two_lorentzians(T, seeds, nu, reverse=rev, window=win, max_iter=max_iter, tol=tol))
            extra = {"chi2": br[..., 6], "evaluations": br[..., 7]}
        out = {"theta": theta, "LP": br[..., 0], "UP": br[..., 1], "splitting": br[..., 1] - br[..., 0],
               "hwhm_LP": br[..., 2], "hwhm_UP": br[..., 3], "amp_LP": br[..., 4], "amp_UP": br[..., 5], **extra}
        with np.errstate(invalid="ignore"):
            spl = np.where(np.isnan(out["splitting"]), np.inf, out["splitting"])
        k = spl.argmin(axis=-1)
        out["min_splitting"] = np.take_along_axis(spl, k[..., None], axis=-1)[..., 0]
        out["theta_at_min"] = theta[k]
        if sweep_d is None:
            out = {key: (v[0] if key != "theta" else v) for key, v in out.items()}
        return out

    # ---- misc -----------------------------------------------------------------------------------------------
    def print_structure(self):
        """Print one line per layer, format of tm.py:630-647."""
        print("Structure Configuration")
        for i, layer in enumerate(self.layers):
            print("Layer {} :".format(i), layer.material, "| d = " + str(int(layer.thickness * 10 ** 3)) + " nm")


    def make_sim_dir(self, out_path, sim_name=None):
        """Path of the output directory of an angle-resolved run (tm.py:648-684).  With sim_name = None the name is
        built from the structure as the reference intends: the layers' materials joined by '_', then
        '<theta_i>--<theta_f>deg_<lambda_i>--<lambda_f>lambda/' with the angles rounded to whole degrees.
        The reference's version cannot run (`sim_name` is never bound, tm.py:669, and the else-branch joins a None,
        tm.py:682); this is that function with `sim_name` as the keyword argument its docstring describes.
        Returns the path; creating the directory is the caller's job, as in tm.py:791-793."""
        if sim_name is None:
            sim_name = ""
            for layer in self.layers:
                sim_name += layer.material + "_"
            theta_i = np.round(np.rad2deg(self.theta[0]))
            theta_f = np.round(np.rad2deg(self.theta[-1]))
            lmbda_i = self.wavelengths[0]
            lmbda_f = self.wavelengths[-1]
            sim_name += str(theta_i) + "--" + str(theta_f) + "deg_" + str(lmbda_i) + "--" + str(lmbda_f) + "lambda/"
        return os.path.join(out_path, sim_name)


def write_tmm_results(theta, wavelengths, trans, refl, output_dir, notebook_names=False):
    """CSV `Wavelength,Transmittance,Reflectance`, file name '<theta>deg_.csv' (tm.py:687-726).
    notebook_names = True writes 'deg<theta in degrees, two decimals>_.csv' instead, the 'deg##.##_' pattern the
    reference's analysis notebooks look for (Polariton Peak Analysis.ipynb, cell 3), with theta taken in radians."""
    name = "deg{:05.2f}_.csv".format(float(np.rad2deg(theta))) if notebook_names else str(theta) + "deg_.csv"
    path = os.path.join(output_dir, name)
    with open(path, "w", encoding="utf-8", newline="") as fh:
        w = csv.writer(fh, delimiter=",")
        w.writerow(["Wavelength", "Transmittance", "Reflectance"])
        for i in range(len(trans)):
            w.writerow([wavelengths[i], trans[i], refl[i]])


def fresnel(n1, n2, k1x, k2x):
    """Fresnel (r_s, t_s, r_p, t_p) from indices and normal wavevector components (tm.py:730-744)."""
    rt, _, shape = _eng().fresnel(n1, n2, k1x, k2x)
    rt = _to_numpy(rt)
    outs = tuple(rt[:, i].reshape(shape) for i in range(4))
    if np.ndim(n1) == 0 and np.ndim(n2) == 0 and np.ndim(k1x) == 0 and np.ndim(k2x) == 0:
        outs = tuple(o.reshape(())[()] for o in outs)
    return outs


def transmission_matrix(r_ij, t_ij):
    """Interface matrix (1/t)[[1, r], [r, 1]] (tm.py:746-751)."""
    return _to_numpy(_eng().interface_matrices(r_ij, t_ij)[0])


# ---- command line (tm.py:755-805) ---------------------------------------------------------------------------------
def parse_arguments(argv=None):
    """`structure output [-p | -s]`, the reference's arguments (tm.py:755-767)."""
    parser = argparse.ArgumentParser(prog="python -m pistachio_b200")
    parser.add_argument("structure", help="Path for a yaml file from config_files describing a multilayered structure.")
    parser.add_argument("output", help="Directory for transfer matrix results.")
    parser.add_argument("-p", "--pwave", help="Boolean. Incident wave is p-wave.", action="store_true")
    parser.add_argument("-s", "--swave", help="Boolean. Incident was is s-wave.", action="store_true")
    parser.add_argument("--notebook-names", action="store_true",
                        help="name the per-angle files deg##.##_.csv (degrees) as the reference's notebooks expect")
    return parser.parse_args(argv)


def main(args):
    """Load the YAML structure, evaluate every angle, write one CSV per angle into make_sim_dir(output)
    (tm.py:770-799; there the run always dies in make_sim_dir after computing).  Like the reference it exits
    silently when neither -p nor -s is given (tm.py:779-780)."""
    print("Loading structure parameters from {}".format(args.structure))
    if args.pwave:
        wave_type = P_WAVE
    elif args.swave:
        wave_type = S_WAVE
    else:
        sys.exit()
    s = Structure()
    s.load_struct_from_config(args.structure)
    start_time = time.time()
    results = s.angle_resolved_spectra(wave_type)
    elapsed_time = np.round(time.time() - start_time, 4)
    print("Calculation time: {} seconds".format(elapsed_time))
    out_path = s.make_sim_dir(args.output)
    if not os.path.exists(out_path):
        os.makedirs(out_path)
    for i, (t_results, r_results) in enumerate(results):
        write_tmm_results(s.theta[i], s.wavelengths, t_results, r_results, out_path,
                          notebook_names=getattr(args, "notebook_names", False))
    print("Wrote results to {}".format(out_path))
    return out_path


if __name__ == "__main__":
    main(parse_arguments())
