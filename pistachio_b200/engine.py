"""Batched device API over the C ABI: grids in, (R, T, A[, M]) out.

PyTorch appears here only as the carrier of device buffers and streams
(`tensor.data_ptr()`, `torch.cuda.current_stream()`); every numeric step runs in
libpistachio_b200.so.  Result layout everywhere: [sweep][pol][wavelength][theta].
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import _cabi
from ._cabi import (FLAG_FORCE_DEEP_KERNEL, FLAG_FORCE_POINT_KERNEL, FLAG_NO_LAYER_CSE, FLAG_NO_SMEM_STAGING,  # noqa: F401
                    FLAG_NO_SWEEP_REUSE, FLAG_NUMERIC_DET, FLAG_SNELL, POL_P,
                    POL_S, GridArgs, check)

S_WAVE = "s-wave"
P_WAVE = "p-wave"
_POL_CODE = {S_WAVE: POL_S, P_WAVE: POL_P}


def pol_mask(pols) -> int:
    if isinstance(pols, str):
        pols = (pols,)
    mask = 0
    for p in pols:
        if p not in _POL_CODE:
            raise ValueError("Must specify 's-wave' or 'p-wave'.")
        mask |= _POL_CODE[p]
    return mask


def pols_of_mask(mask: int):
    return tuple(p for p, c in ((S_WAVE, POL_S), (P_WAVE, POL_P)) if mask & c)


@dataclass
class GridResult:
    """Device-resident results; each tensor has shape (n_sweep, n_pol, n_wl, n_theta) (+ (2, 2) for M)."""

    R: torch.Tensor | None
    T: torch.Tensor | None
    A: torch.Tensor | None
    M: torch.Tensor | None
    pols: tuple

    def numpy(self):
        return {k: (v.cpu().numpy() if v is not None else None) for k, v in (("R", self.R), ("T", self.T), ("A", self.A), ("M", self.M))}


def _dev_f64(x, device) -> torch.Tensor:
    if isinstance(x, torch.Tensor):
        t = x.to(device=device, dtype=torch.float64)
    else:
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.float64))).to(device)
    return t.contiguous()


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


class TMMEngine:
    """One engine per process and GPU.  Not thread-safe (the handle owns scratch memory)."""

    def __init__(self, device: int | None = None):
        if not torch.cuda.is_available():
            raise _cabi.PistachioError("no CUDA device is visible; pistachio_b200 has no CPU fallback")
        self.device_index = torch.cuda.current_device() if device is None else int(device)
        self.device = torch.device("cuda", self.device_index)
        self.handle = _cabi.default_handle(self.device_index)
        self.lib = self.handle.lib

    # ------------------------------------------------------------------ stream helper
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def launch_count(self) -> int:
        return self.handle.launch_count()

    # ------------------------------------------------------------------ K0
    def resample_tables(self, tables, wl):
        """tables: sequence of (tab_wl, tab_n, tab_k) host arrays.  Returns (mat_n, mat_k) device tensors
        [n_tab][n_wl].  Reference: Layer.make_datapoints, transfer_matrix.py:117-170 (interp1d's stable sort of the
        table is host bookkeeping; the search and the arithmetic run on the device)."""
        wl_d = _dev_f64(wl, self.device)
        offs = [0]
        cw, cn, ck = [], [], []
        for tw, tn, tk in tables:
            tw = np.asarray(tw, dtype=np.float64).ravel()
            tn = np.asarray(tn, dtype=np.float64).ravel()
            tk = np.asarray(tk, dtype=np.float64).ravel()
            if not (len(tn) == len(tk)):
                raise ValueError("refractive_index and extinction_coeff tables differ in length")
            if len(tn) == 0:
                raise ValueError("empty refractive index table")
            if len(tn) > 1:
                if len(tw) != len(tn):
                    raise ValueError("x and y arrays must be equal in length along interpolation axis.")
                order = np.argsort(tw, kind="mergesort")
                tw, tn, tk = tw[order], tn[order], tk[order]
            else:
                tw = np.zeros(1)
            cw.append(tw); cn.append(tn); ck.append(tk)
            offs.append(offs[-1] + len(tn))
        n_tab = len(cw)
        tw_d = _dev_f64(np.concatenate(cw), self.device)
        tn_d = _dev_f64(np.concatenate(cn), self.device)
        tk_d = _dev_f64(np.concatenate(ck), self.device)
        out_n = torch.empty((n_tab, wl_d.numel()), dtype=torch.float64, device=self.device)
        out_k = torch.empty_like(out_n)
        off_arr = (C.c_int * (n_tab + 1))(*offs)
        check(self.lib.pst_resample_tables(self.handle.raw, n_tab, off_arr, _ptr(tw_d), _ptr(tn_d), _ptr(tk_d), _ptr(wl_d),
                                           wl_d.numel(), _ptr(out_n), _ptr(out_k), self._stream()), "pst_resample_tables")
        return out_n, out_k

    # ------------------------------------------------------------------ K1/K2
    def _grid_args(self, wl, theta, mat_n, mat_k, layer_mat, layer_d, pols, sweep_layer, sweep_ptr, n_sweep, cos_theta,
                   flags, n_wl, n_theta, n_mat, outs):
        # numpy does the list -> C array conversion (a 10 000-layer list through ctypes costs ~1 ms per call)
        lm = np.ascontiguousarray(layer_mat, dtype=np.int32)
        ld = np.ascontiguousarray(layer_d, dtype=np.float64)
        if lm.ndim != 1 or lm.shape != ld.shape:
            raise ValueError("layer_mat and layer_d differ in length")
        L = lm.shape[0]
        a = GridArgs()
        a.struct_size = C.sizeof(GridArgs)
        a.n_wl, a.n_theta, a.n_mat, a.n_layers = n_wl, n_theta, n_mat, L
        a.sweep_layer = int(sweep_layer)
        a.n_sweep = int(n_sweep)
        a.pol_mask = pol_mask(pols)
        a.flags = int(flags)
        keep = [lm, ld]
        a.layer_mat = lm.ctypes.data_as(_cabi.c_int_p)
        a.layer_d = ld.ctypes.data_as(_cabi.c_double_p)
        a.wl, a.theta, a.cos_theta = wl, theta, cos_theta
        a.mat_n, a.mat_k, a.sweep_d = mat_n, mat_k, sweep_ptr
        a.R, a.T, a.A, a.M = outs
        return a, keep

    def eval_grid(self, wl, theta, mat_n, mat_k, layer_mat, layer_d, pols=(S_WAVE, P_WAVE), *, sweep_layer=-1,
                  sweep_d=None, cos_theta=None, want=("R", "T", "A"), want_matrix=False, flags=0, out=None,
                  out_ptrs=None) -> GridResult:
        """Fused build + chain + R/T/A for every (sweep, pol, wavelength, theta) point (pst_eval_grid).

        wl, theta[, cos_theta], sweep_d: 1-D float64 (host arrays are uploaded; CUDA tensors are used in place).
        mat_n, mat_k: [n_mat][n_wl] on the grid.  layer_mat / layer_d: per layer material column and thickness."""
        wl_d = _dev_f64(wl, self.device)
        th_d = _dev_f64(theta, self.device) if theta is not None else None
        ct_d = _dev_f64(cos_theta, self.device) if cos_theta is not None else None
        n_d = _dev_f64(mat_n, self.device)
        k_d = _dev_f64(mat_k, self.device)
        if n_d.dim() != 2 or n_d.shape != k_d.shape or n_d.shape[1] != wl_d.numel():
            raise ValueError("mat_n / mat_k must have shape (n_mat, n_wl)")
        sw_d = _dev_f64(sweep_d, self.device) if sweep_d is not None else None
        n_wl, n_mat = wl_d.numel(), n_d.shape[0]
        n_theta = (th_d if th_d is not None else ct_d).numel()
        mask = pol_mask(pols)
        n_pol = bin(mask).count("1")
        n_sw = 1 if sw_d is None else sw_d.numel()
        shape = (n_sw, n_pol, n_wl, n_theta)
        if out_ptrs is not None:
            # raw device addresses for R, T, A (e.g. a slab inside an NVLink multicast mapping with
            # PST_FLAG_MULTICAST_OUT, see dist.AssembledResult); the caller owns the memory and its extent
            if want_matrix:
                raise ValueError("out_ptrs does not cover the matrix output")
            peers = None
            if any(isinstance(out_ptrs.get(k), (list, tuple)) for k in ("R", "T", "A")):
                # one address per receiving GPU: unicast fan-out by the kernel (pst_peer_outputs)
                peers = _cabi.PeerOutputs()
                n = max(len(v) for v in out_ptrs.values() if isinstance(v, (list, tuple)))
                if n > _cabi.MAX_PEERS:
                    raise ValueError(f"at most {_cabi.MAX_PEERS} peer buffers")
                peers.n_peers = n
                for k in ("R", "T", "A"):
                    lst = list(out_ptrs.get(k) or [])
                    for i in range(n):
                        getattr(peers, k)[i] = int(lst[i]) if i < len(lst) and lst[i] else None
                raw = (None, None, None, None)
            else:
                raw = tuple(C.c_void_p(int(out_ptrs[k])) if out_ptrs.get(k) else None for k in ("R", "T", "A")) + (None,)
            a, keep = self._grid_args(_ptr(wl_d), _ptr(th_d), _ptr(n_d), _ptr(k_d), layer_mat, layer_d, pols, sweep_layer,
                                      _ptr(sw_d), n_sw, _ptr(ct_d), flags, n_wl, n_theta, n_mat, raw)
            if peers is not None:
                a.peers = C.pointer(peers)
            check(self.lib.pst_eval_grid(self.handle.raw, C.byref(a), self._stream()), "pst_eval_grid")
            del keep
            return GridResult(None, None, None, None, pols_of_mask(mask))
        res = {}
        for key in ("R", "T", "A"):
            if key in want:
                t = None if out is None else out.get(key)
                if t is None:
                    t = torch.empty(shape, dtype=torch.float64, device=self.device)
                elif tuple(t.shape) != shape or t.dtype != torch.float64 or not t.is_contiguous() or not t.is_cuda:
                    raise ValueError(f"out[{key!r}] must be a contiguous float64 CUDA tensor of shape {shape}")
                res[key] = t
            else:
                res[key] = None
        M = None
        if want_matrix:
            M = None if out is None else out.get("M")
            if M is None:
                M = torch.empty(shape + (2, 2), dtype=torch.complex128, device=self.device)
        outs = (_ptr(res["R"]), _ptr(res["T"]), _ptr(res["A"]), _ptr(M))
        a, keep = self._grid_args(_ptr(wl_d), _ptr(th_d), _ptr(n_d), _ptr(k_d), layer_mat, layer_d, pols, sweep_layer,
                                  _ptr(sw_d), n_sw, _ptr(ct_d), flags, n_wl, n_theta, n_mat, outs)
        check(self.lib.pst_eval_grid(self.handle.raw, C.byref(a), self._stream()), "pst_eval_grid")
        del keep
        return GridResult(res["R"], res["T"], res["A"], M, pols_of_mask(mask))

    def field_amplitudes(self, wl, theta, mat_n, mat_k, layer_mat, layer_d, pols=(S_WAVE, P_WAVE), *, cos_theta=None,
                         flags=0) -> torch.Tensor:
        """Forward/backward amplitudes (A_j, B_j) at the left boundary of every layer, normalised to a unit incident
        amplitude (pst_field_amplitudes).  Returns a complex128 CUDA tensor (n_pol, n_wl, n_theta, n_layers, 2)."""
        wl_d = _dev_f64(wl, self.device)
        th_d = _dev_f64(theta, self.device) if theta is not None else None
        ct_d = _dev_f64(cos_theta, self.device) if cos_theta is not None else None
        n_d = _dev_f64(mat_n, self.device)
        k_d = _dev_f64(mat_k, self.device)
        if n_d.dim() != 2 or n_d.shape != k_d.shape or n_d.shape[1] != wl_d.numel():
            raise ValueError("mat_n / mat_k must have shape (n_mat, n_wl)")
        n_wl, n_mat = wl_d.numel(), n_d.shape[0]
        n_theta = (th_d if th_d is not None else ct_d).numel()
        n_pol = bin(pol_mask(pols)).count("1")
        amp = torch.empty((n_pol, n_wl, n_theta, len(layer_d), 2), dtype=torch.complex128, device=self.device)
        a, keep = self._grid_args(_ptr(wl_d), _ptr(th_d), _ptr(n_d), _ptr(k_d), layer_mat, layer_d, pols, -1, None, 1,
                                  _ptr(ct_d), flags, n_wl, n_theta, n_mat, (None, None, None, None))
        check(self.lib.pst_field_amplitudes(self.handle.raw, C.byref(a), _ptr(amp), self._stream()), "pst_field_amplitudes")
        del keep
        return amp

    def eval_grid_host(self, wl, theta, mat_n, mat_k, layer_mat, layer_d, pols=(S_WAVE, P_WAVE), *, sweep_layer=-1,
                       sweep_d=None, cos_theta=None, want=("R", "T", "A"), flags=0, out=None):
        """Host buffers in, host buffers out (pst_eval_grid_host): the plugin-style call.  `out` may carry pinned
        torch CPU tensors / numpy arrays to receive R, T, A; returns a dict of numpy views."""
        def host(x):
            if x is None:
                return None
            if isinstance(x, torch.Tensor):
                assert x.device.type == "cpu" and x.dtype == torch.float64 and x.is_contiguous()
                return x
            return torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.float64)))

        wl_h, th_h, ct_h = host(wl), host(theta), host(cos_theta)
        n_h, k_h, sw_h = host(mat_n), host(mat_k), host(sweep_d)
        n_wl, n_mat = wl_h.numel(), n_h.shape[0]
        n_theta = (th_h if th_h is not None else ct_h).numel()
        mask = pol_mask(pols)
        shape = (1 if sw_h is None else sw_h.numel(), bin(mask).count("1"), n_wl, n_theta)
        res = {}
        for key in ("R", "T", "A"):
            if key in want:
                t = None if out is None else out.get(key)
                if t is None:
                    t = torch.empty(shape, dtype=torch.float64)
                elif not isinstance(t, torch.Tensor):
                    t = torch.from_numpy(t)
                if tuple(t.shape) != shape or t.dtype != torch.float64 or not t.is_contiguous():
                    raise ValueError(f"out[{key!r}] must be a contiguous float64 tensor of shape {shape}")
                res[key] = t
            else:
                res[key] = None
        a, keep = self._grid_args(_ptr(wl_h), _ptr(th_h), _ptr(n_h), _ptr(k_h), layer_mat, layer_d, pols, sweep_layer,
                                  _ptr(sw_h), shape[0], _ptr(ct_h), flags, n_wl, n_theta, n_mat,
                                  (_ptr(res["R"]), _ptr(res["T"]), _ptr(res["A"]), None))
        with torch.cuda.device(self.device):
            check(self.lib.pst_eval_grid_host(self.handle.raw, C.byref(a)), "pst_eval_grid_host")
        del keep
        return {k: (v.numpy() if v is not None else None) for k, v in res.items()}

    # ------------------------------------------------------------------ workloads
    def prepare_workload(self, w):
        """Device-side setup of a `workloads.Workload`: deduplicate the layers' tables into materials, resample them
        onto the grid with K0, and return everything `eval_grid` needs (all tensors already in HBM)."""
        seen, tables, layer_mat = {}, [], []
        for l in w.layers:
            sig = (np.asarray(l.tab_wl).tobytes(), np.asarray(l.tab_n).tobytes(), np.asarray(l.tab_k).tobytes())
            if sig not in seen:
                seen[sig] = len(tables)
                tables.append((l.tab_wl, l.tab_n, l.tab_k))
            layer_mat.append(seen[sig])
        wl_d = _dev_f64(w.wl, self.device)
        mat_n, mat_k = self.resample_tables(tables, wl_d)
        prep = {
            "wl": wl_d,
            "theta": _dev_f64(w.theta, self.device),
            "mat_n": mat_n,
            "mat_k": mat_k,
            "layer_mat": np.asarray(layer_mat, dtype=np.int32),
            "layer_d": np.asarray([l.thickness for l in w.layers], dtype=np.float64),
            "sweep_layer": w.sweep_layer,
            "sweep_d": _dev_f64(w.sweep_d, self.device) if w.sweep_d is not None else None,
        }
        return prep

    def eval_prepared(self, prep, pols=(S_WAVE, P_WAVE), **kw) -> GridResult:
        return self.eval_grid(prep["wl"], prep["theta"], prep["mat_n"], prep["mat_k"], prep["layer_mat"], prep["layer_d"],
                              pols, sweep_layer=prep["sweep_layer"], sweep_d=prep["sweep_d"], **kw)

    # ------------------------------------------------------------------ API-parity kernels
    def layer_matrices(self, n_complex, wl, theta, pol, thickness, cos_theta=None):
        """[D, P, D^-1, M] per wavelength -> complex128 tensor (n_wl, 4, 2, 2) on the device
        (pst_layer_matrices; reference transfer_matrix.py:236-305)."""
        if pol not in _POL_CODE:
            raise ValueError("Must specify 's-wave' or 'p-wave'.")
        n = np.asarray(n_complex, dtype=np.complex128).ravel()
        wl_d = _dev_f64(np.asarray(wl, dtype=np.float64).ravel(), self.device)
        if n.size != wl_d.numel():
            raise ValueError("refractive index and wavelength arrays differ in length")
        nr = _dev_f64(n.real, self.device)
        ni = _dev_f64(n.imag, self.device)
        out = torch.empty((n.size, 4, 2, 2), dtype=torch.complex128, device=self.device)
        ct = float(np.cos(theta)) if cos_theta is None else float(cos_theta)
        check(self.lib.pst_layer_matrices(self.handle.raw, _ptr(nr), _ptr(ni), _ptr(wl_d), n.size, float(theta), ct,
                                          _POL_CODE[pol], float(thickness), _ptr(out), self._stream()), "pst_layer_matrices")
        return out

    def propagation_matrices(self, kx, thickness):
        kx = np.asarray(kx, dtype=np.complex128).ravel()
        kx_d = torch.from_numpy(kx).to(self.device)
        out = torch.empty((kx.size, 2, 2), dtype=torch.complex128, device=self.device)
        d = complex(thickness)
        check(self.lib.pst_propagation_matrices(self.handle.raw, _ptr(kx_d), kx.size, d.real, d.imag, _ptr(out),
                                                self._stream()), "pst_propagation_matrices")
        return out

    def wavenumbers(self, n_complex, wl, theta):
        n = np.asarray(n_complex, dtype=np.complex128).ravel()
        wl_d = _dev_f64(np.asarray(wl, dtype=np.float64).ravel(), self.device)
        if n.size != wl_d.numel():
            raise ValueError("operands could not be broadcast together: refractive index vs wavelengths")
        nr, ni = _dev_f64(n.real, self.device), _dev_f64(n.imag, self.device)
        kx = torch.empty(n.size, dtype=torch.complex128, device=self.device)
        kz = torch.empty_like(kx)
        check(self.lib.pst_wavenumbers(self.handle.raw, _ptr(nr), _ptr(ni), _ptr(wl_d), n.size, float(np.cos(theta)),
                                       float(np.sin(theta)), _ptr(kx), _ptr(kz), self._stream()), "pst_wavenumbers")
        return kx, kz

    def t_r_from_matrices(self, M):
        """(T, R) device tensors from matrices (n, 2, 2) complex128 (pst_t_r_from_matrices; tm.py:464-514)."""
        if isinstance(M, torch.Tensor):
            M_d = M.to(device=self.device, dtype=torch.complex128).contiguous()
        else:
            M_d = torch.from_numpy(np.ascontiguousarray(np.asarray(M, dtype=np.complex128))).to(self.device)
        M_d = M_d.reshape(-1, 2, 2)
        n = M_d.shape[0]
        T = torch.empty(n, dtype=torch.float64, device=self.device)
        R = torch.empty_like(T)
        if n:
            check(self.lib.pst_t_r_from_matrices(self.handle.raw, _ptr(M_d), n, _ptr(T), _ptr(R), self._stream()),
                  "pst_t_r_from_matrices")
        return T, R

    def chain_matrices(self, first_inv, interior, last_d):
        """M = D_0^-1 (M_1 ... M_{N-2}) D_{N-1} from explicit per-layer matrices (pst_chain_matrices; tm.py:559-568).
        first_inv, last_d: (n_wl, 2, 2); interior: (n_interior, n_wl, 2, 2) complex128 device tensors."""
        n_wl = first_inv.shape[0]
        n_int = 0 if interior is None else interior.shape[0]
        out = torch.empty((n_wl, 2, 2), dtype=torch.complex128, device=self.device)
        check(self.lib.pst_chain_matrices(self.handle.raw, _ptr(first_inv.contiguous()),
                                          _ptr(interior.contiguous()) if n_int else None, _ptr(last_d.contiguous()),
                                          n_int, n_wl, _ptr(out), self._stream()), "pst_chain_matrices")
        return out

    def fresnel(self, n1, n2, k1x, k2x, want_matrices=False):
        arrs = [np.atleast_1d(np.asarray(x, dtype=np.complex128)) for x in (n1, n2, k1x, k2x)]
        arrs = np.broadcast_arrays(*arrs)
        n = arrs[0].size
        dev = [torch.from_numpy(np.ascontiguousarray(a).ravel()).to(self.device) for a in arrs]
        rt = torch.empty((n, 4), dtype=torch.complex128, device=self.device)
        D = torch.empty((n, 2, 2, 2), dtype=torch.complex128, device=self.device) if want_matrices else None
        check(self.lib.pst_fresnel(self.handle.raw, _ptr(dev[0]), _ptr(dev[1]), _ptr(dev[2]), _ptr(dev[3]), n, _ptr(rt),
                                   _ptr(D), self._stream()), "pst_fresnel")
        return rt, D, arrs[0].shape

    def interface_matrices(self, r, t):
        """(1/t) [[1, r], [r, 1]] per element (pst_interface_matrices; reference transfer_matrix.py:746-751)."""
        r = np.atleast_1d(np.asarray(r, dtype=np.complex128)).ravel()
        t = np.atleast_1d(np.asarray(t, dtype=np.complex128)).ravel()
        if r.shape != t.shape:
            raise ValueError("r and t differ in length")
        r_d, t_d = torch.from_numpy(r).to(self.device), torch.from_numpy(t).to(self.device)
        out = torch.empty((r.size, 2, 2), dtype=torch.complex128, device=self.device)
        check(self.lib.pst_interface_matrices(self.handle.raw, _ptr(r_d), _ptr(t_d), r.size, _ptr(out), self._stream()),
              "pst_interface_matrices")
        return out

    # ------------------------------------------------------------------ fused spectral reduction
    def find_peaks(self, y, height=None, distance=None, max_peaks=64, axis=None, reverse=False):
        """Peaks along the wavelength axis of a device result y[..., n_wl, n_theta] (e.g. GridResult.T), with the
        semantics of scipy.signal.find_peaks(spectrum, height=, distance=) (pst_find_peaks).  Returns device tensors
        (idx [..., n_theta, max_peaks] int32, heights, count [..., n_theta] int32, spacing [..., n_theta]); only these
        small arrays need to leave the device.  `axis` (n_wl values, e.g. wavenumbers) enables the mean peak spacing."""
        if not (isinstance(y, torch.Tensor) and y.is_cuda and y.dtype == torch.float64 and y.dim() >= 2):
            raise ValueError("y must be a float64 CUDA tensor [..., n_wl, n_theta]")
        y = y.contiguous()
        n_pts, n_inner = y.shape[-2], y.shape[-1]
        lead = tuple(y.shape[:-2])
        n_outer = int(np.prod(lead)) if lead else 1
        if distance is not None and distance < 1:
            raise ValueError("`distance` must be greater or equal to 1")
        dist_i = 0 if distance is None else int(np.ceil(distance))
        ax_d = _dev_f64(axis, self.device) if axis is not None else None
        if ax_d is not None and ax_d.numel() != n_pts:
            raise ValueError("axis must have one value per wavelength")
        shape = lead + (n_inner,)
        idx = torch.empty(shape + (max_peaks,), dtype=torch.int32, device=self.device)
        hts = torch.empty(shape + (max_peaks,), dtype=torch.float64, device=self.device)
        cnt = torch.empty(shape, dtype=torch.int32, device=self.device)
        spc = torch.empty(shape, dtype=torch.float64, device=self.device)
        check(self.lib.pst_find_peaks(self.handle.raw, _ptr(y), n_outer, n_pts, n_inner,
                                      0.0 if height is None else float(height), 0 if height is None else 1, dist_i,
                                      int(max_peaks), 1 if reverse else 0, _ptr(ax_d), _ptr(idx), _ptr(hts), _ptr(cnt),
                                      _ptr(spc), self._stream()), "pst_find_peaks")
        return idx, hts, cnt, spc

    def polariton_branches(self, y, idx, hts, cnt, axis, reverse=False):
        """Lower/upper polariton branch per spectrum from the peak lists `find_peaks` returned for the same `y`, `axis`
        and `reverse` (pst_polariton_branches): the two highest peaks, each refined by the closed-form three-point
        Lorentzian.  Returns a device tensor [..., n_theta, 6] = (centre_lower, centre_upper, hwhm_lower, hwhm_upper,
        amplitude_lower, amplitude_upper) in units of `axis`; NaN where a spectrum has fewer than two peaks."""
        y = y.contiguous()
        n_pts, n_inner = y.shape[-2], y.shape[-1]
        lead = tuple(y.shape[:-2])
        n_outer = int(np.prod(lead)) if lead else 1
        ax_d = _dev_f64(axis, self.device)
        if ax_d.numel() != n_pts:
            raise ValueError("axis must have one value per wavelength")
        out = torch.empty(lead + (n_inner, 6), dtype=torch.float64, device=self.device)
        check(self.lib.pst_polariton_branches(self.handle.raw, _ptr(y), n_outer, n_pts, n_inner, 1 if reverse else 0,
                                              _ptr(ax_d), _ptr(idx.contiguous()), _ptr(hts.contiguous()),
                                              _ptr(cnt.contiguous()), int(idx.shape[-1]), _ptr(out), self._stream()),
              "pst_polariton_branches")
        return out

    def fit_two_lorentzians(self, y, seeds, axis, reverse=False, window=None, sigma0=None, max_iter=200, tol=1.5e-8):
        """Joint two-Lorentzian least-squares fit of every spectrum of y[..., n_wl, n_theta] (pst_fit_two_lorentzians),
        seeded by the rows `polariton_branches` returned for the same `y`, `axis` and `reverse`.  `window` = (i_lo, i_hi)
        in samples of the scan order (default: the whole spectrum).  Returns a device tensor [..., n_theta, 8] =
        (center_lower, center_upper, sigma_lower, sigma_upper, amplitude_lower, amplitude_upper, chi2, evaluations) in
        lmfit's LorentzianModel convention (amplitude = area, sigma = half width at half maximum).  `tol`: relative
        stopping tolerance on chi2 and on the parameters (MINPACK's ftol = xtol; 1.5e-8 is lmfit's default)."""
        y = y.contiguous()
        n_pts, n_inner = y.shape[-2], y.shape[-1]
        lead = tuple(y.shape[:-2])
        n_outer = int(np.prod(lead)) if lead else 1
        ax_d = _dev_f64(axis, self.device)
        if ax_d.numel() != n_pts:
            raise ValueError("axis must have one value per wavelength")
        i_lo, i_hi = (0, n_pts) if window is None else (int(window[0]), int(window[1]))
        if sigma0 is None:
            sigma0 = 4.0 * float((ax_d.max() - ax_d.min()).item()) / max(n_pts - 1, 1)
        seeds = seeds.contiguous()
        out = torch.empty(lead + (n_inner, 8), dtype=torch.float64, device=self.device)
        check(self.lib.pst_fit_two_lorentzians(self.handle.raw, _ptr(y), n_outer, n_pts, n_inner, 1 if reverse else 0, _ptr(ax_d),
                                               _ptr(seeds), i_lo, i_hi, float(sigma0), float(tol), int(max_iter), _ptr(out),
                                               self._stream()),
              "pst_fit_two_lorentzians")
        return out

    # ------------------------------------------------------------------ measurement
    def fp64_peak_tflops(self, iters: int = 4096) -> float:
        out = C.c_double(0.0)
        check(self.lib.pst_fp64_peak_probe(self.handle.raw, int(iters), C.byref(out), self._stream()), "pst_fp64_peak_probe")
        return float(out.value)

    def flush_l2(self):
        check(self.lib.pst_flush_l2(self.handle.raw, self._stream()), "pst_flush_l2")


_engines = {}


def get_engine(device: int | None = None) -> TMMEngine:
    if not torch.cuda.is_available():
        raise _cabi.PistachioError("no CUDA device is visible; pistachio_b200 has no CPU fallback")
    idx = torch.cuda.current_device() if device is None else int(device)
    eng = _engines.get(idx)
    if eng is None:
        eng = _engines[idx] = TMMEngine(idx)
    return eng
