"""ctypes binding of libpistachio_b200.so (C ABI: include/pistachio_b200.h).

Loading is lazy and loud: if the shared library has not been built, or no B200 is
visible, every compute entry point raises -- there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "libpistachio_b200.so")

ABI_VERSION = 2
POL_S, POL_P = 1, 2
FLAG_FORCE_POINT_KERNEL = 1 << 0
FLAG_FORCE_DEEP_KERNEL = 1 << 1
FLAG_NUMERIC_DET = 1 << 2
FLAG_NO_SMEM_STAGING = 1 << 3
FLAG_NO_LAYER_CSE = 1 << 4
FLAG_NO_SWEEP_REUSE = 1 << 5
FLAG_NO_BULK_COPY = 1 << 6
FLAG_NO_PERIODIC_POWER = 1 << 7
FLAG_MULTICAST_OUT = 1 << 8
FLAG_FP32 = 1 << 9
FLAG_POL_DEGENERATE = 1 << 10
FLAG_NO_SHAPE_FASTPATH = 1 << 11
FLAG_SNELL = 1 << 12

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)


MAX_PEERS = 8


class PeerOutputs(C.Structure):
    """Mirror of `pst_peer_outputs`."""

    _fields_ = [("n_peers", C.c_int), ("R", C.c_void_p * MAX_PEERS), ("T", C.c_void_p * MAX_PEERS), ("A", C.c_void_p * MAX_PEERS)]


class GridArgs(C.Structure):
    """Mirror of `pst_grid_args`."""

    _fields_ = [
        ("struct_size", C.c_size_t),
        ("n_wl", C.c_int),
        ("n_theta", C.c_int),
        ("n_mat", C.c_int),
        ("n_layers", C.c_int),
        ("n_sweep", C.c_int),
        ("sweep_layer", C.c_int),
        ("pol_mask", C.c_uint),
        ("flags", C.c_uint),
        ("wl", C.c_void_p),
        ("theta", C.c_void_p),
        ("cos_theta", C.c_void_p),
        ("mat_n", C.c_void_p),
        ("mat_k", C.c_void_p),
        ("layer_mat", c_int_p),
        ("layer_d", c_double_p),
        ("sweep_d", C.c_void_p),
        ("R", C.c_void_p),
        ("T", C.c_void_p),
        ("A", C.c_void_p),
        ("M", C.c_void_p),
        ("peers", C.POINTER(PeerOutputs)),
    ]


# name -> (restype, argtypes); every symbol include/pistachio_b200.h declares
SYMBOLS = {
    "pst_abi_version": (C.c_int, []),
    "pst_last_error": (C.c_char_p, []),
    "pst_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "pst_destroy": (C.c_int, [C.c_void_p]),
    "pst_launch_count": (C.c_int64, [C.c_void_p]),
    "pst_resample_tables": (C.c_int, [C.c_void_p, C.c_int, c_int_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pst_eval_grid": (C.c_int, [C.c_void_p, C.POINTER(GridArgs), C.c_void_p]),
    "pst_eval_grid_host": (C.c_int, [C.c_void_p, C.POINTER(GridArgs)]),
    "pst_layer_matrices": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_double,
                                      C.c_uint, C.c_double, C.c_void_p, C.c_void_p]),
    "pst_propagation_matrices": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_void_p, C.c_void_p]),
    "pst_wavenumbers": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_double,
                                   C.c_void_p, C.c_void_p, C.c_void_p]),
    "pst_t_r_from_matrices": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pst_chain_matrices": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                      C.c_void_p]),
    "pst_field_amplitudes": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pst_fresnel": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                               C.c_void_p, C.c_void_p]),
    "pst_interface_matrices": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "pst_find_peaks": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int,
                                  C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pst_polariton_branches": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "pst_fit_two_lorentzians": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                           C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_void_p]),
    "pst_fp64_peak_probe": (C.c_int, [C.c_void_p, C.c_int, c_double_p, C.c_void_p]),
    "pst_flush_l2": (C.c_int, [C.c_void_p, C.c_void_p]),
}

_lib = None


class PistachioError(RuntimeError):
    pass


def load_library():
    """dlopen the in-tree shared library and type every entry point."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PistachioError(
            f"{LIB_PATH} is missing: build it with `python -m pistachio_b200.build` "
            "(or __graft_entry__.build()).  pistachio_b200 has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the library lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.pst_abi_version() != ABI_VERSION:
        raise PistachioError("libpistachio_b200.so ABI version does not match the Python binding")
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc == 0:
        return
    msg = load_library().pst_last_error().decode("utf-8", "replace")
    if rc == -1:
        raise ValueError(msg or what)
    raise PistachioError(f"{what or 'pistachio_b200'} failed (code {rc}): {msg}")


class Handle:
    """Owner of one `pst_handle_t` (one per process and device)."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        self.device = int(device)
        h = C.c_void_p()
        check(self.lib.pst_create(self.device, C.byref(h)), "pst_create")
        self._h = h

    @property
    def raw(self):
        if self._h is None:
            raise PistachioError("handle already destroyed")
        return self._h

    def launch_count(self) -> int:
        return int(self.lib.pst_launch_count(self.raw))

    def close(self):
        if getattr(self, "_h", None) is not None:
            self.lib.pst_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_handles = {}


def default_handle(device: int | None = None) -> Handle:
    """Process-wide handle for `device` (default: torch's current CUDA device)."""
    if device is None:
        import torch

        if not torch.cuda.is_available():
            raise PistachioError("no CUDA device is visible; pistachio_b200 has no CPU fallback")
        device = torch.cuda.current_device()
    h = _handles.get(device)
    if h is None or h._h is None:
        h = Handle(device)
        _handles[device] = h
    return h
