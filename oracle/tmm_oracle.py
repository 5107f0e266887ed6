"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.

A from-scratch numpy restatement of the transfer-matrix hot path of
garrekstemo/pistachio (reference file: src/pistachio/transfer_matrix.py,
abbreviated tm.py below).  It exists so that the CUDA path can be checked on
the GPU box, where /root/reference is absent.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import it; the product package (pistachio_b200/) never does.

Parity status: PINNED.  tests/test_oracle_golden.py checks this file
  (a) against fixtures produced by the real reference in the build container
      (oracle/make_golden.py -> tests/golden/reference_cases.npz), bit for bit,
  (b) against the reference's own known-answer tests (tests/test_layer.py:67-91,
      tests/test_structure.py:65-71), and
  (c) against the full-precision arrays stored in the reference tutorial's
      plot outputs (tests/golden/tutorial_goldens.npz).

The arithmetic of the reference lives in numpy/scipy calls on 2x2 arrays
(np.linalg.inv, np.linalg.multi_dot, np.matmul, np.linalg.det, np.exp, np.cos,
scipy.interpolate.interp1d).  The restatement issues the *same library calls in
the same order on the same operand types*, so on one machine it is
bit-identical to the reference; only the Python scaffolding around the calls is
new.  scipy's interp1d is restated in numpy from scipy 1.18.1's linear branch
(the reference pins no scipy version -- see SURVEY.md section 8a-3).
"""
from __future__ import annotations

import numpy as np

C0 = 299792458.0  # scipy.constants.c, used at tm.py:186,269
S_WAVE = "s-wave"
P_WAVE = "p-wave"


# --------------------------------------------------------------------------- material data
def complex_index(n_real, n_imag):
    """n + i*k element by element -> complex128 array (tm.py:52-70)."""
    out = []
    for i in range(len(n_real)):
        out.append(n_real[i] + 1j * n_imag[i])
    return np.array(out)


def resample_table(tab_wl, tab_n, tab_k, wl):
    """Put a layer's (lambda, n, k) table onto the structure's wavelength grid.

    tm.py:117-170.  A one-entry table is broadcast (tm.py:158-162).  Otherwise
    the reference builds scipy interp1d(kind='linear', fill_value='extrapolate')
    objects (tm.py:164-168); scipy 1.18.1 evaluates that branch as
        i  = clip(searchsorted(x, xq, 'left'), 1, len(x)-1)
        y  = ((xq-x[i-1])/(x[i]-x[i-1]))*y[i] + ((x[i]-xq)/(x[i]-x[i-1]))*y[i-1]
    after a stable sort of the table by x (interp1d default assume_sorted=False).
    """
    wl = np.asarray(wl)
    num = len(wl)
    if len(tab_n) == 1:
        n_new = np.full(num, tab_n)
        k_new = np.full(num, tab_k)
        return n_new, k_new
    x = np.array(tab_wl, dtype=np.float64)
    order = np.argsort(x, kind="mergesort")
    x = x[order]
    out = []
    for y in (tab_n, tab_k):
        y = np.array(y, dtype=np.float64)[order]
        idx = np.searchsorted(x, wl).clip(1, len(x) - 1).astype(int)
        lo = idx - 1
        hi = idx
        x_lo = x[lo]
        x_hi = x[hi]
        y_new = ((wl - x_lo) / (x_hi - x_lo)) * y[hi] + ((x_hi - wl) / (x_hi - x_lo)) * y[lo]
        out.append(y_new)
    return out[0], out[1]


def wavenumbers(n_complex, wl, theta=0.0):
    """kx, kz attribute arrays of a layer (tm.py:172-188)."""
    omega = 2 * np.pi * C0 / wl
    kx = n_complex * omega / C0 * np.cos(theta)
    kz = n_complex * omega / C0 * np.sin(theta)
    return kx, kz


# --------------------------------------------------------------------------- per-layer matrices
def dynamical(n_complex, theta, pol):
    """2x2 dynamical matrix D (tm.py:190-215)."""
    if pol == S_WAVE:
        return np.array([[1, 1], [n_complex * np.cos(theta), -n_complex * np.cos(theta)]])
    if pol == P_WAVE:
        return np.array([[np.cos(theta), np.cos(theta)], [n_complex, -n_complex]])
    raise ValueError("polarization must be 's-wave' or 'p-wave'")


def propagation(kx, d):
    """2x2 propagation matrix P = diag(exp(-i kx d), exp(+i kx d)) (tm.py:217-234)."""
    return np.array([[np.exp(-1j * kx * d), 0.0], [0.0, np.exp(1j * kx * d)]])


def layer_matrices(n_complex, lam, theta, pol, d):
    """[D, P, D^-1, M] of one layer at one wavelength (tm.py:236-277)."""
    omega = 2 * np.pi * C0 / lam
    kx = (n_complex * omega / C0) * np.cos(theta)
    D = dynamical(n_complex, theta, pol)
    D_inv = np.linalg.inv(D)
    P = propagation(kx, d)
    M = np.linalg.multi_dot([D, P, D_inv])
    return [D, P, D_inv, M]


def layer_matrices_all(n_array, wl, theta, pol, d):
    """(N_lambda, 4, 2, 2) complex128 cache of one layer (tm.py:279-305)."""
    rows = []
    for i, lam in enumerate(wl):
        rows.append(layer_matrices(n_array[i], lam, theta, pol, d))
    return np.array(rows)


# --------------------------------------------------------------------------- stack
def chain(layer_caches, i):
    """Total matrix at wavelength index i from per-layer caches (tm.py:559-568).

    M = D_0^-1 . (M_1 M_2 ... M_{N-2}) . D_{N-1}; the interior product is
    accumulated from the last interior layer towards the first as M_j @ acc.
    """
    acc = np.identity(2)
    first_inv = layer_caches[0][i][2]
    last_d = layer_caches[-1][i][0]
    for j in range(len(layer_caches))[-2:0:-1]:
        acc = np.matmul(layer_caches[j][i][-1], acc)
    return np.linalg.multi_dot([first_inv, acc, last_d])


def t_r(M):
    """(T, R) of a total matrix (tm.py:464-488): t=1/M00, r=M10/M00."""
    t = 1 / M[0, 0]
    r = M[1, 0] / M[0, 0]
    R = np.real(np.abs(r) ** 2)
    T = np.real(np.linalg.det(M) * np.abs(t) ** 2)
    return T, R


def t_r_all(matrices):
    """(T_array, R_array) for an iterable of 2x2 matrices (tm.py:490-514)."""
    T_out, R_out = [], []
    for m in matrices:
        T, R = t_r(m)
        T_out.append(T)
        R_out.append(R)
    return np.array(T_out), np.array(R_out)


class Stack:
    """Plain container: ordered layers as (n_complex[N_lambda], thickness).

    Index 0 is the incidence medium, index -1 the exit medium; their
    thickness never enters (tm.py:561-563).
    """

    def __init__(self, wl, n_list, d_list):
        self.wl = np.asarray(wl, dtype=np.float64)
        self.n = [np.asarray(n, dtype=np.complex128) for n in n_list]
        self.d = [float(d) for d in d_list]
        assert len(self.n) == len(self.d) >= 2
        for n in self.n:
            assert n.shape == self.wl.shape


def spectrum(stack: Stack, theta, pol, want_matrices=False):
    """Reference call sequence for one (theta, pol):
    calculate_layer_matrices (tm.py:516-534) -> Structure.calculate_all_transfer_matrices
    (tm.py:536-572) -> calculate_all_t_r (tm.py:490-514).  Returns (T, R[, M])."""
    caches = [layer_matrices_all(n, stack.wl, theta, pol, d) for n, d in zip(stack.n, stack.d)]
    mats = [chain(caches, i) for i in range(len(stack.wl))]
    T, R = t_r_all(mats)
    if want_matrices:
        return T, R, np.array(mats)
    return T, R


def spectrum_grid(stack: Stack, thetas, pols=(S_WAVE, P_WAVE)):
    """T, R with shape (n_pol, N_lambda, N_theta) -- the layout of the CUDA path."""
    thetas = np.atleast_1d(np.asarray(thetas, dtype=np.float64))
    T = np.empty((len(pols), len(stack.wl), len(thetas)))
    R = np.empty_like(T)
    for ip, pol in enumerate(pols):
        for it, th in enumerate(thetas):
            T[ip, :, it], R[ip, :, it] = spectrum(stack, th, pol)
    return T, R


# --------------------------------------------------------------------------- unused-by-reference helpers
def fresnel(n1, n2, k1x, k2x):
    """Fresnel r,t for s and p (tm.py:730-744)."""
    r_s = (k1x - k2x) / (k1x + k2x)
    t_s = 2 * k1x / (k1x + k2x)
    r_p = (n1 * n1 * k2x - n2 * n2 * k1x) / (n1 * n1 * k2x + n2 * n2 * k1x)
    t_p = (2 * n1 * n1 * k2x) / (n1 * n1 * k2x + n2 * n2 * k1x)
    return r_s, t_s, r_p, t_p


def transmission_matrix(r_ij, t_ij):
    """Interface matrix (1/t)[[1,r],[r,1]] (tm.py:746-751)."""
    return 1 / t_ij * np.array([[1, r_ij], [r_ij, 1]])


# --------------------------------------------------------------------------- closed-form cross-check (vectorised)
def spectrum_closed_form(stack: Stack, thetas, pols=(S_WAVE, P_WAVE)):
    """Vectorised closed-form evaluation of the same model:
    M_j = [[cos p, -i sin p / a], [-i a sin p, cos p]], a = n cos(theta) | n / cos(theta),
    first column propagated, det M = a_last / a_first.  NOT loop-faithful: agrees
    with `spectrum` to rounding only.  Used for large-size property tests and as
    the fast multi-core CPU baseline in bench.py (`kind: "port-vectorised"`).
    Returns (T, R) with shape (n_pol, N_lambda, N_theta)."""
    thetas = np.atleast_1d(np.asarray(thetas, dtype=np.float64))
    wl = stack.wl[:, None]
    cth = np.cos(thetas)[None, :]
    omega = 2 * np.pi * C0 / wl
    T = np.empty((len(pols), stack.wl.size, thetas.size))
    R = np.empty_like(T)
    for ip, pol in enumerate(pols):
        g = cth if pol == S_WAVE else 1.0 / cth
        nf = stack.n[-1][:, None]
        n0 = stack.n[0][:, None]
        v0 = np.ones((stack.wl.size, thetas.size), dtype=np.complex128) * (cth if pol == P_WAVE else 1.0)
        v1 = nf * cth if pol == S_WAVE else nf * np.ones_like(cth)
        for j in range(len(stack.n) - 2, 0, -1):
            n = stack.n[j][:, None]
            kx = (n * omega / C0) * cth
            ph = kx * stack.d[j]
            c, s = np.cos(ph), np.sin(ph)
            a = n * g
            v0, v1 = c * v0 - 1j * s / a * v1, -1j * a * s * v0 + c * v1
        a0 = n0 * g
        scale = 1.0 if pol == S_WAVE else 1.0 / cth
        m00 = 0.5 * (v0 * scale + v1 / (a0 * (1.0 if pol == S_WAVE else cth)))
        m10 = 0.5 * (v0 * scale - v1 / (a0 * (1.0 if pol == S_WAVE else cth)))
        R[ip] = np.abs(m10 / m00) ** 2
        T[ip] = np.real(nf / n0) * np.abs(1 / m00) ** 2
    return T, R
