"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY (refraction mode, PST_FLAG_SNELL).

Parity status: NO REFERENCE COUNTERPART.  garrekstemo/pistachio keeps the incidence angle in every layer
(src/pistachio/transfer_matrix.py:270, abbreviated tm.py), so this mode cannot be pinned to reference outputs.
It is pinned instead (tests/test_snell.py) to
  (a) the reference-pinned oracle (oracle/tmm_oracle.py) at normal incidence, where both models coincide,
  (b) closed forms: Fresnel's single-interface formulas (Brewster angle, total internal reflection) and the Airy
      formula of one slab,
  (c) an independently formulated textbook implementation (`rta_textbook`: Fresnel interface coefficients +
      propagation, power factors Re(n cos)/Re(n cos), e.g. Yeh, Optical Waves in Layered Media, ch. 5), and
  (d) a 50-digit mpmath evaluation (`truth_point`).

`rta` restates the reference's own matrices with the refracted angle of each layer: the dynamical matrix D_j
(tm.py:190-215), the propagation matrix P_j (tm.py:217-234), M_j = D_j P_j D_j^-1 (tm.py:236-277), the chain
M = D_0^-1 (M_1 ... M_{N-2}) D_{N-1} (tm.py:559-568) and t = 1/M00, r = M10/M00, R = |r|^2, T = Re(det M |t|^2)
(tm.py:482-486), using the same numpy calls.  The relations between layers are the ones the reference lists (unused)
in fresnel()/transmission_matrix() (tm.py:730-751).
"""
from __future__ import annotations

import numpy as np

C0 = 299792458.0
S_WAVE = "s-wave"
P_WAVE = "p-wave"


def layer_cosines(n_list, n0, theta0):
    """cos(theta_j) = sqrt(1 - (n_0 sin(theta_0) / n_j)^2) with the root that makes Im(n_j cos(theta_j)) >= 0
    (forward wave decays; k > 0 absorbs in the reference's sign convention, tm.py:217-234)."""
    s2 = (n0 * np.sin(theta0)) ** 2
    out = []
    for n in n_list:
        c = np.sqrt(1.0 + 0j - s2 / (n * n))
        kz = n * c
        if kz.imag < 0 or (kz.imag == 0 and kz.real < 0):
            c = -c
        out.append(c)
    return out


def dynamical_matrix(n, c, pol):
    """tm.py:190-215 with cos(theta_j) = c."""
    if pol == S_WAVE:
        return np.array([[1, 1], [n * c, -n * c]], dtype=complex)
    return np.array([[c, c], [n, -n]], dtype=complex)


def point(lam, n_list, d_list, theta0, pol):
    """(T, R, M) of one spectral point; n_list complex per layer at this wavelength."""
    omega = 2 * np.pi * C0 / lam
    cs = layer_cosines(n_list, n_list[0], theta0)
    M = np.eye(2, dtype=complex)
    for j in range(len(n_list) - 2, 0, -1):
        kx = n_list[j] * omega / C0 * cs[j]
        D = dynamical_matrix(n_list[j], cs[j], pol)
        P = np.array([[np.exp(-1j * kx * d_list[j]), 0], [0, np.exp(1j * kx * d_list[j])]])
        Mj = np.linalg.multi_dot([D, P, np.linalg.inv(D)])
        M = Mj @ M
    D0 = dynamical_matrix(n_list[0], cs[0], pol)
    Df = dynamical_matrix(n_list[-1], cs[-1], pol)
    M = np.linalg.inv(D0) @ (M @ Df)
    t = 1 / M[0, 0]
    r = M[1, 0] / M[0, 0]
    R = (r * np.conj(r)).real
    # det M = n_f cos(theta_f) / (n_0 cos(theta_0)) analytically; the numeric LU determinant loses eps/T
    det = (n_list[-1] * cs[-1]) / (n_list[0] * cs[0])
    T = (det * t * np.conj(t)).real
    return T, R, M


def rta(wl, n_layers, d_list, thetas, pols=(S_WAVE, P_WAVE)):
    """R, T arrays (n_pol, n_wl, n_theta).  n_layers[j][il]: complex index of layer j at wavelength il."""
    R = np.empty((len(pols), len(wl), len(thetas)))
    T = np.empty_like(R)
    for ip, pol in enumerate(pols):
        for il, lam in enumerate(wl):
            ns = [complex(n[il]) for n in n_layers]
            for it, th in enumerate(thetas):
                T[ip, il, it], R[ip, il, it], _ = point(lam, ns, d_list, th, pol)
    return R, T


# ------------------------------------------------------------------------------ independent textbook formulation
def point_textbook(lam, n_list, d_list, theta0, pol):
    """Interface/propagation form with field Fresnel coefficients and power factors (independent of `point`)."""
    k0 = 2 * np.pi / lam
    s = n_list[0] * np.sin(theta0)
    kz = []
    for n in n_list:
        k = np.sqrt(n * n - s * s + 0j)
        if k.imag < 0 or (k.imag == 0 and k.real < 0):
            k = -k
        kz.append(k)

    def interface(i, j):
        if pol == S_WAVE:
            r = (kz[i] - kz[j]) / (kz[i] + kz[j])
            t = 2 * kz[i] / (kz[i] + kz[j])
        else:
            ni2, nj2 = n_list[i] ** 2, n_list[j] ** 2
            r = (nj2 * kz[i] - ni2 * kz[j]) / (nj2 * kz[i] + ni2 * kz[j])
            t = 2 * n_list[i] * n_list[j] * kz[i] / (nj2 * kz[i] + ni2 * kz[j])
        return r, t

    N = len(n_list)
    M = np.eye(2, dtype=complex)
    for j in range(N - 1):
        r, t = interface(j, j + 1)
        M = M @ (np.array([[1, r], [r, 1]], dtype=complex) / t)
        if j + 1 < N - 1:
            ph = kz[j + 1] * k0 * d_list[j + 1]
            M = M @ np.array([[np.exp(-1j * ph), 0], [0, np.exp(1j * ph)]])
    t = 1 / M[0, 0]
    r = M[1, 0] / M[0, 0]
    R = abs(r) ** 2
    cf, c0 = kz[-1] / n_list[-1], kz[0] / n_list[0]
    if pol == S_WAVE:
        fac = (n_list[-1] * cf).real / (n_list[0] * c0).real
    else:
        fac = (n_list[-1] * np.conj(cf)).real / (n_list[0] * np.conj(c0)).real
    T = fac * abs(t) ** 2
    return T, R


def truth_point(lam, n_list, d_list, theta0, pol, dps=50):
    """50-digit (mpmath) characteristic-matrix evaluation of the same model; inputs taken as exact binary64."""
    import mpmath as mp

    with mp.workdps(dps):
        lam = mp.mpf(float(lam))
        k0 = 2 * mp.pi / lam
        ns = [mp.mpc(float(n.real), float(n.imag)) for n in n_list]
        ct0 = mp.cos(mp.mpf(float(theta0)))
        s2 = ns[0] ** 2 * (1 - ct0 ** 2)
        kz = []
        for n in ns:
            k = mp.sqrt(n * n - s2)
            if k.imag < 0 or (k.imag == 0 and k.real < 0):
                k = -k
            kz.append(k)
        is_s = pol == S_WAVE
        a = [k if is_s else n * n / k for n, k in zip(ns, kz)]
        v0 = mp.mpc(1) if is_s else kz[-1] / ns[-1]
        v1 = kz[-1] if is_s else ns[-1]
        for j in range(len(ns) - 2, 0, -1):
            ph = kz[j] * k0 * mp.mpf(float(d_list[j]))
            c, s = mp.cos(ph), mp.sin(ph)
            v0, v1 = c * v0 - 1j * s / a[j] * v1, -1j * a[j] * s * v0 + c * v1
        if is_s:
            m00 = (v0 + v1 / kz[0]) / 2
            m10 = (v0 - v1 / kz[0]) / 2
        else:
            c0 = kz[0] / ns[0]
            m00 = (v0 / c0 + v1 / ns[0]) / 2
            m10 = (v0 / c0 - v1 / ns[0]) / 2
        R = abs(m10 / m00) ** 2
        T = (kz[-1] / kz[0]).real * abs(1 / m00) ** 2
        return float(T), float(R)


# ------------------------------------------------------------------------------ per-layer field amplitudes
def field_amplitudes_point(lam, n_list, d_list, theta0, pol, snell=True):
    """(A_j, B_j) at the left boundary of every layer j for a unit incident amplitude (pst_field_amplitudes):
    D_j^-1 applied to the tangential field pair, D_j as tm.py:190-215.  snell=False is the reference's model (the
    incidence angle in every layer).  Layer 0 holds (1, r), the exit medium (t, 0).  Returns complex (L, 2)."""
    omega = 2 * np.pi * C0 / lam
    L = len(n_list)
    cs = layer_cosines(n_list, n_list[0], theta0) if snell else [np.cos(theta0) + 0j] * L
    Df = dynamical_matrix(n_list[-1], cs[-1], pol)
    v = Df @ np.array([1.0, 0.0], dtype=complex)  # unit transmitted amplitude
    raw = np.zeros((L, 2), dtype=complex)
    raw[L - 1] = (1.0, 0.0)
    for j in range(L - 2, 0, -1):
        kx = n_list[j] * omega / C0 * cs[j]
        D = dynamical_matrix(n_list[j], cs[j], pol)
        P = np.array([[np.exp(-1j * kx * d_list[j]), 0], [0, np.exp(1j * kx * d_list[j])]])
        v = np.linalg.multi_dot([D, P, np.linalg.inv(D)]) @ v
        raw[j] = np.linalg.inv(D) @ v
    raw[0] = np.linalg.inv(dynamical_matrix(n_list[0], cs[0], pol)) @ v
    return raw / raw[0, 0]
