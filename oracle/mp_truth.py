"""TEST INFRASTRUCTURE ONLY.  50-digit evaluation (mpmath) of the reference's model for adjudicating points where the
reference and the CUDA path disagree by more than the tight tolerance: whose error is it?

The inputs are taken as the exact binary64 numbers both implementations receive (wavelength, n, k, thickness and the
already rounded cos(theta)); the arithmetic is exact to ~50 digits, so the result is "the truth" of the model
   M = D_0^-1 (M_1 ... M_{N-2}) D_{N-1},  M_j = [[cos p, -i sin p / a], [-i a sin p, cos p]],  p = n w/c cos(theta) d
(reference transfer_matrix.py:190-277, 559-568, 482-486) for those inputs.
"""
from __future__ import annotations

import mpmath as mp

C0 = mp.mpf(299792458)


def truth_point(lam, n_list, d_list, cos_theta, pol, dps=50):
    """(T, R) of one spectral point.  n_list: complex per layer at this wavelength; returns mpf pair."""
    with mp.workdps(dps):
        lam = mp.mpf(float(lam))
        ct = mp.mpf(float(cos_theta))
        k0 = 2 * mp.pi / lam  # (n w / c) = n 2 pi / lambda
        is_s = pol == "s-wave"
        g = ct if is_s else 1 / ct
        nf = mp.mpc(float(n_list[-1].real), float(n_list[-1].imag))
        n0 = mp.mpc(float(n_list[0].real), float(n_list[0].imag))
        v0 = mp.mpc(1) if is_s else mp.mpc(ct)
        v1 = nf * ct if is_s else nf
        for j in range(len(n_list) - 2, 0, -1):
            n = mp.mpc(float(n_list[j].real), float(n_list[j].imag))
            ph = n * k0 * ct * mp.mpf(float(d_list[j]))
            c, s = mp.cos(ph), mp.sin(ph)
            a = n * g
            v0, v1 = c * v0 - 1j * s / a * v1, -1j * a * s * v0 + c * v1
        if is_s:
            m00 = (v0 + v1 / (n0 * ct)) / 2
            m10 = (v0 - v1 / (n0 * ct)) / 2
        else:
            m00 = (v0 / ct + v1 / n0) / 2
            m10 = (v0 / ct - v1 / n0) / 2
        R = abs(m10 / m00) ** 2
        T = (nf / n0).real * abs(1 / m00) ** 2
        return T, R


def truth_grid(wl, n_list, d_list, cos_thetas, pols=("s-wave", "p-wave")):
    """float arrays (n_pol, n_wl, n_theta) of the 50-digit result rounded to binary64."""
    import numpy as np

    T = np.empty((len(pols), len(wl), len(cos_thetas)))
    R = np.empty_like(T)
    for ip, pol in enumerate(pols):
        for il, lam in enumerate(wl):
            ns = [n[il] for n in n_list]
            for it, ct in enumerate(cos_thetas):
                t, r = truth_point(lam, ns, d_list, ct, pol)
                T[ip, il, it], R[ip, il, it] = float(t), float(r)
    return T, R
