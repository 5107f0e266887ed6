"""Low-contrast periodic stacks for the parity tests of the closed-form period power (typed kernel, pst_chain.cuh).

A quarter-wave stack of n = 1.46 / 1.45 has a stop band only 0.4 % wide: over the whole neighbourhood of the design
wavelength the period matrix is close to -I and the Bloch phase close to pi, which is where a power taken through the
trace alone (Chebyshev recurrence) loses O(k^2 eps) while the reference's layer-by-layer product loses O(k eps)
(ADVICE round 1).  The reference stays finite on these stacks at every depth used here.
"""
from __future__ import annotations

import numpy as np


def band_edge_wavelengths(n_hi, n_lo, n_wide, n_edge, lo=0.97, hi=1.03, lam0=1.0):
    """Wavelengths over [lo, hi] plus a dense set around the two edges of the first-order stop band."""
    half = (2.0 / np.pi) * np.arcsin(abs(n_hi - n_lo) / (n_hi + n_lo)) * lam0  # half width in wavelength, to first order
    edges = np.concatenate([lam0 - half * np.linspace(0.6, 1.6, n_edge), lam0 + half * np.linspace(0.6, 1.6, n_edge)])
    return np.unique(np.concatenate([np.linspace(lo, hi, n_wide), edges]))


def quarter_wave_stack(pairs, wl, n_hi=1.46, n_lo=1.45, n_in=1.0, n_out=1.5, spacer=None, mirrored=True):
    """(material columns [n_mat][n_wl] complex, layer_mat, layer_d).  spacer = (n complex, thickness) puts that layer
    between two mirrors of `pairs` pairs each (second mirror in mirrored order when `mirrored`)."""
    nw = len(wl)
    cols = [np.full(nw, n_in + 0j), np.full(nw, n_hi + 0j), np.full(nw, n_lo + 0j), np.full(nw, n_out + 0j)]
    d_hi, d_lo = 1.0 / (4 * n_hi), 1.0 / (4 * n_lo)
    lm = [0] + [1, 2] * pairs
    ld = [0.0] + [d_hi, d_lo] * pairs
    if spacer is not None:
        cols.append(np.full(nw, complex(spacer[0])))
        lm += [4] + ([2, 1] if mirrored else [1, 2]) * pairs
        ld += [float(spacer[1])] + ([d_lo, d_hi] if mirrored else [d_hi, d_lo]) * pairs
    lm.append(3)
    ld.append(0.0)
    return np.array(cols), lm, ld


def per_layer(cols, lm):
    """The same stack as one index array per layer (what the oracle and the host emulation take)."""
    return [cols[m] for m in lm]


# (pairs, wide samples, samples per band edge): sized so that the oracle (25 us per layer and wavelength, four passes
# with the sensitivity runs) stays within seconds per case
CASES = {50: (24, 6), 100: (24, 6), 500: (8, 4), 1500: (4, 2)}
THETAS = np.array([0.0, 0.4])
