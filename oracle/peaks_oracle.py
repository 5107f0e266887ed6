"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.

Restatement of the peak search the reference's analysis notebook applies to spectra
(notebooks/Cavity Length Determiner.ipynb, cell 3: `signal.find_peaks(intensity, height=1.0, distance=500)`, then the
mean peak spacing -> cavity length L = 1e4 / (2 n_eff <d nu>), cells 5-6).  The arithmetic lives in scipy
(scipy.signal._peak_finding / _peak_finding_utils, version unpinned by the reference; 1.18.1 here):

  local maxima   strict rise on the left, plateaus reported at their midpoint (left+right)//2, strict fall on the right;
                 the first and last sample can never be peaks                         (_local_maxima_1d)
  height         keep peaks with x[peak] >= height                                    (_select_by_property)
  distance       visit peaks from the highest to the lowest; a visited, still-kept peak removes every neighbour closer
                 than `distance` samples                                              (_select_by_peak_distance)

tests/test_peaks.py pins this restatement to scipy.signal.find_peaks itself.
"""
from __future__ import annotations

import numpy as np


def local_maxima(x):
    x = np.asarray(x, dtype=np.float64)
    peaks = []
    i, i_max = 1, len(x) - 1
    while i < i_max:
        if x[i - 1] < x[i]:
            ahead = i + 1
            while ahead < i_max and x[ahead] == x[i]:
                ahead += 1
            if x[ahead] < x[i]:
                peaks.append((i + ahead - 1) // 2)
                i = ahead
        i += 1
    return np.array(peaks, dtype=np.intp)


def select_by_distance(peaks, heights, distance):
    n = len(peaks)
    distance = int(np.ceil(distance))
    keep = np.ones(n, dtype=bool)
    order = np.argsort(heights)
    for i in range(n - 1, -1, -1):
        j = order[i]
        if not keep[j]:
            continue
        k = j - 1
        while k >= 0 and peaks[j] - peaks[k] < distance:
            keep[k] = False
            k -= 1
        k = j + 1
        while k < n and peaks[k] - peaks[j] < distance:
            keep[k] = False
            k += 1
    return keep


def find_peaks(x, height=None, distance=None):
    x = np.asarray(x, dtype=np.float64)
    peaks = local_maxima(x)
    if height is not None:
        peaks = peaks[x[peaks] >= height]
    if distance is not None:
        peaks = peaks[select_by_distance(peaks, x[peaks], distance)]
    return peaks


def mean_spacing(axis, peaks):
    """pandas `Series(axis[peaks]).diff().mean()` of the notebook (cell 5): telescopes to (last - first)/(k - 1)."""
    if len(peaks) < 2:
        return float("nan")
    vals = np.asarray(axis, dtype=np.float64)[peaks]
    return float(np.mean(np.diff(vals)))


def lorentzian_three_point(x, y, i):
    """Closed-form Lorentzian through samples i-1, i, i+1: for y = A / (1 + ((x - x0)/g)^2) the reciprocal is a
    parabola, so divided differences of 1/y give (x0, g, A) exactly.  Restatement of pst_polariton_branches' refinement
    (there is no reference oracle: the notebook's lmfit two-Lorentzian fit is not part of the reference repository)."""
    x0, x1, x2 = (float(x[i - 1]), float(x[i]), float(x[i + 1]))
    z0, z1, z2 = (1.0 / float(y[i - 1]), 1.0 / float(y[i]), 1.0 / float(y[i + 1]))
    s01, s12 = (z1 - z0) / (x1 - x0), (z2 - z1) / (x2 - x1)
    a = (s12 - s01) / (x2 - x0)
    b1 = s01 + a * (x1 - x0)
    if not (a > 0.0 and np.isfinite(a) and np.isfinite(b1)):
        return x1, float("nan"), float(y[i])
    zmin = z1 - b1 * b1 / (4.0 * a)
    if zmin <= 0.0:
        return x1 - b1 / (2.0 * a), float("nan"), float("nan")
    return x1 - b1 / (2.0 * a), float(np.sqrt(zmin / a)), 1.0 / zmin


def polariton_branches(x, y, height=None, distance=None):
    """(x0_lower, x0_upper, g_lower, g_upper, A_lower, A_upper) of the two highest peaks of one spectrum."""
    peaks = find_peaks(y, height=height, distance=distance)
    if len(peaks) < 2:
        return (float("nan"),) * 6
    h = np.asarray(y)[peaks]
    j1 = int(np.argmax(h))
    rest = [k for k in range(len(peaks)) if k != j1]
    j2 = rest[int(np.argmax(h[rest]))]
    a, b = lorentzian_three_point(x, y, peaks[j1]), lorentzian_three_point(x, y, peaks[j2])
    lo, hi = (a, b) if a[0] <= b[0] else (b, a)
    return lo[0], hi[0], lo[1], hi[1], lo[2], hi[2]


# --------------------------------------------------------------------------- joint two-Lorentzian fit
def lorentzian(x, amplitude, center, sigma):
    """lmfit's LorentzianModel line shape (lmfit.lineshapes.lorentzian; lmfit is what the reference's notebook imports,
    "Polariton Peak Analysis" cell 2, and is not part of the reference repository)."""
    return (amplitude / np.pi) * sigma / ((x - center) ** 2 + sigma ** 2)


def two_lorentzians(x, a1, c1, s1, a2, c2, s2):
    return lorentzian(x, a1, c1, s1) + lorentzian(x, a2, c2, s2)


def fit_two_lorentzians(x, y, seed, sigma0=None, tol=1e-15):
    """The notebook's fit (cells 8-14): `LorentzianModel(prefix='l1_') + LorentzianModel(prefix='l2_')` least-squares
    fitted to one spectrum -- here with scipy.optimize.curve_fit (MINPACK's Levenberg-Marquardt, the minimiser lmfit
    calls by default), tolerances at machine level by default so that the minimum itself is the oracle (tol = 1.5e-8 is
    lmfit's default ftol = xtol).  `seed` = (center_lower,
    center_upper, hwhm_lower, hwhm_upper, height_lower, height_upper), the row pst_polariton_branches produces.
    Returns (center_lower, center_upper, sigma_lower, sigma_upper, amplitude_lower, amplitude_upper, chi2)."""
    from scipy import optimize

    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    c1, c2, g1, g2, h1, h2 = (float(v) for v in seed)
    if sigma0 is not None:
        g1 = g1 if g1 > 0 else sigma0
        g2 = g2 if g2 > 0 else sigma0
    p0 = [np.pi * g1 * h1, c1, g1, np.pi * g2 * h2, c2, g2]
    p, _ = optimize.curve_fit(two_lorentzians, x, y, p0=p0, method="lm", ftol=tol, xtol=tol, gtol=min(tol, 1e-15), maxfev=20000)
    p = [float(v) for v in p]
    p[2], p[5] = abs(p[2]), abs(p[5])  # the model is even in sigma up to the sign of the amplitude
    lo, hi = (p[:3], p[3:]) if p[1] <= p[4] else (p[3:], p[:3])
    chi2 = float(np.sum((y - two_lorentzians(x, *p)) ** 2))
    return lo[1], hi[1], lo[2], hi[2], abs(lo[0]), abs(hi[0]), chi2
