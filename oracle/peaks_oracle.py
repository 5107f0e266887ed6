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
