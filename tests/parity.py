"""Parity rules between the CUDA path and the reference (via the oracle / golden fixtures).

north_star asks for <= 1e-12 relative agreement in fp64.  SURVEY.md section 7.2-1 shows that the reference itself is not
that accurate everywhere, so the rule is stated per point as

    |got - ref| <= 1e-12 * max(|ref|, 1e-3)                      ("tight": relative 1e-12, absolute 1e-15 below 1e-3)
                   + 4 * sens                                       (backward-error clause)
                   + ref_noise                                      (T only)

  sens      = how much the REFERENCE's own output moves when its inputs move by a few ulp: the oracle is re-run with
              every wavelength scaled by (1 +- 4 eps) and with every thickness scaled by an independent random
              (1 +- 4 eps).  Near a high-finesse resonance a 1-ulp change of any phase moves R/T by finesse * ulp, and
              two correct evaluations with different rounding cannot agree better than that.
  ref_noise = the rounding-error bound of the reference's numeric determinant in T = Re(det(M) |t|^2) (tm.py:486):
              8 eps (|M00 M11| + |M01 M10|) / |M00|^2.

The report counts how many points needed more than the tight term; tests bound that fraction so the extra clauses
cannot hide a wrong kernel.
"""
from __future__ import annotations

import numpy as np

from oracle import tmm_oracle as orc

EPS = np.finfo(np.float64).eps
POLS = (orc.S_WAVE, orc.P_WAVE)


def oracle_grid(wl, n_list, d_list, thetas, pols=POLS, want_m=False):
    stack = orc.Stack(wl, n_list, d_list)
    thetas = np.atleast_1d(thetas)
    T = np.empty((len(pols), len(wl), len(thetas)))
    R = np.empty_like(T)
    M = np.empty(T.shape + (2, 2), dtype=np.complex128) if want_m else None
    for ip, pol in enumerate(pols):
        for it, th in enumerate(thetas):
            if want_m:
                T[ip, :, it], R[ip, :, it], M[ip, :, it] = orc.spectrum(stack, float(th), pol, want_matrices=True)
            else:
                T[ip, :, it], R[ip, :, it] = orc.spectrum(stack, float(th), pol)
    return (T, R, M) if want_m else (T, R)


def sensitivity(wl, n_list, d_list, thetas, T0, R0, pols=POLS, seed=1234):
    """Per-point movement of the oracle's (T, R) under few-ulp input perturbations."""
    rng = np.random.default_rng(seed)
    sT = np.zeros_like(T0)
    sR = np.zeros_like(R0)
    wl = np.asarray(wl, dtype=np.float64)
    d = np.asarray(d_list, dtype=np.float64)
    for kind in ("wl+", "wl-", "d"):
        if kind == "wl+":
            T, R = oracle_grid(wl * (1 + 4 * EPS), n_list, d, thetas, pols)
        elif kind == "wl-":
            T, R = oracle_grid(wl * (1 - 4 * EPS), n_list, d, thetas, pols)
        else:
            T, R = oracle_grid(wl, n_list, d * (1 + 4 * EPS * rng.choice([-1.0, 1.0], size=d.shape)), thetas, pols)
        with np.errstate(invalid="ignore"):
            sT = np.fmax(sT, np.abs(T - T0))
            sR = np.fmax(sR, np.abs(R - R0))
    return sT, sR


def det_noise(M):
    """Rounding-error bound of the reference's T from its numeric determinant; M: (..., 2, 2)."""
    m00, m01, m10, m11 = M[..., 0, 0], M[..., 0, 1], M[..., 1, 0], M[..., 1, 1]
    with np.errstate(all="ignore"):
        return 8 * EPS * (np.abs(m00 * m11) + np.abs(m01 * m10)) / np.abs(m00) ** 2


class Report:
    def __init__(self, name):
        self.name = name
        self.n = 0
        self.n_loose = 0
        self.n_fail = 0
        self.max_tight_ratio = 0.0  # max |d| / tight tolerance over points that pass on the tight term
        self.worst = None

    def add(self, what, got, ref, extra):
        got, ref, extra = np.asarray(got), np.asarray(ref), np.asarray(extra)
        finite = np.isfinite(ref)
        d = np.abs(got - ref)
        tight = 1e-12 * np.maximum(np.abs(ref), 1e-3)
        ok_tight = d <= tight
        ok = d <= tight + extra
        ok_tight &= finite
        bad = finite & ~ok
        self.n += int(finite.sum())
        self.n_loose += int((finite & ok & ~ok_tight).sum())
        self.n_fail += int(bad.sum())
        if ok_tight.any():
            self.max_tight_ratio = max(self.max_tight_ratio, float((d[ok_tight] / tight[ok_tight]).max()))
        if bad.any():
            i = np.unravel_index(np.argmax(np.where(bad, d / (tight + extra), 0)), d.shape)
            self.worst = (what, i, float(got[i]), float(ref[i]), float(d[i]), float((tight + extra)[i]))
        # results must be finite wherever the reference is
        assert np.isfinite(got[finite]).all(), f"{self.name}/{what}: non-finite CUDA result where the reference is finite"

    @property
    def loose_fraction(self):
        return self.n_loose / max(self.n, 1)

    def summary(self):
        return (f"{self.name}: {self.n} values, {self.n_fail} fail, {self.n_loose} needed the conditioning clause "
                f"({100 * self.loose_fraction:.3f}%), worst tight ratio {self.max_tight_ratio:.3f}")

    def check(self, max_loose_fraction=0.02):
        print(self.summary())
        assert self.n_fail == 0, f"{self.summary()} worst={self.worst}"
        assert self.loose_fraction <= max_loose_fraction, self.summary()


def compare_grid(name, got_T, got_R, wl, n_list, d_list, thetas, ref_T=None, ref_R=None, pols=POLS, got_A=None):
    """Full parity check of (n_pol, n_wl, n_theta) arrays against the oracle (or supplied reference values)."""
    T0, R0, M0 = oracle_grid(wl, n_list, d_list, thetas, pols, want_m=True)
    if ref_T is None:
        ref_T, ref_R = T0, R0
    sT, sR = sensitivity(wl, n_list, d_list, thetas, T0, R0, pols)
    rep = Report(name)
    rep.add("R", got_R, ref_R, 4 * sR)
    rep.add("T", got_T, ref_T, 4 * sT + det_noise(M0))
    if got_A is not None:
        assert np.array_equal(got_A, (1.0 - got_R) - got_T), f"{name}: A is not (1 - R) - T"
    return rep
