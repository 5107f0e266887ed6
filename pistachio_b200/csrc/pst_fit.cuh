// Joint two-Lorentzian least-squares fit of one spectrum.
//
// The reference's "Polariton Peak Analysis" notebook fits every angle's spectrum with the sum of two lmfit
// LorentzianModels (cells 8-14: `LorentzianModel(prefix='l1_') + LorentzianModel(prefix='l2_')`, six parameters
// amplitude / center / sigma per line, least squares over the selected wavenumber window) and reads the lower and upper
// polariton off the two centers (cell 17).  lmfit is not part of the reference repository; its LorentzianModel is
//     f(x; A, mu, sigma) = (A / pi) sigma / ((x - mu)^2 + sigma^2)
// and its default minimiser is Levenberg-Marquardt (MINPACK through scipy).  This file is that fit for the device:
// Levenberg-Marquardt with Marquardt's diagonal scaling on the six parameters, seeded by the three-point estimates of
// pst_polariton_branches, one pass over the samples per evaluation.  Near the anticrossing the two lines overlap and a
// per-peak estimate is biased by the neighbour's tail; the joint fit is not -- that is where the Rabi splitting is read.
//
// Internal parameters: height h = A / (pi sigma), center mu, half width sigma per line (better conditioned than the
// area); results are reported in lmfit's (amplitude, center, sigma).  __host__ __device__: tests/host_emul runs the same
// source on the CPU against scipy.optimize.curve_fit.
#pragma once
#include "pst_math.cuh"

namespace pst {

constexpr int kFitParams = 6;  // h1, mu1, s1, h2, mu2, s2
constexpr int kFitTri = 21;    // packed lower triangle of the 6 x 6 normal matrix

struct FitProblem {
  const double* y;     // first sample of the spectrum (scan order handled by `stride` / `x` below)
  long long y_stride;  // element stride between consecutive samples of the window (negative when scanning in reverse)
  const double* x;     // axis value of the first sample of the window
  long long x_stride;
  int n;               // samples in the window
};

PST_HD double fit_ld(const double* p) {
#ifdef __CUDA_ARCH__
  return __ldg(p);
#else
  return *p;
#endif
}

// chi^2 at p; with JAC also the packed normal matrix N = J^T J and the gradient side g = J^T r (r = y - f)
template <bool JAC>
PST_HD double fit_eval(const FitProblem& pb, const double (&p)[kFitParams], double (&N)[kFitTri], double (&g)[kFitParams]) {
  if (JAC) {
#pragma unroll
    for (int k = 0; k < kFitTri; ++k) N[k] = 0.0;
#pragma unroll
    for (int k = 0; k < kFitParams; ++k) g[k] = 0.0;
  }
  double chi2 = 0.0;
  const double s1sq = p[2] * p[2], s2sq = p[5] * p[5];
  for (int i = 0; i < pb.n; ++i) {
    const double xv = fit_ld(pb.x + (long long)i * pb.x_stride), yv = fit_ld(pb.y + (long long)i * pb.y_stride);
    const double d1 = xv - p[1], d2 = xv - p[4];
    const double q1 = 1.0 / fma(d1, d1, s1sq), q2 = 1.0 / fma(d2, d2, s2sq);  // 1 / ((x - mu)^2 + sigma^2)
    const double l1 = s1sq * q1, l2 = s2sq * q2;                               // unit-height line shapes
    const double r = yv - fma(p[0], l1, p[3] * l2);
    chi2 = fma(r, r, chi2);
    if (JAC) {
      double J[kFitParams];
      J[0] = l1;
      J[1] = 2.0 * p[0] * l1 * d1 * q1;             // d f / d mu    = h sigma^2 2 (x - mu) / D^2
      J[2] = 2.0 * p[0] * p[2] * d1 * d1 * q1 * q1;  // d f / d sigma = 2 h sigma (x - mu)^2 / D^2
      J[3] = l2;
      J[4] = 2.0 * p[3] * l2 * d2 * q2;
      J[5] = 2.0 * p[3] * p[5] * d2 * d2 * q2 * q2;
      int t = 0;
#pragma unroll
      for (int a = 0; a < kFitParams; ++a) {
#pragma unroll
        for (int b = 0; b <= a; ++b) {
          N[t] = fma(J[a], J[b], N[t]);
          ++t;
        }
        g[a] = fma(J[a], r, g[a]);
      }
    }
  }
  return chi2;
}

// Solve (N + lambda diag(N)) delta = g by Cholesky on the packed triangle.  false: not positive definite.
PST_HD bool fit_solve(const double (&N)[kFitTri], const double (&g)[kFitParams], double lambda, double (&delta)[kFitParams]) {
  double L[kFitTri];
  int t = 0;
#pragma unroll
  for (int a = 0; a < kFitParams; ++a) {
#pragma unroll
    for (int b = 0; b <= a; ++b) {
      double s = N[t];
      if (a == b) s = fma(lambda, N[t], s);
#pragma unroll
      for (int k = 0; k < b; ++k) s = fma(-L[a * (a + 1) / 2 + k], L[b * (b + 1) / 2 + k], s);
      if (a == b) {
        if (!(s > 0.0)) return false;
        L[t] = sqrt(s);
      } else {
        L[t] = s / L[b * (b + 1) / 2 + b];
      }
      ++t;
    }
  }
  double z[kFitParams];
#pragma unroll
  for (int a = 0; a < kFitParams; ++a) {
    double s = g[a];
#pragma unroll
    for (int k = 0; k < a; ++k) s = fma(-L[a * (a + 1) / 2 + k], z[k], s);
    z[a] = s / L[a * (a + 1) / 2 + a];
  }
#pragma unroll
  for (int a = kFitParams - 1; a >= 0; --a) {
    double s = z[a];
#pragma unroll
    for (int k = a + 1; k < kFitParams; ++k) s = fma(-L[k * (k + 1) / 2 + a], delta[k], s);
    delta[a] = s / L[a * (a + 1) / 2 + a];
  }
  return true;
}

// seed / result layout (pst_polariton_branches' order): (center_lower, center_upper, hwhm_lower, hwhm_upper, X_lower, X_upper)
// with X = peak height in the seed and X = lmfit's amplitude (area, pi sigma height) in the result; out[6] = chi^2,
// out[7] = evaluations used (negative: stopped at max_iter without meeting the tolerance).  A seed width that is NaN or
// not positive (the three-point estimate gives up on flat tops) is replaced by sigma0; other NaN seeds give NaN results.
// Stopping rule (MINPACK's ftol / xtol, which lmfit sets to 1.5e-8): an accepted step that lowers chi^2 by no more than
// tol chi^2, or moves no parameter by more than tol of its size; tol is clamped to >= 1e-15.  On spectra that ARE two
// Lorentzians the misfit vanishes and convergence is quadratic, so a tight tol costs a few evaluations; on measured-like
// spectra (model misfit) Levenberg-Marquardt converges linearly along flat valleys and the reference's own result is only
// defined to lmfit's tolerance.  Steps that would carry a centre more than one window span outside the window, or a width
// beyond four spans, are refused like uphill steps (the unbounded fit can otherwise run away on a window holding one line).
PST_HD void fit_two_lorentzians(const FitProblem& pb, const double* seed, double sigma0, double tol, int max_iter, double* out) {
  const double nan = from_words(0x7ff80000, 0);
  double p[kFitParams] = {seed[4], seed[0], seed[2], seed[5], seed[1], seed[3]};
  if (!(p[2] > 0.0)) p[2] = sigma0;
  if (!(p[5] > 0.0)) p[5] = sigma0;
  bool ok = pb.n >= kFitParams;
#pragma unroll
  for (int k = 0; k < kFitParams; ++k) ok = ok && (p[k] == p[k]);
  ok = ok && p[2] > 0.0 && p[5] > 0.0;
  if (!ok) {
    for (int k = 0; k < 8; ++k) out[k] = nan;
    return;
  }
  tol = fmax(tol, 1.0e-15);
  const double xa = fit_ld(pb.x), xb = fit_ld(pb.x + (long long)(pb.n - 1) * pb.x_stride);
  const double span = fabs(xb - xa), c_lo = fmin(xa, xb) - span, c_hi = fmax(xa, xb) + span, s_hi = 4.0 * span;
  double N[kFitTri], g[kFitParams], delta[kFitParams], trial[kFitParams], dummyN[kFitTri], dummyg[kFitParams];
  double chi2 = fit_eval<true>(pb, p, N, g);
  double lambda = 1.0e-3;
  int evals = 1;
  bool converged = false;
  while (evals < max_iter) {
    if (!fit_solve(N, g, lambda, delta)) {
      lambda *= 10.0;
      if (lambda > 1.0e12) break;
      continue;
    }
#pragma unroll
    for (int k = 0; k < kFitParams; ++k) trial[k] = p[k] + delta[k];
    // the model is even in sigma (height parameterisation), so a width may cross zero freely -- as it may in the
    // unbounded lmfit model, where (amplitude, sigma) -> (-amplitude, -sigma) is a symmetry; |sigma| is reported
    if (!(fabs(trial[2]) <= s_hi && fabs(trial[5]) <= s_hi && trial[1] >= c_lo && trial[1] <= c_hi && trial[4] >= c_lo &&
          trial[4] <= c_hi)) {  // a line leaving the window (or a NaN): shorten the step
      lambda *= 10.0;
      if (lambda > 1.0e12) break;
      continue;
    }
    const double chi2_try = fit_eval<false>(pb, trial, dummyN, dummyg);
    ++evals;
    if (chi2_try <= chi2) {
      double step = 0.0, size = 0.0;
#pragma unroll
      for (int k = 0; k < kFitParams; ++k) {
        step = fmax(step, fabs(delta[k]) / fmax(fabs(trial[k]), 1e-300));
        size = fmax(size, fabs(trial[k]));
        p[k] = trial[k];
      }
      const double gain = chi2 - chi2_try;
      chi2 = fit_eval<true>(pb, p, N, g);
      ++evals;
      lambda = fmax(lambda * 0.1, 1.0e-12);
      // stop when neither the parameters nor the misfit still move by more than tol (relative)
      if (step <= tol || gain <= tol * chi2) {
        converged = true;
        break;
      }
    } else {
      lambda *= 10.0;
      if (lambda > 1.0e12) {  // no downhill step left at any damping: a minimum to working precision
        converged = true;
        break;
      }
    }
  }
  const int lo = p[1] <= p[4] ? 0 : 3, hi = 3 - lo;
  const double kPi = 3.141592653589793;
  out[0] = p[lo + 1];
  out[1] = p[hi + 1];
  out[2] = fabs(p[lo + 2]);
  out[3] = fabs(p[hi + 2]);
  out[4] = kPi * fabs(p[lo + 2]) * p[lo];  // lmfit amplitude = area
  out[5] = kPi * fabs(p[hi + 2]) * p[hi];
  out[6] = chi2;
  out[7] = converged ? (double)evals : -(double)evals;
}

}  // namespace pst
