// Per-point arithmetic of the TMM chain: layer-matrix build, layer step, exact rescaling, ordered segment combine,
// epilogue.  Everything here is __host__ __device__ so that tests/host_emul can execute the same source on the CPU
// box; the shipped library only ever calls it from kernels (pst_kernels.cuh).
//
// Model (reference tm.py:190-277, 559-568, 482-486), phi = kx d, a = n g with
// g = cos(theta) for s-waves and 1/cos(theta) for p-waves (same theta in every layer: the reference has no Snell
// refraction):
//   M_j = D_j P_j D_j^-1 = [[cos phi, -i sin phi / a], [-i a sin phi, cos phi]]
//   M   = D_0^-1 (M_1 ... M_{N-2}) D_{N-1},  t = 1/M00, r = M10/M00, R = |r|^2, T = Re(det M |t|^2)
// Only the requested columns of M are propagated, from the exit side: v <- M_j v.
#pragma once
#include "pst_math.cuh"

namespace pst {

// Exponent pulled out by a rescale: [-1022, 1000] keeps 2^-e a normal number (a zero or denormal state is scaled by
// 2^1022, harmlessly; inf/nan stay what they are)
PST_HD int clamp_exp(int e) { return e > 1000 ? 1000 : (e < -1022 ? -1022 : e); }

// State per polarisation and column: v = (v0.re, v0.im, v1.re, v1.im); true value = v * 2^kexp.
template <int NPOL, int NC>
struct ChainState {
  double v[NPOL][NC][4];
  int kexp[NPOL];
};

// Optical constants of one (material, wavelength), in shared-memory order (48 bytes, three 16-byte loads):
//   q = (n w)(1/c) in the reference's rounding order, n, 1/n.  The lossless path reads only the first 24 bytes + ni.
struct OptRow {
  double qr, nr, inr, ni, qi, ini;
};

template <int NPOL>
PST_HD bool pol_is_s(int p, int pol_first) {
  return (NPOL == 2) ? (p == 0) : (pol_first == 0);
}

// ---- layer matrices ------------------------------------------------------------------------------------------
// Lossless layer (k = 0 at every wavelength of its material): phi and n real,
//   M = [[c, -i be], [-i al, c]],  al = a sin(phi), be = sin(phi)/a, one (al, be) pair per polarisation.
template <int NPOL>
struct RealLayer {
  double c;
  double al[NPOL], be[NPOL];
  double sn, si;  // n sin(phi), sin(phi)/n: the polarisation-independent parts of al = g sn, be = sin(phi)/(n g) = si/g
};
// General layer: M = 2^sc [[cphi, m01], [m10, cphi]] with complex entries.
template <int NPOL>
struct CplxLayer {
  cplx cphi;
  cplx m10[NPOL], m01[NPOL];
  int sc;
};

// Layer-matrix construction is split in two so that callers can batch the transcendental part of several layers:
//   phase  x = (q.re cos) d, y = (q.im cos) d in the reference's rounding order (tm.py:270, 234)
//   finish entries of M from sin x, cos x (and cosh y, sinh y)
PST_HD double layer_phase(double q, double ct, double d) { return mul_rn(mul_rn(q, ct), d); }

template <int NPOL>
PST_HD RealLayer<NPOL> finish_real_layer(double sx, double cx, double nr, double inr, double ct, double ict,
                                         int pol_first) {
  const double sn = sx * nr, si = sx * inr;
  RealLayer<NPOL> m;
  m.c = cx;
  m.sn = sn;
  m.si = si;
#pragma unroll
  for (int p = 0; p < NPOL; ++p) {
    const bool is_s = pol_is_s<NPOL>(p, pol_first);
    m.al[p] = (is_s ? ct : ict) * sn;
    m.be[p] = (is_s ? ict : ct) * si;
  }
  return m;
}

// the same with the polarisation factors resolved by the caller (g = cos for s, 1/cos for p; gi = 1/g): loops over many
// layers hoist the selection
template <int NPOL>
PST_HD RealLayer<NPOL> finish_real_layer_g(double sx, double cx, double nr, double inr, const double (&g)[NPOL],
                                           const double (&gi)[NPOL]) {
  const double sn = sx * nr, si = sx * inr;
  RealLayer<NPOL> m;
  m.c = cx;
  m.sn = sn;
  m.si = si;
#pragma unroll
  for (int p = 0; p < NPOL; ++p) {
    m.al[p] = g[p] * sn;
    m.be[p] = gi[p] * si;
  }
  return m;
}

template <int NPOL>
PST_HD CplxLayer<NPOL> finish_cplx_layer(double sx, double cx, double y, const OptRow& o, double ct, double ict,
                                         int pol_first) {
  double ch, sh;
  CplxLayer<NPOL> m;
  m.sc = cosh_sinh_scaled(y, ch, sh);
  // cos(x+iy) = cos x cosh y - i sin x sinh y ; sin(x+iy) = sin x cosh y + i cos x sinh y   (both times 2^sc)
  m.cphi = {cx * ch, -(sx * sh)};
  const cplx sphi = {sx * ch, cx * sh};
  const cplx sn = cmul(sphi, {o.nr, o.ni});
  const cplx si = cmul(sphi, {o.inr, o.ini});
#pragma unroll
  for (int p = 0; p < NPOL; ++p) {
    const bool is_s = pol_is_s<NPOL>(p, pol_first);
    const double g = is_s ? ct : ict, gi = is_s ? ict : ct;
    m.m10[p] = {g * sn.im, -(g * sn.re)};    // -i a sin(phi)
    m.m01[p] = {gi * si.im, -(gi * si.re)};  // -i sin(phi) / a
  }
  return m;
}

// IN_RANGE: the caller has bounded the phase inside sincos' fast range (|x| < 2^30): no range test and no call to the
// out-of-line large-argument path (a possible call inside a layer loop costs a convergence barrier per layer)
template <int NPOL, bool IN_RANGE = false>
PST_HD RealLayer<NPOL> build_real_layer(double qr, double nr, double inr, double d, double ct, double ict,
                                        int pol_first) {
  double sx, cx;
  if constexpr (IN_RANGE) sincos_core(layer_phase(qr, ct, d), sx, cx);
  else sincos_fast(layer_phase(qr, ct, d), sx, cx);
  return finish_real_layer<NPOL>(sx, cx, nr, inr, ct, ict, pol_first);
}

template <int NPOL, bool IN_RANGE = false>
PST_HD CplxLayer<NPOL> build_cplx_layer(const OptRow& o, double d, double ct, double ict, int pol_first) {
  double sx, cx;
  if constexpr (IN_RANGE) sincos_core(layer_phase(o.qr, ct, d), sx, cx);
  else sincos_fast(layer_phase(o.qr, ct, d), sx, cx);
  return finish_cplx_layer<NPOL>(sx, cx, layer_phase(o.qi, ct, d), o, ct, ict, pol_first);
}

// v <- M v for every polarisation/column held by the thread
template <int NPOL, int NC>
PST_HD void apply_real_layer(ChainState<NPOL, NC>& st, const RealLayer<NPOL>& m) {
#pragma unroll
  for (int p = 0; p < NPOL; ++p) {
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      double* v = st.v[p][c];
      const double v0r = v[0], v0i = v[1], v1r = v[2], v1i = v[3];
      v[0] = fma(m.c, v0r, m.be[p] * v1i);
      v[1] = fma(m.c, v0i, -(m.be[p] * v1r));
      v[2] = fma(m.c, v1r, m.al[p] * v0i);
      v[3] = fma(m.c, v1i, -(m.al[p] * v0r));
    }
  }
}

template <int NPOL, int NC>
PST_HD void apply_cplx_layer(ChainState<NPOL, NC>& st, const CplxLayer<NPOL>& m) {
#pragma unroll
  for (int p = 0; p < NPOL; ++p) {
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      double* v = st.v[p][c];
      const cplx v0 = {v[0], v[1]}, v1 = {v[2], v[3]};
      v[0] = fma(m.cphi.re, v0.re, fma(-m.cphi.im, v0.im, fma(m.m01[p].re, v1.re, -(m.m01[p].im * v1.im))));
      v[1] = fma(m.cphi.re, v0.im, fma(m.cphi.im, v0.re, fma(m.m01[p].re, v1.im, m.m01[p].im * v1.re)));
      v[2] = fma(m.m10[p].re, v0.re, fma(-m.m10[p].im, v0.im, fma(m.cphi.re, v1.re, -(m.cphi.im * v1.im))));
      v[3] = fma(m.m10[p].re, v0.im, fma(m.m10[p].im, v0.re, fma(m.cphi.re, v1.im, m.cphi.im * v1.re)));
    }
    st.kexp[p] += m.sc;
  }
}

// A cached layer matrix in either form (the typed paths).  Real form: a[0] = c, a[1 + 2p] = al[p], a[2 + 2p] = be[p],
// a[5] = sn, a[6] = si (polarisation-independent parts, used by the period power).
// Complex form: a[0..1] = cphi, a[2 + 4p .. 5 + 4p] = m10[p], m01[p]; sc = power-of-two exponent.
template <int NPOL>
struct TypeRegs {
  double a[10];
  int sc;
  PST_HD void set_real(const RealLayer<NPOL>& m) {
    a[0] = m.c;
#pragma unroll
    for (int p = 0; p < NPOL; ++p) { a[1 + 2 * p] = m.al[p]; a[2 + 2 * p] = m.be[p]; }
    a[5] = m.sn;
    a[6] = m.si;
    sc = 0;
  }
  PST_HD void set_cplx(const CplxLayer<NPOL>& m) {
    a[0] = m.cphi.re; a[1] = m.cphi.im;
#pragma unroll
    for (int p = 0; p < NPOL; ++p) {
      a[2 + 4 * p] = m.m10[p].re; a[3 + 4 * p] = m.m10[p].im;
      a[4 + 4 * p] = m.m01[p].re; a[5 + 4 * p] = m.m01[p].im;
    }
    sc = m.sc;
  }
};

template <int NPOL, int NC>
PST_HD void apply_type(ChainState<NPOL, NC>& st, const TypeRegs<NPOL>& t, bool lossy) {
  if (!lossy) {
    RealLayer<NPOL> m;
    m.c = t.a[0];
#pragma unroll
    for (int p = 0; p < NPOL; ++p) { m.al[p] = t.a[1 + 2 * p]; m.be[p] = t.a[2 + 2 * p]; }
    apply_real_layer<NPOL, NC>(st, m);
  } else {
    CplxLayer<NPOL> m;
    m.cphi = {t.a[0], t.a[1]};
#pragma unroll
    for (int p = 0; p < NPOL; ++p) {
      m.m10[p] = {t.a[2 + 4 * p], t.a[3 + 4 * p]};
      m.m01[p] = {t.a[4 + 4 * p], t.a[5 + 4 * p]};
    }
    m.sc = t.sc;
    apply_cplx_layer<NPOL, NC>(st, m);
  }
}

// k repetitions of "apply A, then B" with both matrices in registers (the typed path's inner loop)
template <int NPOL, int NC, bool LA, bool LB>
PST_HD void pair_loop(ChainState<NPOL, NC>& st, const TypeRegs<NPOL>& A, const TypeRegs<NPOL>& B, int k) {
  for (int i = 0; i < k; ++i) {
    apply_type<NPOL, NC>(st, A, LA);
    apply_type<NPOL, NC>(st, B, LB);
  }
}

// k repetitions of a LOSSLESS pair in closed form.  One period is U = B A (A is applied first) with
//   U = [[u00, -i u01], [-i u10, u11]],  u00 = cB cA - beB alA,  u11 = cB cA - alB beA,
//                                         u01 = cB beA + beB cA,  u10 = alB cA + cB alA      (all real, det U = 1).
// With al = g sn and be = si / g (g = cos(theta) for s-waves, 1/cos(theta) for p-waves; sn = n sin(phi), si = sin(phi)/n)
// the diagonal of U does not depend on the polarisation and the off-diagonal entries only through one factor:
//   u00 = cB cA - siB snA,  u11 = cB cA - snB siA,  u01 = (cB siA + siB cA) / g,  u10 = g (snB cA + cB snA),
// so the period is formed ONCE per point from the bare quantities (u01, u10 below are the bracketed sums) and the
// polarisations part ways only when the finished power is applied to their state vectors -- the same sharing as for the
// phase and the sincos of a layer.
// Split off the trace:  U = x I + W,  x = (u00 + u11)/2,  W = [[w, -i u01/g], [-i g u10, -w]],  w = (u00 - u11)/2, and
//   W^2 = -s2 I,  s2 = u01 u10 - w^2  (= sin^2 of the Bloch phase; negative inside a stop band),
// so every power of U lives in the two-dimensional commutative ring {t I + s W}:  U^k = t_k I + s_k W with
//   (t, s)(t', s') = (t t' - s2 s s', t s' + s t').
// The power is taken by binary exponentiation in that ring (left to right over the bits of k: one squaring
// (t^2 - s2 s^2, 2 t s) per bit, one multiplication by U = (x, 1) per set bit): ceil(log2 k) steps instead of k
// layer-pair products, with the forward error of a product of k matrices, O(k eps).
//
// Why this form and not U^k = S_{k-1} U - S_{k-2} I with the Chebyshev recurrence S_j = 2x S_{j-1} - S_{j-2} (Yeh,
// Optical Waves in Layered Media; the form round 1 shipped): near a band edge and for low-contrast stacks U is close
// to +-I, the S_j grow like 1/sin(theta) and U^k comes out of a cancellation between S_{k-1} U and S_{k-2} I; the
// recurrence sees the Bloch phase only through 1 - |x| (absolute error eps on a quantity of size sin^2/2) and loses
// O(k^2 eps) -- measured 4e-12 on R at 500 pairs of n = 1.46/1.45, where the reference's layer-by-layer product is
// good to 3e-14.  Here the phase enters through s2, which is computed from the small quantities themselves (u01 u10
// and w^2 carry relative, not absolute, rounding errors), x only contributes sin(theta) eps, and t I + s W has no
// cancellation.
//
// The ring element depends on the period only through (x, s2), which are the same for U = B A and its mirror image
// A B (w changes sign, u01 and u10 stay); `mirror_image` makes the sums round in the operand order of the period's
// first occurrence, so that a period and its mirror image agree bit for bit: the second Bragg stack of a symmetric
// cavity is the first one's matrix with p00 and p11 exchanged, and the powering runs once.
struct Period {
  double x, s2, w, u01, u10;
};
// the polarisation-independent part of a lossless layer matrix: cos(phi), n sin(phi), sin(phi)/n
struct BareLayer {
  double c, sn, si;
};
PST_HD BareLayer finish_bare_layer(double sx, double cx, double nr, double inr) { return {cx, sx * nr, sx * inr}; }

PST_HD void period_form(const BareLayer& A, const BareLayer& B, bool mirror_image, Period& U) {
  const double cA = A.c, cB = B.c;
  const double snA = A.sn, siA = A.si, snB = B.sn, siB = B.si;
  const double cc = cB * cA;
  double sum, dif;  // siB snA + snB siA = 2 cc - (u00 + u11);  snB siA - siB snA = u00 - u11
  if (!mirror_image) {
    const double m = siB * snA;
    sum = fma(snB, siA, m);
    dif = fma(snB, siA, -m);
    U.u01 = fma(cB, siA, siB * cA);
    U.u10 = fma(snB, cA, cB * snA);
  } else {  // the same numbers as the call with A and B exchanged, w negated
    const double m = siA * snB;
    sum = fma(snA, siB, m);
    dif = -fma(snA, siB, -m);
    U.u01 = fma(cA, siB, siA * cB);
    U.u10 = fma(snA, cB, cA * snB);
  }
  U.x = fma(-0.5, sum, cc);
  U.w = 0.5 * dif;
  U.s2 = fma(-U.w, U.w, U.u01 * U.u10);
}
template <int NPOL>
PST_HD void period_form(const TypeRegs<NPOL>& A, const TypeRegs<NPOL>& B, bool mirror_image, Period& U) {
  period_form(BareLayer{A.a[0], A.a[5], A.a[6]}, BareLayer{B.a[0], B.a[5], B.a[6]}, mirror_image, U);
}

// (t_k, s_k) 2^ks with U^k = t_k I + s_k W.
struct PairPower {
  double t, s;
  int ks;
};

// One bit of the exponent: square, multiply by U when the bit is set (block-uniform), and -- RESCALE -- pull an exact
// power of two out (needed when |eigenvalue|^k can leave the fp64 range: deep stop bands).
template <bool RESCALE>
PST_HD void pair_power_step(const Period& U, bool mul, PairPower& S) {
  const double t = S.t, s = S.s;
  double tn = fma(t, t, -((U.s2 * s) * s));  // (t, s)^2
  double sn = (t + t) * s;
  if (__builtin_expect(mul, 0)) {  // times U = (x, 1).  `mul` is block-uniform; marked unlikely so that it compiles to
    const double tm = fma(U.x, tn, -(U.s2 * sn));  // predicated instructions (issue slots only when the bit is clear)
    sn = fma(U.x, sn, tn);                         // instead of three unconditional FP64 instructions and four selects
    tn = tm;
  }
  if (RESCALE) {
    unsigned m = (unsigned)hi_word(tn) & 0x7fffffffu;
    const unsigned h = (unsigned)hi_word(sn) & 0x7fffffffu;
    m = h > m ? h : m;
    int e = (int)(m >> 20) - 1023;
    e = clamp_exp(e);
    const double f = from_words((1023 - e) << 20, 0);
    tn *= f;
    sn *= f;
    S.ks = 2 * S.ks + e;
  }
  S.t = tn;
  S.s = sn;
}

// Left-to-right binary exponentiation.  k is block-uniform (it comes from the program): the bits below the leading
// one are dispatched through one switch into straight-line code (k < 8192), so nothing loop-carried moves between
// registers; deeper stacks take the loop.
PST_HD int log2_floor(int k) {
  int hb = 0;
  while ((k >> (hb + 1)) != 0) ++hb;
  return hb;
}
// hb = log2_floor(k): callers that know it (the host resolves it for the cavity kernel) pass it in
template <bool RESCALE>
PST_HD void pair_power_coeffs_t(const Period& U, int k, int hb, PairPower& S) {
  S.t = U.x;  // U^1
  S.s = 1.0;
  S.ks = 0;
  if (hb > 12) {
    for (int bit = hb - 1; bit >= 12; --bit) pair_power_step<RESCALE>(U, (k >> bit) & 1, S);
    hb = 12;
  }
  switch (hb) {
    case 12: pair_power_step<RESCALE>(U, (k >> 11) & 1, S);  // fall through
    case 11: pair_power_step<RESCALE>(U, (k >> 10) & 1, S);  // fall through
    case 10: pair_power_step<RESCALE>(U, (k >> 9) & 1, S);   // fall through
    case 9: pair_power_step<RESCALE>(U, (k >> 8) & 1, S);    // fall through
    case 8: pair_power_step<RESCALE>(U, (k >> 7) & 1, S);    // fall through
    case 7: pair_power_step<RESCALE>(U, (k >> 6) & 1, S);    // fall through
    case 6: pair_power_step<RESCALE>(U, (k >> 5) & 1, S);    // fall through
    case 5: pair_power_step<RESCALE>(U, (k >> 4) & 1, S);    // fall through
    case 4: pair_power_step<RESCALE>(U, (k >> 3) & 1, S);    // fall through
    case 3: pair_power_step<RESCALE>(U, (k >> 2) & 1, S);    // fall through
    case 2: pair_power_step<RESCALE>(U, (k >> 1) & 1, S);    // fall through
    case 1: pair_power_step<RESCALE>(U, k & 1, S);           // fall through
    default: break;
  }
}

// Pair counts up to kPowerFixedMax: the bits of k resolved at COMPILE time.  k is block-uniform, so one jump selects a
// straight-line chain of exactly the squarings and multiplications k needs (k = 20: 4 squarings + 1 multiplication =
// 23 FP64 instructions, no bit tests, no predicated multiplications, no register moves between steps).  The operations
// and their order are those of pair_power_coeffs_t<false>, so the two agree bit for bit (the interpreter keeps the
// generic form; tests/test_gpu_parity.py::test_cavity_shape_fast_path_is_bit_identical compares them).
constexpr int kPowerFixedMax = 40;
template <int K>
PST_HD void pair_power_fixed(const Period& U, PairPower& S) {
  if constexpr (K <= 1) {
    S.t = U.x;
    S.s = 1.0;
    S.ks = 0;
  } else {
    pair_power_fixed<K / 2>(U, S);
    pair_power_step<false>(U, (K & 1) != 0, S);
  }
}
// false: k is outside the compiled range (the caller takes the generic form)
PST_HD bool pair_power_small(const Period& U, int k, PairPower& S) {
#define PST_PPF(K) case K: pair_power_fixed<K>(U, S); return true;
  switch (k) {
    PST_PPF(1) PST_PPF(2) PST_PPF(3) PST_PPF(4) PST_PPF(5) PST_PPF(6) PST_PPF(7) PST_PPF(8) PST_PPF(9) PST_PPF(10)
    PST_PPF(11) PST_PPF(12) PST_PPF(13) PST_PPF(14) PST_PPF(15) PST_PPF(16) PST_PPF(17) PST_PPF(18) PST_PPF(19) PST_PPF(20)
    PST_PPF(21) PST_PPF(22) PST_PPF(23) PST_PPF(24) PST_PPF(25) PST_PPF(26) PST_PPF(27) PST_PPF(28) PST_PPF(29) PST_PPF(30)
    PST_PPF(31) PST_PPF(32) PST_PPF(33) PST_PPF(34) PST_PPF(35) PST_PPF(36) PST_PPF(37) PST_PPF(38) PST_PPF(39) PST_PPF(40)
    default: return false;
  }
#undef PST_PPF
}

// Whether the powering must rescale: |t_k|, |s_k| <= k G^(2k) with G <= 2^lg the per-layer growth bound.  Block-uniform
// by construction (lg is grid-wide, k comes from the program).  Without rescaling the entries of U^k stay below
// 2^kPowerNoRescaleLog2.
constexpr int kPowerNoRescaleLog2 = 440;
PST_HD bool pair_power_needs_rescale(int k, int lg) { return 2 * lg * k + 40 > kPowerNoRescaleLog2; }

PST_HD void pair_power_coeffs(const Period& U, int k, int lg, PairPower& S) {
  if (pair_power_needs_rescale(k, lg)) pair_power_coeffs_t<true>(U, k, log2_floor(k), S);
  else if (!pair_power_small(U, k, S)) pair_power_coeffs_t<false>(U, k, log2_floor(k), S);
}

// U^k = t I + s W as a matrix [[p00, -i p01], [-i p10, p11]] (real entries) times 2^ks: the diagonal is shared by the
// polarisations, the off-diagonal entries take the factors 1/g and g
template <int NPOL>
struct PairMatrix {
  double p00, p11;
  double p01[NPOL], p10[NPOL];
  int ks;
};

template <int NPOL>
PST_HD void pair_power_matrix(const Period& U, const PairPower& S, double ct, double ict, int pol_first, PairMatrix<NPOL>& P) {
  P.p00 = fma(S.s, U.w, S.t);
  P.p11 = fma(-S.s, U.w, S.t);
  const double b01 = S.s * U.u01, b10 = S.s * U.u10;
#pragma unroll
  for (int p = 0; p < NPOL; ++p) {
    const bool is_s = pol_is_s<NPOL>(p, pol_first);
    P.p01[p] = (is_s ? ict : ct) * b01;
    P.p10[p] = (is_s ? ct : ict) * b10;
  }
  P.ks = S.ks;
}

// V0_REAL: the caller knows Im v0 = 0 exactly (the exit column straight from init_exit_columns); the two products with
// that zero are left out -- the same values (finite operands), two FP64 instructions fewer per polarisation
template <int NPOL, int NC, bool V0_REAL = false>
PST_HD void apply_pair_matrix(ChainState<NPOL, NC>& st, const PairMatrix<NPOL>& P) {
#pragma unroll
  for (int p = 0; p < NPOL; ++p) {
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      double* v = st.v[p][c];
      const double v0r = v[0], v0i = v[1], v1r = v[2], v1i = v[3];
      v[0] = fma(P.p00, v0r, P.p01[p] * v1i);
      v[3] = fma(P.p11, v1i, -(P.p10[p] * v0r));
      if constexpr (V0_REAL) {
        v[1] = -(P.p01[p] * v1r);
        v[2] = P.p11 * v1r;
      } else {
        v[1] = fma(P.p00, v0i, -(P.p01[p] * v1r));
        v[2] = fma(P.p11, v1r, P.p10[p] * v0i);
      }
    }
    st.kexp[p] += P.ks;
  }
}

// v <- U^k v with U = B A (A applied first); lg = growth bound of the two layers (growth_log2)
template <int NPOL, int NC>
PST_HD void apply_pair_power(ChainState<NPOL, NC>& st, const TypeRegs<NPOL>& A, const TypeRegs<NPOL>& B, int k, int lg,
                             double ct, double ict, int pol_first) {
  Period U;
  period_form<NPOL>(A, B, false, U);
  PairPower S;
  pair_power_coeffs(U, k, lg, S);
  PairMatrix<NPOL> P;
  pair_power_matrix<NPOL>(U, S, ct, ict, pol_first, P);
  apply_pair_matrix<NPOL, NC>(st, P);
}

// Build + apply in one go (the untyped path).  `lossless` is the MATERIAL's flag (k = 0 at every wavelength), so
// the branch is uniform over the block.
template <int NPOL, int NC>
PST_HD void apply_layer(ChainState<NPOL, NC>& st, const OptRow& o, bool lossless, double d, double ct, double ict,
                        int pol_first) {
  if (lossless) apply_real_layer<NPOL, NC>(st, build_real_layer<NPOL>(o.qr, o.nr, o.inr, d, ct, ict, pol_first));
  else apply_cplx_layer<NPOL, NC>(st, build_cplx_layer<NPOL>(o, d, ct, ict, pol_first));
}

// Pull a power of two out of the state (exact) so that products cannot overflow in deep stop bands.  The largest
// magnitude is found on the high words with integer max (one issue slot each instead of an FP64 compare + selects).
template <int NPOL, int NC>
PST_HD void renormalise(ChainState<NPOL, NC>& st) {
#pragma unroll
  for (int p = 0; p < NPOL; ++p) {
    unsigned m = 0;
#pragma unroll
    for (int c = 0; c < NC; ++c)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const unsigned h = (unsigned)hi_word(st.v[p][c][k]) & 0x7fffffffu;
        m = h > m ? h : m;
      }
    int e = (int)(m >> 20) - 1023;  // floor(log2 max|v|); -1023 for zero/denormal, 1024 for inf/nan
    e = clamp_exp(e);
    const double s = from_words((1023 - e) << 20, 0);
#pragma unroll
    for (int c = 0; c < NC; ++c)
#pragma unroll
      for (int k = 0; k < 4; ++k) st.v[p][c][k] *= s;
    st.kexp[p] += e;
  }
}

// lg >= log2 G with max|M v| <= G max|v| for every layer matrix of the grid.  E = e_n + e_g + 2 >= 0 with e_n, e_g the
// exponents (floor(log2)) of the largest |n|, |1/n| entry and of the largest 1/cos(theta) of the grid (prep_kernel):
// every al = g n sin(phi), be = sin(phi)/(n g) is below 2^E, and the entries of an absorbing layer's mantissa matrix
// (cosh, sinh scaled into [1/2, 5/2]) below 2^(E + 3).
//   lossless   G = 1 + max(|al|, |be|)                       <= 2 max(1, m)  ->  1 + E
//   absorbing  G = |c.re| + |c.im| + max(|m.re| + |m.im|)    <= 4 max(1, m)  ->  2 + E + 3
// Grid-wide, hence block-uniform; it only decides WHEN exact power-of-two rescalings happen, never a result bit.
PST_HD int growth_log2(int E, bool any_lossy) { return E + (any_lossy ? 5 : 1); }
// layout of the device word the typed kernel reads (prep_kernel writes it): bits 0..7 lossy bit per layer type,
// bits 8..19 E, bit 31 some material absorbs
PST_HD int growth_E_of_word(unsigned w) { return (int)((w >> 8) & 0xfffu); }

// Exit-medium columns D_f[:, c]:  s: (1, +-n_f cos)   p: (cos, +-n_f)                       (tm.py:209,212)
template <int NPOL, int NC>
PST_HD void init_exit_columns(ChainState<NPOL, NC>& st, double nfr, double nfi, double ct, int pol_first) {
#pragma unroll
  for (int p = 0; p < NPOL; ++p) {
    const bool is_s = pol_is_s<NPOL>(p, pol_first);
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const double sg = c == 0 ? 1.0 : -1.0;
      st.v[p][c][0] = is_s ? 1.0 : ct;
      st.v[p][c][1] = 0.0;
      st.v[p][c][2] = sg * (is_s ? nfr * ct : nfr);
      st.v[p][c][3] = sg * (is_s ? nfi * ct : nfi);
    }
    st.kexp[p] = 0;
  }
}

template <int NPOL>
PST_HD void init_identity(ChainState<NPOL, 2>& st) {
#pragma unroll
  for (int p = 0; p < NPOL; ++p) {
    st.v[p][0][0] = 1.0; st.v[p][0][1] = 0.0; st.v[p][0][2] = 0.0; st.v[p][0][3] = 0.0;
    st.v[p][1][0] = 0.0; st.v[p][1][1] = 0.0; st.v[p][1][2] = 1.0; st.v[p][1][3] = 0.0;
    st.kexp[p] = 0;
  }
}

// v <- S v for one polarisation, S = [col0 col1] given as 8 doubles (a00, a10 | a01, a11) + exponent.
template <int NPOL, int NC>
PST_HD void apply_segment(ChainState<NPOL, NC>& st, int p, const double (&s)[8], int s_exp) {
  const cplx a00 = {s[0], s[1]}, a10 = {s[2], s[3]}, a01 = {s[4], s[5]}, a11 = {s[6], s[7]};
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    double* v = st.v[p][c];
    const cplx v0 = {v[0], v[1]}, v1 = {v[2], v[3]};
    const cplx n0 = cadd(cmul(a00, v0), cmul(a01, v1));
    const cplx n1 = cadd(cmul(a10, v0), cmul(a11, v1));
    v[0] = n0.re; v[1] = n0.im; v[2] = n1.re; v[3] = n1.im;
  }
  st.kexp[p] += s_exp;
}

// x * 2^e for any int e: two exact power-of-two factors (flushes to 0 / overflows to inf like ldexp, without the
// library routine's denormal handling -- results below 1e-308 are not part of the path's contract)
PST_HD double scale2(double x, int e) {
  e = e > 2000 ? 2000 : (e < -2000 ? -2000 : e);
  const int h = e / 2;
  return x * pow2i(h) * pow2i(e - h);
}
// x * 2^(2h): one power of two applied twice (same result as scale2(x, 2h), fewer integer instructions)
PST_HD double scale_sq(double x, int h) {
  h = h > 1023 ? 1023 : (h < -1023 ? -1023 : h);
  const double f = from_words((h + 1023) << 20, 0);  // h = -1023 gives 0.0
  return x * f * f;
}

struct PointOut {
  double R, T, A;
  double M[8];  // row-major 2x2 complex, only filled when NCOL == 2
};

// Rows of D_0^-1, then t, r, R, T, A for one polarisation                            (tm.py:482-486, 568)
//   D_0^-1 = 1/2 [[w0, w1], [w0, -w1]]:  s: w0 = 1, w1 = 1/(n_0 cos)    p: w0 = 1/cos, w1 = 1/n_0
template <int NPOL, int NCOL>
PST_HD PointOut finish_point_det(const ChainState<NPOL, NCOL>& fin, int p, double in0r, double in0i, double det_re,
                                 double ict, int pol_first, unsigned flags) {
  const bool is_s = pol_is_s<NPOL>(p, pol_first);
  const double w0 = is_s ? 1.0 : ict;
  const cplx w1 = is_s ? cplx{in0r * ict, in0i * ict} : cplx{in0r, in0i};
  // The 1/2 of D_0^-1 is an exact power of two: without the matrix output it is left out of M00 and M10 (it cancels in
  // R) and returned to T through the exponent -- the same bits with eight multiplications fewer per polarisation.
  constexpr double kHalf = NCOL == 2 ? 0.5 : 1.0;
  cplx m[NCOL][2];
#pragma unroll
  for (int c = 0; c < NCOL; ++c) {
    const double* v = fin.v[p][c];
    const cplx a = {kHalf * w0 * v[0], kHalf * w0 * v[1]};
    const cplx b = cscale(cmul(w1, {v[2], v[3]}), kHalf);
    m[c][0] = cadd(a, b);  // M[0][c]
    m[c][1] = csub(a, b);  // M[1][c]
  }
  PointOut out;
  const double den = fma(m[0][0].re, m[0][0].re, m[0][0].im * m[0][0].im);  // |M00|^2 (scaled by 2^-2kexp)
  const double num = fma(m[0][1].re, m[0][1].re, m[0][1].im * m[0][1].im);  // |M10|^2
  const double iden = recip_nr(den);
  out.R = num * iden;
  out.T = scale_sq(det_re * iden, (NCOL == 2 ? 0 : 1) - fin.kexp[p]);
  if constexpr (NCOL == 2) {
    if (flags & 4u) {
      // numeric determinant like np.linalg.det (tm.py:486); the 2^kexp factors cancel
      const cplx det = csub(cmul(m[0][0], m[1][1]), cmul(m[1][0], m[0][1]));
      out.T = det.re * iden;
    }
    const int e = fin.kexp[p];
    out.M[0] = scale2(m[0][0].re, e); out.M[1] = scale2(m[0][0].im, e);
    out.M[2] = scale2(m[1][0].re, e); out.M[3] = scale2(m[1][0].im, e);
    out.M[4] = scale2(m[0][1].re, e); out.M[5] = scale2(m[0][1].im, e);
    out.M[6] = scale2(m[1][1].re, e); out.M[7] = scale2(m[1][1].im, e);
  }
  out.A = (1.0 - out.R) - out.T;
  return out;
}

// Same-angle model of the reference: det M = a_last / a_first = n_f / n_0 for both polarisations (the cos factors cancel)
template <int NPOL, int NCOL>
PST_HD PointOut finish_point(const ChainState<NPOL, NCOL>& fin, int p, double in0r, double in0i, double nfr, double nfi,
                             double ict, int pol_first, unsigned flags) {
  const double det_re = fma(nfr, in0r, -(nfi * in0i));
  return finish_point_det<NPOL, NCOL>(fin, p, in0r, in0i, det_re, ict, pol_first, flags);
}

// ---- refraction mode (PST_FLAG_SNELL) ---------------------------------------------------
// The reference keeps the incidence angle in every layer.  With Snell refraction the conserved quantity is the
// tangential index s = n_0 sin(theta_0); in the reference's notation layer j then has
//   cos(theta_j) = sqrt(1 - s^2 / n_j^2)        (complex in absorbing layers and beyond the critical angle),
//   phi_j = ((n_j w)(1/c) cos(theta_j)) d_j,    a_j = n_j cos(theta_j) (s)  or  n_j / cos(theta_j) (p),
//   M_j = [[cos phi, -i sin phi / a], [-i a sin phi, cos phi]]     -- the same characteristic matrix (tm.py:236-277),
// evaluated by the same operations in the same order as the absorbing-layer path above, so at normal incidence
// (cos(theta_j) = 1 exactly) the two modes agree bit for bit.  M_j is even in the sign of the root; the exit medium
// needs the root whose forward wave decays: Im(n cos(theta)) >= 0 (k > 0 absorbs in the reference, tm.py:217-234).
PST_HD cplx csqrt_principal(cplx w) {
  const double m = hypot(w.re, w.im);
  if (m == 0.0) return {0.0, 0.0};
  const double t = sqrt(0.5 * (m + fabs(w.re)));
  const double u = 0.5 * w.im / t;
  if (w.re >= 0.0) return {t, u};
  return {fabs(u), w.im < 0.0 ? -t : t};
}

struct SnellAngle {
  cplx c, ic;  // cos(theta_j), 1 / cos(theta_j)
  cplx kzn;    // n_j cos(theta_j): normal index, Im >= 0
};

// s2 = (n_0 sin(theta_0))^2.  An exact zero of the cosine (grazing propagation inside the layer) is nudged so that
// sin(phi)/a keeps its finite limit.
PST_HD SnellAngle snell_angle(const OptRow& o, cplx s2) {
  const cplx n = {o.nr, o.ni}, inv_n = {o.inr, o.ini};
  const cplx r = cmul(s2, cmul(inv_n, inv_n));
  SnellAngle g;
  g.c = csqrt_principal({1.0 - r.re, -r.im});
  if (g.c.re == 0.0 && g.c.im == 0.0) g.c = {0.0, 1e-150};
  g.kzn = cmul(n, g.c);
  if (g.kzn.im < 0.0 || (g.kzn.im == 0.0 && g.kzn.re < 0.0)) {
    g.c = {-g.c.re, -g.c.im};
    g.kzn = {-g.kzn.re, -g.kzn.im};
  }
  g.ic = crecip(g.c);
  return g;
}

template <int NPOL>
PST_HD CplxLayer<NPOL> build_snell_layer(const OptRow& o, double d, cplx s2, int pol_first) {
  const SnellAngle g = snell_angle(o, s2);
  const cplx kx = cmul({o.qr, o.qi}, g.c);  // (n w)(1/c) cos(theta_j)
  double sx, cx, ch, sh;
  sincos_fast(mul_rn(kx.re, d), sx, cx);
  CplxLayer<NPOL> m;
  m.sc = cosh_sinh_scaled(mul_rn(kx.im, d), ch, sh);
  m.cphi = {cx * ch, -(sx * sh)};
  const cplx sphi = {sx * ch, cx * sh};
  const cplx sn = cmul(sphi, {o.nr, o.ni});
  const cplx si = cmul(sphi, {o.inr, o.ini});
#pragma unroll
  for (int p = 0; p < NPOL; ++p) {
    const bool is_s = pol_is_s<NPOL>(p, pol_first);
    const cplx as = cmul(sn, is_s ? g.c : g.ic), is = cmul(si, is_s ? g.ic : g.c);
    m.m10[p] = {as.im, -as.re};  // -i a sin(phi)
    m.m01[p] = {is.im, -is.re};  // -i sin(phi) / a
  }
  return m;
}

// Exit-medium columns D_f[:, c] with the refracted angle:  s: (1, +-n_f cos(theta_f))   p: (cos(theta_f), +-n_f)
template <int NPOL, int NC>
PST_HD void init_exit_columns_snell(ChainState<NPOL, NC>& st, double nfr, double nfi, const SnellAngle& gf,
                                    int pol_first) {
#pragma unroll
  for (int p = 0; p < NPOL; ++p) {
    const bool is_s = pol_is_s<NPOL>(p, pol_first);
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const double sg = c == 0 ? 1.0 : -1.0;
      st.v[p][c][0] = is_s ? 1.0 : gf.c.re;
      st.v[p][c][1] = is_s ? 0.0 : gf.c.im;
      st.v[p][c][2] = sg * (is_s ? gf.kzn.re : nfr);
      st.v[p][c][3] = sg * (is_s ? gf.kzn.im : nfi);
    }
    st.kexp[p] = 0;
  }
}

// Tangential index squared for an incidence medium n_0 at cos(theta_0) = ct, and Re det M = Re(n_f cos(theta_f) /
// (n_0 cos(theta_0))) (both polarisations; T = Re(det M) |t|^2 as tm.py:486)
PST_HD cplx snell_s2(double n0r, double n0i, double ct) {
  const double sin2 = fmax(fma(-ct, ct, 1.0), 0.0);
  return cscale(cmul({n0r, n0i}, {n0r, n0i}), sin2);
}
PST_HD double snell_det_re(const SnellAngle& gf, double in0r, double in0i, double ict) {
  return fma(gf.kzn.re, in0r, -(gf.kzn.im * in0i)) * ict;
}

// q, n, 1/n of one (material, wavelength) in the reference's rounding order (tm.py:269-270; numpy divides a
// complex by a real by multiplying with the rounded reciprocal, see pst_math.cuh)
PST_HD OptRow make_opt_row(double wl, double nr, double ni) {
  const double omega = div_rn(kTwoPiC, wl);
  const cplx inv = crecip({nr, ni});
  OptRow r;
  r.qr = mul_rn(mul_rn(nr, omega), kInvC0);
  r.qi = mul_rn(mul_rn(ni, omega), kInvC0);
  r.nr = nr; r.ni = ni; r.inr = inv.re; r.ini = inv.im;
  return r;
}

}  // namespace pst
