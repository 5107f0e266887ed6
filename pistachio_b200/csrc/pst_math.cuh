// Device-side fp64 building blocks of the TMM path (sm_100a).
//
// Two rules shape this file:
//  (1) The propagation phase must carry the reference's rounding sequence, because a
//      cavity of finesse F amplifies a phase error by F.  The reference evaluates
//          omega = (2*pi*c)/lambda                                  tm.py:269
//          kx    = ((n*omega) * (1/c)) * cos(theta)                 tm.py:270 (numpy divides a
//                   complex by a real by multiplying with the rounded reciprocal)
//          arg   = (Im kx * d, -Re kx * d)                          tm.py:234
//      each step rounded once.  mul_rn/div_rn below forbid FMA contraction for these.
//  (2) Everything after the phase is free of catastrophic amplification, so it is
//      written for the FP64 pipe: FMA everywhere, no divisions in the layer loop,
//      exponentials split into mantissa and a power of two that is carried as an
//      integer so that thick absorbers and deep stop bands cannot overflow.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

// Device math is also compilable for the host (nvcc host pass) so that tests/host_emul can run the *same*
// layer-step / epilogue source on the CPU box before GPU time is spent.  That harness is test infrastructure;
// libpistachio_b200.so never calls these functions from host code.
#define PST_HD __host__ __device__ __forceinline__

namespace pst {

#ifdef __CUDA_ARCH__
PST_HD double mul_rn(double a, double b) { return __dmul_rn(a, b); }
PST_HD double add_rn(double a, double b) { return __dadd_rn(a, b); }
PST_HD double sub_rn(double a, double b) { return __dsub_rn(a, b); }
PST_HD double div_rn(double a, double b) { return __ddiv_rn(a, b); }
PST_HD int hi_word(double x) { return __double2hiint(x); }
PST_HD double from_words(int hi, int lo) { return __hiloint2double(hi, lo); }
#else
// host emulation: separate statements through volatiles keep the compiler from contracting a*b+c
PST_HD double mul_rn(double a, double b) { volatile double r = a * b; return r; }
PST_HD double add_rn(double a, double b) { volatile double r = a + b; return r; }
PST_HD double sub_rn(double a, double b) { volatile double r = a - b; return r; }
PST_HD double div_rn(double a, double b) { volatile double r = a / b; return r; }
PST_HD int hi_word(double x) { long long u; memcpy(&u, &x, 8); return (int)(u >> 32); }
PST_HD double from_words(int hi, int lo) {
  long long u = ((long long)(unsigned)hi << 32) | (unsigned)lo; double x; memcpy(&x, &u, 8); return x;
}
#endif

#ifdef __CUDA_ARCH__
PST_HD double xor_sign(double x, int mask) { return __hiloint2double(__double2hiint(x) ^ mask, __double2loint(x)); }
#else
PST_HD double xor_sign(double x, int mask) {
  long long u; memcpy(&u, &x, 8); u ^= (long long)(unsigned)mask << 32; memcpy(&x, &u, 8); return x;
}
#endif

constexpr double kC0 = 299792458.0;                        // scipy.constants.c
constexpr double kTwoPiC = (2.0 * 3.141592653589793) * kC0;  // (2*np.pi)*c, one rounding, folded on the host
constexpr double kInvC0 = 1.0 / kC0;                       // numpy's reciprocal-scale in complex/real division

// 2^e as a double for e in [-1022, 1023]; 0 below, +inf above.
PST_HD double pow2i(int e) {
  if (e < -1022) return 0.0;
  if (e > 1023) return from_words(0x7ff00000, 0);
  return from_words((e + 1023) << 20, 0);
}

// Unbiased exponent of |x| (floor(log2|x|)) for normal x; 0 for zero/denormal, 1024 for inf/nan.
PST_HD int exponent_of(double x) {
  int hi = hi_word(x);
  int e = (hi >> 20) & 0x7ff;
  return e == 0 ? 0 : e - 1023;
}

// y = k*ln2 + r, |r| <= ln2/2:  e^{+-r} = E(r^2) +- r*O(r^2) with one shared even/odd Taylor
// split (degree 13, truncation < 4e-18).  Returns even = E, odd = r*O; k is returned, not applied.
#define PST_EXP_COEFS                                                                                          \
  {1.4426950408889634, -6.93147180369123816490e-01, -1.90821492927058770002e-10, 1.0 / 479001600.0,            \
   1.0 / 3628800.0, 1.0 / 40320.0, 1.0 / 720.0, 1.0 / 24.0, 1.0 / 6227020800.0, 1.0 / 39916800.0,             \
   1.0 / 362880.0, 1.0 / 5040.0, 1.0 / 120.0, 1.0 / 6.0}
// 0: log2(e); 1, 2: -ln2 split (low 21 bits of the high part are zero); 3..7: 1/12! .. 1/4!; 8..13: 1/13! .. 1/3!
#ifdef __CUDACC__
__constant__ double c_exp[14] = PST_EXP_COEFS;
#endif
static const double h_exp[14] = PST_EXP_COEFS;
#ifdef __CUDA_ARCH__
#define PST_EC(i) c_exp[i]
#else
#define PST_EC(i) h_exp[i]
#endif
PST_HD void exp_split(double y, double& even, double& odd, int& k) {
  // |y| >= 2^30 (not NaN) is clamped to +-2^30 on the high word: the power of two then saturates the caller's exponent
  // bookkeeping instead of overflowing it (five integer instructions; fmin/fmax on doubles cost fifteen)
  {
    const int hy = hi_word(y);
    if ((unsigned)((hy & 0x7fffffff) - 0x41d00000) < (unsigned)(0x7ff00000 - 0x41d00000)) y = from_words((hy & 0x80000000) | 0x41d00000, 0);
  }
  const double kd = rint(y * PST_EC(0));
  double r = fma(kd, PST_EC(1), y);
  r = fma(kd, PST_EC(2), r);
  k = (int)kd;
  double r2 = r * r;
  double e = PST_EC(3);                    // 1/12!
  e = fma(e, r2, PST_EC(4));               // 1/10!
  e = fma(e, r2, PST_EC(5));               // 1/8!
  e = fma(e, r2, PST_EC(6));               // 1/6!
  e = fma(e, r2, PST_EC(7));               // 1/4!
  e = fma(e, r2, 0.5);                     // 1/2!
  e = fma(e, r2, 1.0);
  double o = PST_EC(8);                    // 1/13!
  o = fma(o, r2, PST_EC(9));               // 1/11!
  o = fma(o, r2, PST_EC(10));              // 1/9!
  o = fma(o, r2, PST_EC(11));              // 1/7!
  o = fma(o, r2, PST_EC(12));              // 1/5!
  o = fma(o, r2, PST_EC(13));              // 1/3!
  o = fma(o, r2, 1.0);
  even = e;
  odd = r * o;
}

// cosh(y) = 2^s * ch, sinh(y) = 2^s * sh with ch, sh of order one; returns s = |k| - 1.
// With f = 2^(-2|k|):  cosh|y| = 2^(|k|-1) [E(1+f) + rO(1-f)],  sinh|y| = 2^(|k|-1) [E(1-f) + rO(1+f)].
// Never overflows (the huge factor stays in the integer) and keeps sinh accurate for tiny y (f = 1).
PST_HD int cosh_sinh_scaled(double y, double& ch, double& sh) {
  double e, ro;
  int k;
  exp_split(y, e, ro, k);
  int ka = k < 0 ? -k : k;
  double f = pow2i(ka > 600 ? -1200 : -2 * ka);
  double sg = k >= 0 ? 1.0 : -1.0;  // for k < 0 evaluate at -y: r -> -r, sinh flips sign
  double opf = 1.0 + f, omf = 1.0 - f;
  ch = fma(sg * ro, omf, e * opf);
  sh = fma(ro, opf, sg * e * omf);
  return ka - 1;
}

// ---- sincos for the layer loop ---------------------------------------------------------------------------
// Coefficients live in the constant bank so that DFMA takes them as c[bank][offset] operands: the library
// sincos() materialises each 64-bit constant with two moves, which made kernel v1 issue-bound (profiles/r01_v1).
// 0..5: sin minimax S1..S6, 6..11: cos minimax C1..C6 on |r| <= pi/4 (fdlibm __kernel_sin/__kernel_cos, error < 2^-58)
#define PST_SINCOS_COEFS                                                                                       \
  {-1.66666666666666324348e-01, 8.33333333332248946124e-03, -1.98412698298579493134e-04,                       \
   2.75573137070700676789e-06, -2.50507602534068634195e-08, 1.58969099521155010221e-10,                        \
   4.16666666666666019037e-02, -1.38888888888741095749e-03, 2.48015872894767294178e-05,                        \
   -2.75573143513906633035e-07, 2.08757232129817482790e-09, -1.13596475577881948265e-11,                       \
   6.36619772367581382433e-01, -1.5707963267948966, -6.123233995736757e-17, -8.478427660368898e-32}
// 12: 2/pi; 13..15: -(pi/2) split into three doubles (0x3ff921fb54442d18, 0x3c91a62633145c00, 0x397b839a252049c0);
// the middle one is truncated (trailing zero bits) so that kd*mid stays exact for |kd| < 2^30
#ifdef __CUDACC__
__constant__ double c_sincos[16] = PST_SINCOS_COEFS;
#endif
static const double h_sincos[16] = PST_SINCOS_COEFS;
#ifdef __CUDA_ARCH__
#define PST_SC(i) c_sincos[i]
#else
#define PST_SC(i) h_sincos[i]
#endif

// sin and cos of x.  |x| < 2^30: k = rint(x 2/pi) by the magic-number add (no F2I/I2F), two-term Cody-Waite
// reduction with FMA (pi/2 split into doubles whose leading two are used), two degree-6 polynomials in r^2, quadrant fix-up with integer
// selects.  Larger |x| (and inf/nan) take the library's Payne-Hanek path.  Accuracy ~1 ulp.
#ifdef __CUDA_ARCH__
__device__ __noinline__ void sincos_slow(double x, double* s, double* c) { sincos(x, s, c); }  // cold: keep it out of line
#else
inline void sincos_slow(double x, double* s, double* c) { sincos(x, s, c); }
#endif
PST_HD bool sincos_in_fast_range(double x) { return fabs(x) < 1073741824.0; }  // false for nan too
PST_HD void sincos_core(double x, double& s, double& c);
PST_HD void sincos_fast(double x, double& s, double& c) {
  if (!sincos_in_fast_range(x)) {
    double ts, tc;  // the address-taken pair lives on the cold branch only: s and c themselves stay in registers
    sincos_slow(x, &ts, &tc);
    s = ts;
    c = tc;
    return;
  }
  sincos_core(x, s, c);
}
// straight-line part (no branches): several independent calls interleave in one basic block
PST_HD void sincos_core(double x, double& s, double& c) {
  const double kMagic = 6755399441055744.0;  // 1.5 * 2^52
  const double t = fma(x, PST_SC(12), kMagic);
  const double kd = t - kMagic;
  int q;  // low word of the magic sum = k mod 2^32 (read as a 32-bit register: the quadrant tests stay 32-bit compares)
#ifdef __CUDA_ARCH__
  q = __double2loint(t);
#else
  {
    long long u;
    memcpy(&u, &t, 8);
    q = (int)u;
  }
#endif
  // two-term Cody-Waite: kd * hi is exact in the FMA and the middle word carries pi/2 to 2^-107; the third word
  // (8.5e-32 kd < 6e-23 on the fast range) is below half an ulp of any result that is not itself below 1e-6
  double r = fma(kd, PST_SC(13), x);
  r = fma(kd, PST_SC(14), r);
  const double z = r * r;
  double ps = fma(PST_SC(5), z, PST_SC(4));
  double pc = fma(PST_SC(11), z, PST_SC(10));
  ps = fma(ps, z, PST_SC(3));
  pc = fma(pc, z, PST_SC(9));
  ps = fma(ps, z, PST_SC(2));
  pc = fma(pc, z, PST_SC(8));
  ps = fma(ps, z, PST_SC(1));
  pc = fma(pc, z, PST_SC(7));
  ps = fma(ps, z, PST_SC(0));
  pc = fma(pc, z, PST_SC(6));
  const double sr = fma(r * z, ps, r);
  const double cr = fma(z, fma(z, pc, -0.5), 1.0);
  // quadrant: q&1 swaps, q&2 negates sin, (q+1)&2 negates cos
  const double ss = (q & 1) ? cr : sr;
  const double cc = (q & 1) ? sr : cr;
  // sign flips as XOR on the high word (one integer op each instead of an FP64 negate + select)
  s = xor_sign(ss, (q & 2) << 30);
  c = xor_sign(cc, ((q + 1) & 2) << 30);
}

// 1/d for the epilogues: hardware seed (2^-23) and two Newton steps, every path the same function.  The compiler's 1.0/d
// adds an exactly-rounded last step and a slow-path branch for denormal or huge operands (ten instructions more per
// call); the operands here are |M00|^2 of a freshly rescaled state.  <= 1 ulp; the host emulation divides.
PST_HD double recip_nr(double d) {
#ifdef __CUDA_ARCH__
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  double e = fma(-d, r, 1.0);
  r = fma(r, e, r);
  e = fma(-d, r, 1.0);
  return fma(r, e, r);
#else
  return 1.0 / d;
#endif
}

struct cplx {
  double re, im;
};
PST_HD cplx cmul(cplx a, cplx b) {
  return {fma(a.re, b.re, -(a.im * b.im)), fma(a.re, b.im, a.im * b.re)};
}
PST_HD cplx cadd(cplx a, cplx b) { return {a.re + b.re, a.im + b.im}; }
PST_HD cplx csub(cplx a, cplx b) { return {a.re - b.re, a.im - b.im}; }
PST_HD cplx cscale(cplx a, double s) { return {a.re * s, a.im * s}; }
PST_HD cplx crecip(cplx a) {
  // scaled to dodge overflow of |a|^2; used outside the layer loop only
  double m = fmax(fabs(a.re), fabs(a.im));
  if (m == 0.0) return {1.0 / a.re, 0.0};
  double sr = a.re / m, si = a.im / m;
  double den = (sr * sr + si * si) * m;
  return {sr / den, -si / den};
}
PST_HD cplx cdiv(cplx a, cplx b) { return cmul(a, crecip(b)); }

}  // namespace pst
