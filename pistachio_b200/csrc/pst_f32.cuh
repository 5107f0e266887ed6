// Optional fp32 mode of the point kernel (PST_FLAG_FP32; north_star: "<= 1e-5 for an optional fp32 mode").
//
// What stays in fp64: the propagation phase in the reference's rounding order and its range reduction (a cavity
// amplifies phase error by its finesse, and fp32 cannot even hold phi ~ 100 rad to 1e-5), the table values, and the
// epilogue (t, r, R, T, A).  What runs in fp32: the sin/cos/exp polynomials on the reduced argument, the layer-matrix
// entries and the ordered chain.  The two polarisations of a point live in the two halves of packed f32x2 registers
// (Blackwell FFMA2 / FMUL2: one instruction steps both polarisations), because in the reference's model they differ
// only in the admittance factor g = cos(theta) vs 1/cos(theta).
// One thread per (lambda, theta) point, layer loop without CSE (the fp64 typed path is the fast one for periodic
// stacks; this mode pays off on stacks without repetition).
#pragma once
#include "pst_bulk.cuh"
#include "pst_chain.cuh"

namespace pst {

#ifdef __CUDACC__
__device__ __forceinline__ float2 splat2(float a) { return make_float2(a, a); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }

// sin and cos of a double phase as floats: fp64 Cody-Waite reduction to |r| <= pi/4 (two terms: the third is below
// fp32 resolution), fp32 minimax polynomials (Cephes sinf/cosf, < 1 ulp on the reduced range), quadrant fix-up.
__device__ __forceinline__ void sincos_f32(double x, float& s, float& c) {
  if (!sincos_in_fast_range(x)) {
    double sd, cd;
    sincos_slow(x, &sd, &cd);
    s = (float)sd;
    c = (float)cd;
    return;
  }
  const double kMagic = 6755399441055744.0;
  const double t = fma(x, PST_SC(12), kMagic);
  const double kd = t - kMagic;
  const int q = (int)__double_as_longlong(t);
  double rd = fma(kd, PST_SC(13), x);
  rd = fma(kd, PST_SC(14), rd);
  const float r = (float)rd;
  const float z = r * r;
  float ps = fmaf(-1.9515295891e-4f, z, 8.3321608736e-3f);
  float pc = fmaf(2.443315711809948e-5f, z, -1.388731625493765e-3f);
  ps = fmaf(ps, z, -1.6666654611e-1f);
  pc = fmaf(pc, z, 4.166664568298827e-2f);
  const float sr = fmaf(r * z, ps, r);
  const float cr = fmaf(z * z, pc, fmaf(z, -0.5f, 1.0f));
  const float ss = (q & 1) ? cr : sr;
  const float cc = (q & 1) ? sr : cr;
  s = __int_as_float(__float_as_int(ss) ^ ((q & 2) << 30));
  c = __int_as_float(__float_as_int(cc) ^ (((q + 1) & 2) << 30));
}

// cosh(y) = 2^sc ch, sinh(y) = 2^sc sh with ch, sh of order one (same construction as cosh_sinh_scaled, fp64 argument
// reduction, fp32 even/odd Taylor split of e^{+-r}, |r| <= ln2/2)
__device__ __forceinline__ int cosh_sinh_f32(double y, float& ch, float& sh) {
  double kd = rint(y * PST_EC(0));
  kd = fmin(fmax(kd, -1.0e9), 1.0e9);
  const float r = (float)fma(kd, PST_EC(2), fma(kd, PST_EC(1), y));
  const int k = (int)kd;
  const float r2 = r * r;
  float e = fmaf(1.0f / 720.0f, r2, 1.0f / 24.0f);
  e = fmaf(e, r2, 0.5f);
  e = fmaf(e, r2, 1.0f);
  float o = fmaf(1.0f / 5040.0f, r2, 1.0f / 120.0f);
  o = fmaf(o, r2, 1.0f / 6.0f);
  o = fmaf(o, r2, 1.0f);
  const float ro = r * o;
  const int ka = k < 0 ? -k : k;
  const float f = ka > 60 ? 0.0f : __int_as_float((127 - 2 * ka) << 23);  // 2^(-2|k|)
  const float sg = k >= 0 ? 1.0f : -1.0f;
  const float opf = 1.0f + f, omf = 1.0f - f;
  ch = fmaf(sg * ro, omf, e * opf);
  sh = fmaf(ro, opf, sg * e * omf);
  return ka - 1;
}

// state of one point: v = (v0, v1) complex per polarisation; .x = s-wave, .y = p-wave
struct PairState {
  float2 v0r, v0i, v1r, v1i;
  int kexp[2];
};

__device__ __forceinline__ void renormalise_f32(PairState& st) {
  const unsigned mx = max(max(__float_as_uint(st.v0r.x) & 0x7fffffffu, __float_as_uint(st.v0i.x) & 0x7fffffffu),
                          max(__float_as_uint(st.v1r.x) & 0x7fffffffu, __float_as_uint(st.v1i.x) & 0x7fffffffu));
  const unsigned my = max(max(__float_as_uint(st.v0r.y) & 0x7fffffffu, __float_as_uint(st.v0i.y) & 0x7fffffffu),
                          max(__float_as_uint(st.v1r.y) & 0x7fffffffu, __float_as_uint(st.v1i.y) & 0x7fffffffu));
  int ex = (int)(mx >> 23) - 127, ey = (int)(my >> 23) - 127;  // floor(log2 max|v|); -127 for zero, 128 for inf/nan
  ex = ex > 100 ? 100 : (ex < -100 ? 0 : ex);
  ey = ey > 100 ? 100 : (ey < -100 ? 0 : ey);
  const float2 sc = make_float2(__int_as_float((127 - ex) << 23), __int_as_float((127 - ey) << 23));
  st.v0r = mul2(st.v0r, sc); st.v0i = mul2(st.v0i, sc);
  st.v1r = mul2(st.v1r, sc); st.v1i = mul2(st.v1i, sc);
  st.kexp[0] += ex;
  st.kexp[1] += ey;
}

constexpr int kRenormEveryF32 = 4;  // fp32 range 2^+-126: |a| of a metal reaches ~2^5 per layer

template <bool STAGE>
__global__ void __launch_bounds__(128)
tmm_chain_f32_kernel(const GridParams prm) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int PB = 128;
  const unsigned npts = (unsigned)prm.points, nth = (unsigned)prm.n_theta;
  const unsigned p0 = blockIdx.x * (unsigned)PB;
  const unsigned p_last = min(p0 + (unsigned)PB, npts) - 1u;
  const int l_lo = prm.il0 + (int)(p0 / nth);
  const int TL = prm.il0 + (int)(p_last / nth) - l_lo + 1;
  const int L = prm.n_layers;
  unsigned char* s_opt = smem_raw;
  LayerEntry* s_layers = reinterpret_cast<LayerEntry*>(smem_raw + (size_t)prm.tl_rows * prm.opt_row_bytes);
  if constexpr (STAGE)
    stage_block(prm.use_bulk != 0, s_opt, reinterpret_cast<const unsigned char*>(prm.opt + (size_t)l_lo * prm.n_mat * 6), TL,
                prm.n_mat * 48, prm.opt_row_bytes, reinterpret_cast<unsigned char*>(s_layers),
                reinterpret_cast<const unsigned char*>(prm.layers), L * (int)sizeof(LayerEntry), threadIdx.x, PB);
  if (p0 + threadIdx.x >= npts) return;
  const unsigned p = p0 + threadIdx.x;
  const int il_loc = (int)(p / nth);
  const int it = (int)(p - (unsigned)il_loc * nth);
  const int il = prm.il0 + il_loc;
  const double2 tr = __ldg(prm.trig + it);
  const double ct = tr.x, ict = tr.y;
  const double d_sweep = prm.sweep_d ? __ldg(prm.sweep_d + blockIdx.y) : 0.0;
  const unsigned char* optp = STAGE ? (s_opt + (size_t)(il - l_lo) * prm.opt_row_bytes)
                                    : reinterpret_cast<const unsigned char*>(prm.opt + (size_t)il * prm.n_mat * 6);
  const LayerEntry* lay = STAGE ? s_layers : prm.layers;
  auto opt_row = [&](int mat) {
    const double2* o = reinterpret_cast<const double2*>(optp + (size_t)mat * 48);
    const double2 a = o[0], b = o[1], c = o[2];
    OptRow r;
    r.qr = a.x; r.nr = a.y; r.inr = b.x; r.ni = b.y; r.qi = c.x; r.ini = c.y;
    return r;
  };
  const OptRow exit_row = opt_row(lay[L - 1].mat);
  const float ctf = (float)ct, ictf = (float)ict;
  const float2 G = make_float2(ctf, ictf), Gi = make_float2(ictf, ctf);      // admittance factor g and 1/g per polarisation
  const float2 nG = make_float2(-ctf, -ictf), nGi = make_float2(-ictf, -ctf);

  // exit-medium column D_f[:, 0]:  s: (1, n_f cos)   p: (cos, n_f)
  PairState st;
  st.v0r = make_float2(1.0f, ctf);
  st.v0i = make_float2(0.0f, 0.0f);
  st.v1r = make_float2((float)(exit_row.nr * ct), (float)exit_row.nr);
  st.v1i = make_float2((float)(exit_row.ni * ct), (float)exit_row.ni);
  st.kexp[0] = st.kexp[1] = 0;

  // Norm drift.  A lossless layer matrix has determinant 1; its fp32 entries give 1 + delta_j with |delta_j| ~ 1e-7, and
  // identical layers (a Bragg mirror) repeat the same delta, so the product's norm drifts linearly with the depth: |M00|^2
  // and with it T is off by prod(1 + delta_j) while R, a ratio, is not (measured on 256 low-contrast layers: T 2.9e-5, R
  // 6e-7).  The deltas are known exactly from the rounded entries, so they are summed and divided out of T at the end.
  float2 det_drift = make_float2(0.0f, 0.0f);
  int since = 0;
  for (int j = L - 2; j >= 1; --j) {
    const LayerEntry le = lay[j];
    const double d = (j == prm.sweep_layer) ? d_sweep : le.d;
    const double2* o = reinterpret_cast<const double2*>(optp + (size_t)le.mat * 48);
    const double2 a = o[0], b = o[1];  // (q.re, n.re), ((1/n).re, n.im)
    float sx, cx;
    sincos_f32(layer_phase(a.x, ct, d), sx, cx);
    if (le.type == 0) {  // lossless material: M = [[c, -i be], [-i al, c]], real entries
      const float sn = sx * (float)a.y, si = sx * (float)b.x;
      const float2 c2 = splat2(cx);
      const float2 al = mul2(splat2(sn), G), nal = mul2(splat2(sn), nG);
      const float2 be = mul2(splat2(si), Gi), nbe = mul2(splat2(si), nGi);
      {
        // det M_j - 1 = c^2 + al be - 1 of the ROUNDED entries, to ~1e-14 with fp32 error-free products: the two big terms
        // cancel exactly (Sterbenz) once 1 is taken from the one that is >= 1/2
        const float2 pc = mul2(c2, c2), ec = fma2(c2, c2, make_float2(-pc.x, -pc.y));
        const float2 pq = mul2(al, be), eq = fma2(al, be, make_float2(-pq.x, -pq.y));
        const bool cbig = pc.x >= 0.5f;
        const float2 hi = cbig ? pc : pq, lo = cbig ? pq : pc;
        const float2 t = __fadd2_rn(__fadd2_rn(hi, make_float2(-1.0f, -1.0f)), lo);
        det_drift = __fadd2_rn(det_drift, __fadd2_rn(t, __fadd2_rn(ec, eq)));
      }
      const float2 n0r = fma2(c2, st.v0r, mul2(be, st.v1i));
      const float2 n0i = fma2(c2, st.v0i, mul2(nbe, st.v1r));
      const float2 n1r = fma2(c2, st.v1r, mul2(al, st.v0i));
      const float2 n1i = fma2(c2, st.v1i, mul2(nal, st.v0r));
      st.v0r = n0r; st.v0i = n0i; st.v1r = n1r; st.v1i = n1i;
    } else {
      const double2 c = o[2];  // (q.im, (1/n).im)
      float ch, sh;
      const int sc = cosh_sinh_f32(layer_phase(c.x, ct, d), ch, sh);
      // cos(x+iy) = cos x cosh y - i sin x sinh y ; sin(x+iy) = sin x cosh y + i cos x sinh y   (both times 2^sc)
      const float cr = cx * ch, ci = -(sx * sh);
      const float sr = sx * ch, sim = cx * sh;
      const float nr = (float)a.y, ni = (float)b.y, inr = (float)b.x, ini = (float)c.y;
      const float snr = fmaf(sr, nr, -(sim * ni)), sni = fmaf(sr, ni, sim * nr);      // sin(phi) n
      const float sir = fmaf(sr, inr, -(sim * ini)), sii = fmaf(sr, ini, sim * inr);  // sin(phi) / n
      // m10 = -i g sin(phi) n = (g sn.im, -g sn.re);  m01 = -i sin(phi) / (g n) = (gi si.im, -gi si.re)
      const float2 m10r = mul2(splat2(sni), G), m10i = mul2(splat2(snr), nG), nm10i = mul2(splat2(snr), G);
      const float2 m01r = mul2(splat2(sii), Gi), m01i = mul2(splat2(sir), nGi), nm01i = mul2(splat2(sir), Gi);
      const float2 cr2 = splat2(cr), ci2 = splat2(ci), nci2 = splat2(-ci);
      const float2 n0r = fma2(cr2, st.v0r, fma2(nci2, st.v0i, fma2(m01r, st.v1r, mul2(nm01i, st.v1i))));
      const float2 n0i = fma2(cr2, st.v0i, fma2(ci2, st.v0r, fma2(m01r, st.v1i, mul2(m01i, st.v1r))));
      const float2 n1r = fma2(m10r, st.v0r, fma2(nm10i, st.v0i, fma2(cr2, st.v1r, mul2(nci2, st.v1i))));
      const float2 n1i = fma2(m10r, st.v0i, fma2(m10i, st.v0r, fma2(cr2, st.v1i, mul2(ci2, st.v1r))));
      st.v0r = n0r; st.v0i = n0i; st.v1r = n1r; st.v1i = n1i;
      st.kexp[0] += sc;
      st.kexp[1] += sc;
    }
    if (++since == kRenormEveryF32) { renormalise_f32(st); since = 0; }
  }
  renormalise_f32(st);

  // ---- epilogue in fp64 (t, r, R, T, A as in the fp64 kernels)
  ChainState<2, 1> fin;
  fin.v[0][0][0] = st.v0r.x; fin.v[0][0][1] = st.v0i.x; fin.v[0][0][2] = st.v1r.x; fin.v[0][0][3] = st.v1i.x;
  fin.v[1][0][0] = st.v0r.y; fin.v[1][0][1] = st.v0i.y; fin.v[1][0][2] = st.v1r.y; fin.v[1][0][3] = st.v1i.y;
  fin.kexp[0] = st.kexp[0];
  fin.kexp[1] = st.kexp[1];
  const OptRow in_row = opt_row(lay[0].mat);
  const size_t plane = (size_t)prm.points;
  const int npol_out = (prm.pol_mask == 3u) ? 2 : 1;
  int slot = 0;
#pragma unroll
  for (int pp = 0; pp < 2; ++pp) {
    if (!(prm.pol_mask & (1u << pp))) continue;
    PointOut r = finish_point<2, 1>(fin, pp, in_row.inr, in_row.ini, exit_row.nr, exit_row.ni, ict, 0, prm.flags);
    const double dd = (double)(pp == 0 ? det_drift.x : det_drift.y);
    r.T *= fma(dd, fma(dd, 0.5, 1.0), 1.0);  // prod(1 + delta_j) = exp(sum delta_j) to second order
    r.A = (1.0 - r.R) - r.T;
    store_point<false>(prm, ((size_t)blockIdx.y * npol_out + slot) * plane + (size_t)p, r.R, r.T, r.A);
    ++slot;
  }
}
#endif  // __CUDACC__

}  // namespace pst
