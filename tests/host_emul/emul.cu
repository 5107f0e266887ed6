// TEST INFRASTRUCTURE ONLY.  Runs the SAME per-point source the CUDA kernels execute
// (pistachio_b200/csrc/pst_chain.cuh, compiled by nvcc's host pass) in a plain CPU loop, so that the algebra of
// the layer step, the exact rescaling, the ordered segment combine and the epilogue can be checked against the
// oracle on the GPU-less build box.  Built into tests/host_emul/_build/ by tests/test_host_emul.py; never part of
// libpistachio_b200.so and never reachable from the pistachio_b200 package.
#include <vector>
#include "../../pistachio_b200/csrc/pst_chain.cuh"
#include "../../pistachio_b200/csrc/pst_fit.cuh"

using namespace pst;

template <int NPOL, int NCOL>
static void run(int n_wl, int n_theta, int L, const double* wl, const double* cos_theta, const double* n_re,
                const double* n_im, const double* d, int pol_first, int W, unsigned flags, double* R, double* T,
                double* A, double* M) {
  const int Li = L - 2;
  std::vector<char> lossless(L, 1);  // the kernels branch on the MATERIAL's flag: k = 0 at every wavelength
  for (int j = 0; j < L; ++j)
    for (int il = 0; il < n_wl; ++il)
      if (n_im[(size_t)j * n_wl + il] != 0.0) lossless[j] = 0;
  // grid-wide growth bound as prep_kernel forms it: exponents of the largest |n|, |1/n| component and 1/cos(theta)
  int growth_E = 0;
  {
    unsigned hn = 0, hg = 0;
    auto up = [](unsigned& m, double x) { const unsigned h = (unsigned)hi_word(x) & 0x7fffffffu; m = h > m ? h : m; };
    for (int j = 0; j < L; ++j)
      for (int il = 0; il < n_wl; ++il) {
        const OptRow r = make_opt_row(wl[il], n_re[(size_t)j * n_wl + il], n_im[(size_t)j * n_wl + il]);
        up(hn, r.nr); up(hn, r.ni); up(hn, r.inr); up(hn, r.ini);
      }
    for (int it = 0; it < n_theta; ++it) up(hg, 1.0 / cos_theta[it]);
    const int E = (int)(hn >> 20) - 1023 + (int)(hg >> 20) - 1023 + 2;
    growth_E = E < 0 ? 0 : (E > 4095 ? 4095 : E);
  }
  for (int il = 0; il < n_wl; ++il) {
    std::vector<OptRow> rows(L);
    for (int j = 0; j < L; ++j) rows[j] = make_opt_row(wl[il], n_re[(size_t)j * n_wl + il], n_im[(size_t)j * n_wl + il]);
    for (int it = 0; it < n_theta; ++it) {
      const double ct = cos_theta[it], ict = 1.0 / ct;
      ChainState<NPOL, NCOL> fin;
      if (flags & 4096u) {
        // refraction mode (PST_FLAG_SNELL): the loop of tmm_snell_kernel
        const cplx s2 = snell_s2(rows[0].nr, rows[0].ni, ct);
        const SnellAngle gf = snell_angle(rows[L - 1], s2);
        init_exit_columns_snell<NPOL, NCOL>(fin, rows[L - 1].nr, rows[L - 1].ni, gf, pol_first);
        int j = L - 2;
        while (j >= 1) {
          const int j_stop = j - 16 > 0 ? j - 16 : 0;
          for (; j > j_stop; --j) apply_cplx_layer<NPOL, NCOL>(fin, build_snell_layer<NPOL>(rows[j], d[j], s2, pol_first));
          renormalise<NPOL, NCOL>(fin);
        }
        const double det_re = snell_det_re(gf, rows[0].inr, rows[0].ini, ict);
        for (int p = 0; p < NPOL; ++p) {
          const PointOut r = finish_point_det<NPOL, NCOL>(fin, p, rows[0].inr, rows[0].ini, det_re, ict, pol_first, flags);
          const size_t o = ((size_t)p * n_wl + il) * n_theta + it;
          R[o] = r.R; T[o] = r.T; A[o] = r.A;
          if (NCOL == 2 && M) for (int k = 0; k < 8; ++k) M[o * 8 + k] = r.M[k];
        }
        continue;
      }
      if (W <= 1 && (flags & 256u)) {
        // typed path with run-length ops and the closed-form power of repeated lossless pairs (what the kernel does
        // by default for periodic stacks): layer types by (n at this wavelength, thickness)
        std::vector<int> seq;
        for (int j = L - 2; j >= 1; --j) {
          int t = j;
          for (int u = L - 2; u > j; --u)
            if (rows[u].nr == rows[j].nr && rows[u].ni == rows[j].ni && d[u] == d[j]) { t = u; break; }
          seq.push_back(t);
        }
        init_exit_columns<NPOL, NCOL>(fin, rows[L - 1].nr, rows[L - 1].ni, ct, pol_first);
        auto build = [&](int j) {
          TypeRegs<NPOL> t;
          if (lossless[j]) t.set_real(build_real_layer<NPOL>(rows[j].qr, rows[j].nr, rows[j].inr, d[j], ct, ict, pol_first));
          else t.set_cplx(build_cplx_layer<NPOL>(rows[j], d[j], ct, ict, pol_first));
          return t;
        };
        size_t i = 0;
        const size_t n = seq.size();
        while (i < n) {
          if (i + 1 < n) {
            const int ta = seq[i], tb = seq[i + 1];
            int k = 1;
            while (i + 2 * (size_t)k + 1 < n && seq[i + 2 * k] == ta && seq[i + 2 * k + 1] == tb) ++k;
            const TypeRegs<NPOL> A = build(ta), B = build(tb);
            if (lossless[ta] && lossless[tb] && k >= 3) {
              renormalise<NPOL, NCOL>(fin);
              apply_pair_power<NPOL, NCOL>(fin, A, B, k, growth_log2(growth_E, false), ct, ict, pol_first);
            } else {
              for (int r = 0; r < k; ++r) {
                apply_type<NPOL, NCOL>(fin, A, !lossless[ta]);
                apply_type<NPOL, NCOL>(fin, B, !lossless[tb]);
                renormalise<NPOL, NCOL>(fin);
              }
            }
            renormalise<NPOL, NCOL>(fin);
            i += 2 * (size_t)k;
          } else {
            apply_type<NPOL, NCOL>(fin, build(seq[i]), !lossless[seq[i]]);
            i += 1;
          }
        }
        renormalise<NPOL, NCOL>(fin);
      } else if (W <= 1) {
        init_exit_columns<NPOL, NCOL>(fin, rows[L - 1].nr, rows[L - 1].ni, ct, pol_first);
        int since = 0;
        for (int j = L - 2; j >= 1; --j) {
          apply_layer<NPOL, NCOL>(fin, rows[j], lossless[j], d[j], ct, ict, pol_first);
          if (++since == 16) { renormalise<NPOL, NCOL>(fin); since = 0; }
        }
        renormalise<NPOL, NCOL>(fin);
      } else {
        std::vector<ChainState<NPOL, 2>> segs(W);
        for (int w = 0; w < W; ++w) {
          const int seg_q = Li / W, seg_r = Li - seg_q * W;
          const int j_lo = 1 + w * seg_q + (w < seg_r ? w : seg_r), j_hi = j_lo + seg_q + (w < seg_r ? 1 : 0);
          init_identity<NPOL>(segs[w]);
          int since = 0;
          for (int j = j_hi - 1; j >= j_lo; --j) {
            apply_layer<NPOL, 2>(segs[w], rows[j], lossless[j], d[j], ct, ict, pol_first);
            if (++since == 16) { renormalise<NPOL, 2>(segs[w]); since = 0; }
          }
          renormalise<NPOL, 2>(segs[w]);
        }
        init_exit_columns<NPOL, NCOL>(fin, rows[L - 1].nr, rows[L - 1].ni, ct, pol_first);
        for (int w = W - 1; w >= 0; --w) {
          for (int p = 0; p < NPOL; ++p) {
            double s[8];
            for (int c = 0; c < 2; ++c) for (int k = 0; k < 4; ++k) s[c * 4 + k] = segs[w].v[p][c][k];
            apply_segment<NPOL, NCOL>(fin, p, s, segs[w].kexp[p]);
          }
          renormalise<NPOL, NCOL>(fin);
        }
      }
      for (int p = 0; p < NPOL; ++p) {
        const PointOut r = finish_point<NPOL, NCOL>(fin, p, rows[0].inr, rows[0].ini, rows[L - 1].nr, rows[L - 1].ni,
                                                    ict, pol_first, flags);
        const size_t o = ((size_t)p * n_wl + il) * n_theta + it;
        R[o] = r.R; T[o] = r.T; A[o] = r.A;
        if (NCOL == 2 && M) for (int k = 0; k < 8; ++k) M[o * 8 + k] = r.M[k];
      }
    }
  }
}

// pol: 0 = s only, 1 = p only, 2 = both.  n_re/n_im: [L][n_wl].  Outputs [n_pol][n_wl][n_theta].
extern "C" int emul_eval_grid(int n_wl, int n_theta, int L, const double* wl, const double* cos_theta,
                              const double* n_re, const double* n_im, const double* d, int pol, int ncol, int W,
                              unsigned flags, double* R, double* T, double* A, double* M) {
  if (L < 2) return -1;
  if (pol == 2) {
    if (ncol == 2) run<2, 2>(n_wl, n_theta, L, wl, cos_theta, n_re, n_im, d, 0, W, flags, R, T, A, M);
    else run<2, 1>(n_wl, n_theta, L, wl, cos_theta, n_re, n_im, d, 0, W, flags, R, T, A, M);
  } else {
    if (ncol == 2) run<1, 2>(n_wl, n_theta, L, wl, cos_theta, n_re, n_im, d, pol, W, flags, R, T, A, M);
    else run<1, 1>(n_wl, n_theta, L, wl, cos_theta, n_re, n_im, d, pol, W, flags, R, T, A, M);
  }
  return 0;
}

// ---- period power (pst_chain.cuh): the compile-time chain of a pair count against the generic runtime bit loop
// out = (t, s, 1 when the count is inside the compiled range)
extern "C" void emul_pair_power(double x, double s2, int k, int fixed, double* out) {
  Period U;
  U.x = x; U.s2 = s2; U.w = 0.0; U.u01 = 0.0; U.u10 = 0.0;
  PairPower S;
  S.t = S.s = 0.0; S.ks = 0;
  bool ok = true;
  if (fixed) ok = pair_power_small(U, k, S);
  else pair_power_coeffs_t<false>(U, k, log2_floor(k), S);
  out[0] = S.t; out[1] = S.s; out[2] = ok ? 1.0 : 0.0;
}

// ---- elementary functions of the layer loop, exposed for accuracy tests against mpmath
extern "C" void emul_sincos(int n, const double* x, double* s, double* c) {
  for (int i = 0; i < n; ++i) sincos_fast(x[i], s[i], c[i]);
}
extern "C" void emul_cosh_sinh(int n, const double* y, double* ch, double* sh, int* sc) {
  for (int i = 0; i < n; ++i) sc[i] = cosh_sinh_scaled(y[i], ch[i], sh[i]);
}

// ---- joint two-Lorentzian fit (pst_fit.cuh) on one contiguous spectrum, for the comparison with scipy's curve_fit
extern "C" void emul_fit_two_lorentzians(int n, const double* x, const double* y, const double* seed, double sigma0, double tol,
                                         int max_iter, double* out) {
  FitProblem pb;
  pb.y = y; pb.y_stride = 1; pb.x = x; pb.x_stride = 1; pb.n = n;
  fit_two_lorentzians(pb, seed, sigma0, tol, max_iter, out);
}
