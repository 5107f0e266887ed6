/* pistachio_b200 -- C ABI of the B200-native transfer-matrix (TMM) hot path.
 *
 * This is the drop-in boundary for the path BASELINE.json's north_star names.  The
 * reference (garrekstemo/pistachio, src/pistachio/transfer_matrix.py, "tm.py" below)
 * is pure Python and has no FFI; each entry point states the reference method(s) it
 * replaces.  INTEGRATION.md shows the ctypes stubs a maintainer of the reference
 * would add inside those methods.
 *
 * Conventions
 *  - every function returns 0 on success or a negative PST_E_* code; the message of
 *    the last failure on the calling thread is pst_last_error().  No C++ exceptions
 *    and no torch types cross this boundary: plain pointers and sizes only.
 *  - pointers marked [dev] are CUDA device pointers on the handle's device, [host]
 *    are ordinary host pointers.  The library never frees or retains caller buffers
 *    (small [host] metadata is copied during the call).
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Work
 *    is enqueued asynchronously; the caller synchronises.  A handle owns device
 *    scratch memory (optical tables, layer lists); calls that use it are ordered
 *    across streams by the library itself (each such call waits, on the device, for
 *    the previous one's work before overwriting the tables), so one handle may be
 *    used from several streams in turn.  A handle is NOT safe for concurrent calls
 *    from several host threads: use one handle per thread.
 *  - complex numbers are interleaved (re, im) doubles, numpy complex128 layout.
 *  - polarisation codes: PST_POL_S / PST_POL_P, the reference's "s-wave"/"p-wave".
 *  - angles are radians; wavelengths and thicknesses share one (arbitrary) unit.
 *  - result layout of grid evaluations: [sweep][pol][wavelength][theta], theta fastest.
 */
#ifndef PISTACHIO_B200_H
#define PISTACHIO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PST_ABI_VERSION 2

#define PST_OK 0
#define PST_E_INVALID (-1)  /* bad argument */
#define PST_E_CUDA (-2)     /* CUDA runtime error (message has the cudaError string) */
#define PST_E_NOMEM (-3)    /* host or device allocation failed */
#define PST_E_NODEVICE (-4) /* no usable sm_100 device */

#define PST_POL_S 1u /* "s-wave" */
#define PST_POL_P 2u /* "p-wave" */

/* flags of pst_grid_args */
#define PST_FLAG_FORCE_POINT_KERNEL (1u << 0) /* always one thread per point (K1) */
#define PST_FLAG_FORCE_DEEP_KERNEL (1u << 1)  /* always the layer-parallel scan kernel (K2) */
#define PST_FLAG_NUMERIC_DET (1u << 2)        /* T = Re(det(M) |t|^2) with det from the computed matrix, as
                                                 tm.py:486 does, instead of the closed form a_last/a_first */
#define PST_FLAG_NO_SMEM_STAGING (1u << 3)    /* read the optical table from global memory (debug/ablation) */
#define PST_FLAG_NO_SWEEP_REUSE (1u << 5)     /* thickness sweeps: rebuild the whole chain for every thickness instead
                                                 of reusing the prefix/suffix products (K3) -- ablation */
#define PST_FLAG_NO_PERIODIC_POWER (1u << 7)  /* apply a k-times repeated lossless layer pair k times instead of through
                                                 the closed-form power of its period (binary powering of U = x I + W in
                                                 the ring spanned by I and the traceless part W, W^2 = -s2 I: log2 k
                                                 steps, forward error O(k eps) like the product); both agree to
                                                 rounding -- ablation */
#define PST_FLAG_NO_BULK_COPY (1u << 6)       /* stage tables with plain 128-bit loads instead of cp.async.bulk (ablation) */
#define PST_FLAG_MULTICAST_OUT (1u << 8)       /* R, T, A are addresses inside an NVLink multicast mapping (CUDA multicast
                                                 object, e.g. torch symmetric memory's multicast_ptr, offset to this
                                                 rank's slab): results are written with multimem.st and land on every
                                                 GPU of the mapping -- a sharded sweep is assembled on all devices
                                                 without an all-gather.  The caller synchronises the devices afterwards.
                                                 Not valid with M or with pst_eval_grid_host */
#define PST_FLAG_FP32 (1u << 9)                /* optional fp32 mode: phase and its range reduction stay in fp64,
                                                 polynomials, layer matrices and the chain run in fp32 with the two
                                                 polarisations packed in f32x2 registers; inputs/outputs stay fp64.
                                                 Accuracy: |dR|, |dT| <= 1e-5 OF FULL SCALE (R, T in [0, 1]), i.e. 1e-5
                                                 relative only for values of order one; measured 3e-7 on few-layer
                                                 stacks, 9e-6 on a 2 x 20-pair DBR cavity (its finesse amplifies fp32
                                                 rounding).  Identical layers repeat their rounding error, so the error
                                                 of a periodic stack grows linearly with its depth: measured 1.1e-7 of
                                                 full scale per layer on low-contrast quarter-wave stacks (2.9e-5 at 256
                                                 layers, 2.7e-4 at 10 000).  Stacks of more than PST_FP32_MAX_LAYERS
                                                 interior layers -- where that growth reaches 1e-5 -- are refused with
                                                 PST_E_INVALID.  Point kernel without layer CSE; not valid with M */
#ifndef PST_FP32_MAX_LAYERS
#define PST_FP32_MAX_LAYERS 88
#endif
#define PST_FLAG_POL_DEGENERATE (1u << 10)     /* opt-in: compute the s-wave chain only and write it to both polarisation
                                                 planes.  In the reference's model (same theta in every layer, tm.py:209,
                                                 212,270) the p-wave matrices are a similarity transform of the s-wave
                                                 ones, M_p = L M_s L^-1 with L = diag(1, 1/cos^2 theta), and D_p = cos(theta)
                                                 L D_s, so R and T are the same function of the inputs for both
                                                 polarisations; the reference's two results differ only by their
                                                 rounding.  Needs pol_mask = S | P; not valid with M */
#define PST_FLAG_NO_SHAPE_FASTPATH (1u << 11)  /* run "mirror / spacer / mirror" programs through the general op interpreter
                                                 of the typed kernel instead of its straight-line form (ablation; results
                                                 are bit-identical) */
#define PST_FLAG_NO_LAYER_CSE (1u << 4)       /* rebuild the matrix of every layer even when layers repeat (ablation;
                                                 results are bit-identical with and without) */

#define PST_FLAG_SNELL (1u << 12)              /* refraction mode (additive; the reference has no counterpart -- it keeps
                                                 the incidence angle in every layer, tm.py:270): theta is the angle in the
                                                 incidence medium and layer j propagates at its refracted angle,
                                                 cos(theta_j) = sqrt(1 - (n_0 sin(theta_0) / n_j)^2), complex in absorbing
                                                 layers and beyond the critical angle; the root of the exit medium is the
                                                 one whose wave decays.  Same characteristic matrices, t, r, R and
                                                 T = Re(det M) |t|^2 (tm.py:236-277, 482-486); at normal incidence the
                                                 results equal the default mode's.  Built on the relations the reference
                                                 lists in fresnel()/transmission_matrix() (tm.py:730-751).  Excludes
                                                 PST_FLAG_FP32, PST_FLAG_POL_DEGENERATE and PST_FLAG_FORCE_DEEP_KERNEL. */

typedef struct pst_handle_s* pst_handle_t;

/* ---- lifetime ------------------------------------------------------------------------------------- */
int pst_abi_version(void);
const char* pst_last_error(void);
/* Bind a handle to CUDA device `device`.  Fails with PST_E_NODEVICE when no GPU is visible:
 * there is no CPU fallback anywhere in this library. */
int pst_create(int device, pst_handle_t* out);
int pst_destroy(pst_handle_t h);
/* Number of kernels this handle has launched so far (bench.py's gpu_launches). */
int64_t pst_launch_count(pst_handle_t h);

/* ---- K0: dispersion-table lookup ------------------------------------------------------------------
 * Replaces Layer.make_datapoints (tm.py:117-170): resample n_tab (lambda, n, k) tables onto the grid
 * wl[n_wl].  Table t occupies rows tab_off[t] .. tab_off[t+1]-1 of tab_wl/tab_n/tab_k and must be sorted
 * by wavelength (the host wrapper applies interp1d's stable sort).  A one-row table is broadcast
 * (tm.py:158-162); otherwise linear inter/extrapolation with scipy's two-term formula
 *   y = ((x-x_lo)/(x_hi-x_lo))*y_hi + ((x_hi-x)/(x_hi-x_lo))*y_lo,  i_hi = clip(lower_bound(x), 1, rows-1),
 * evaluated without FMA contraction so the result is bit-identical to the reference's.
 * out_n, out_k: [n_tab][n_wl]. */
int pst_resample_tables(pst_handle_t h, int n_tab, const int* tab_off /*[host] n_tab+1*/,
                        const double* tab_wl /*[dev]*/, const double* tab_n /*[dev]*/, const double* tab_k /*[dev]*/,
                        const double* wl /*[dev]*/, int n_wl, double* out_n /*[dev]*/, double* out_k /*[dev]*/,
                        void* stream);

/* ---- K1/K2/K3: fused matrix build + ordered chain + R/T/A over a (sweep, pol, lambda, theta) grid -----
 * Replaces, for every grid point at once, the reference call sequence
 *   Structure.calculate_layer_matrices(theta, pol)      tm.py:516-534  (-> Layer.calculate_transfer_matrix 236-277)
 *   Structure.calculate_all_transfer_matrices(...)      tm.py:536-572  (M = D_0^-1 (M_1..M_{N-2}) D_{N-1})
 *   Structure.calculate_all_t_r(M)                      tm.py:490-514  (t = 1/M00, r = M10/M00, R, T)
 * and the fan-out over angles Structure.angle_resolved_spectra (tm.py:597-628).
 * The reference model is kept: the same theta in every layer (no Snell refraction, tm.py:209,212,270). */
/* Optional fan-out of the results to several GPUs by the compute kernel itself (unicast stores over NVLink into
 * peer-mapped memory, e.g. the per-rank buffer_ptrs of a torch symmetric-memory allocation).  Entry i addresses the
 * first element of THIS call's output block inside GPU i's array; every result is stored once per entry (the caller's
 * own GPU is normally one of them).  See also PST_FLAG_MULTICAST_OUT for the NVSwitch-multicast variant. */
#define PST_MAX_PEERS 8
typedef struct pst_peer_outputs {
  int n_peers;
  double* R[PST_MAX_PEERS]; /* [dev, peer-mapped]; an array may be left out by setting all its entries to NULL */
  double* T[PST_MAX_PEERS];
  double* A[PST_MAX_PEERS];
} pst_peer_outputs;

typedef struct pst_grid_args {
  size_t struct_size; /* = sizeof(pst_grid_args) */
  int n_wl;           /* wavelengths */
  int n_theta;        /* incidence angles */
  int n_mat;          /* distinct (n,k) columns */
  int n_layers;       /* layers including the incidence (0) and exit (n_layers-1) media; >= 2 */
  int n_sweep;        /* thickness-sweep length; 1 when sweep_layer < 0 */
  int sweep_layer;    /* index of the interior layer whose thickness takes sweep_d[...]; -1 = none */
  unsigned pol_mask;  /* PST_POL_S | PST_POL_P; output pol axis has popcount(pol_mask) entries, s first */
  unsigned flags;
  const double* wl;        /* [dev] n_wl */
  const double* theta;     /* [dev] n_theta, radians */
  const double* cos_theta; /* [dev] n_theta or NULL.  When given it is used instead of cos(theta[i]) so a host
                              that already evaluated np.cos (as tm.py:209,270 do) gets the same bits.  Cosines of real
                              angles: |cos_theta[i]| <= 1 (the kernels bound the layer phases by |q d| on that basis
                              when they choose loops without large-argument handling; there is a factor 4 of slack) */
  const double* mat_n;     /* [dev] [n_mat][n_wl] real index on the grid */
  const double* mat_k;     /* [dev] [n_mat][n_wl] extinction coefficient (k > 0 absorbs) */
  const int* layer_mat;    /* [host] n_layers: column of mat_n/mat_k used by each layer */
  const double* layer_d;   /* [host] n_layers: thickness (entries 0 and n_layers-1 are ignored, tm.py:561-563) */
  const double* sweep_d;   /* [dev] n_sweep or NULL */
  double* R;               /* [dev] outputs [n_sweep][n_pol][n_wl][n_theta]; any may be NULL */
  double* T;
  double* A;               /* A = 1 - R - T (not computed by the reference; north_star) */
  double* M;               /* [dev] or NULL: total 2x2 matrices, complex128 [n_sweep][n_pol][n_wl][n_theta][2][2] */
  const pst_peer_outputs* peers; /* [host] or NULL.  When given, R, T, A above must be NULL: the results go to the
                                    listed peer buffers instead (pst_eval_grid only) */
} pst_grid_args;

int pst_eval_grid(pst_handle_t h, const pst_grid_args* args, void* stream);

/* Same computation with HOST buffers: wl, theta, cos_theta, mat_n, mat_k, sweep_d and the outputs are host
 * pointers (pinned memory recommended).  The wavelength axis is cut into chunks whose device->host copies overlap
 * the next chunk's kernels on internal streams.  Blocks until the results are in host memory.  This is the call
 * bench.py times as `e2e`. */
int pst_eval_grid_host(pst_handle_t h, const pst_grid_args* args_with_host_pointers);

/* ---- per-layer matrices (API parity) ------------------------------------------------------------------
 * Replaces Layer.calculate_all_transfer_matrices / calculate_transfer_matrix (tm.py:236-305) incl.
 * dynamical_matrix (190-215) and propagation_matrix (217-234): for each wavelength the four 2x2 matrices
 * [D, P, D^-1, M].  out: complex128 [n_wl][4][2][2].  cos_theta as in pst_grid_args (pass cos(theta)). */
int pst_layer_matrices(pst_handle_t h, const double* n_re /*[dev]*/, const double* n_im /*[dev]*/,
                       const double* wl /*[dev]*/, int n_wl, double theta, double cos_theta, unsigned pol,
                       double thickness, double* out /*[dev]*/, void* stream);

/* Replaces Layer.propagation_matrix (tm.py:217-234) for explicit (kx, d): out complex128 [n][2][2].
 * d may be complex (the reference's own test passes d = pi/(2 kx) with a complex kx); d_im = 0 for a thickness. */
int pst_propagation_matrices(pst_handle_t h, const double* kx /*[dev] complex n*/, int n, double d_re, double d_im,
                             double* out /*[dev]*/, void* stream);

/* Replaces Layer.calculate_wavenumbers (tm.py:172-188): kx = n w/c cos(theta), kz = n w/c sin(theta);
 * outputs complex128 [n_wl]. */
int pst_wavenumbers(pst_handle_t h, const double* n_re /*[dev]*/, const double* n_im /*[dev]*/,
                    const double* wl /*[dev]*/, int n_wl, double cos_theta, double sin_theta,
                    double* kx /*[dev]*/, double* kz /*[dev]*/, void* stream);

/* Replaces Structure.calculate_t_r / calculate_all_t_r (tm.py:464-514) for caller-supplied matrices:
 * M complex128 [n][2][2] -> T[n], R[n] with T = Re(det(M) |1/M00|^2), R = |M10/M00|^2. */
int pst_t_r_from_matrices(pst_handle_t h, const double* M /*[dev]*/, int64_t n, double* T /*[dev]*/,
                          double* R /*[dev]*/, void* stream);

/* Replaces the per-wavelength chain of Structure.calculate_all_transfer_matrices (tm.py:559-568) for explicit,
 * already-built layer matrices (the reference's cached Layer.transfer_matrices): for each wavelength i
 *   acc = I;  for j = n_interior-1 .. 0: acc = interior[j][i] . acc;   out[i] = first_inv[i] . (acc . last_d[i]).
 * first_inv, last_d, out: complex128 [n_wl][2][2]; interior: complex128 [n_interior][n_wl][2][2] (may be NULL
 * when n_interior == 0). */
int pst_chain_matrices(pst_handle_t h, const double* first_inv /*[dev]*/, const double* interior /*[dev]*/,
                       const double* last_d /*[dev]*/, int n_interior, int n_wl, double* out /*[dev]*/, void* stream);

/* Fresnel coefficients and interface matrices (tm.py:730-751; unused by the reference's own chain but part of
 * the path's vocabulary): for n pairs, out_rt complex128 [n][4] = (r_s, t_s, r_p, t_p); out_D (optional)
 * complex128 [n][2 (s,p)][2][2] = (1/t)[[1,r],[r,1]]. */
int pst_fresnel(pst_handle_t h, const double* n1 /*[dev] complex n*/, const double* n2, const double* k1x,
                const double* k2x, int64_t n, double* out_rt /*[dev]*/, double* out_D /*[dev] or NULL*/, void* stream);

/* Replaces transmission_matrix (tm.py:746-751): out complex128 [n][2][2] = (1/t[i]) [[1, r[i]], [r[i], 1]]. */
int pst_interface_matrices(pst_handle_t h, const double* r /*[dev] complex n*/, const double* t /*[dev] complex n*/,
                           int64_t n, double* out /*[dev]*/, void* stream);

/* Per-layer field amplitudes (additive; SURVEY.md section 8f-4 -- the reference's YAML files carry A0/B0 amplitude keys
 * that its code never reads, default/fabry-perot.yaml:13-14).  For every (polarisation, wavelength, theta) point of the
 * grid described by `a` (same meaning as pst_eval_grid; outputs R/T/A/M and peers are ignored, no thickness sweep, flags
 * = 0 or PST_FLAG_SNELL) and every layer j: the forward and backward amplitudes (A_j, B_j) = D_j^-1 (E, H)_tangential
 * at the LEFT boundary of layer j, with D_j the reference's dynamical matrix (tm.py:190-215), normalised to a unit
 * incident amplitude.  Layer 0 holds (1, r), the exit medium (t, 0), with r and t as in tm.py:482-483.
 * amp: device, complex128 [n_pol][n_wl][n_theta][n_layers][2]. */
int pst_field_amplitudes(pst_handle_t h, const pst_grid_args* a, double* amp /*[dev]*/, void* stream);

/* ---- fused spectral reduction --------------------------------------------------------
 * Peak search along the middle axis of y[n_outer][n_pts][n_inner] (a grid result: outer = (sweep, pol), pts =
 * wavelength, inner = theta) with the semantics of scipy.signal.find_peaks(y, height=, distance=) that the reference's
 * notebook applies to spectra (notebooks/Cavity Length Determiner.ipynb cell 3), so that only peak lists leave the
 * device instead of whole spectra.  use_height = 0 ignores `height`; distance <= 1 disables the distance condition.
 * reverse != 0 scans each spectrum from its last sample to its first (ascending wavenumber on an ascending
 * wavelength grid); indices then count from the end.  Outputs per spectrum s = outer * n_inner + inner:
 *   out_idx[s][max_peaks] (first out_count[s] entries valid), out_height[s][max_peaks] (may be NULL),
 *   out_count[s] (-1: more than max_peaks peaks -- candidates the distance condition removes do not count, however
 *   many), out_spacing[s] = mean difference of axis[] between consecutive peaks (NaN below two peaks; may be NULL; axis
 *   may be NULL) -- the notebook's `diff().mean()` from which the cavity length L = 1e4 / (2 n_eff spacing) follows
 *   (cells 5-6). */
int pst_find_peaks(pst_handle_t h, const double* y /*[dev]*/, int n_outer, int n_pts, int n_inner, double height,
                   int use_height, int distance, int max_peaks, int reverse, const double* axis /*[dev] n_pts or NULL*/,
                   int* out_idx /*[dev]*/, double* out_height /*[dev] or NULL*/, int* out_count /*[dev]*/,
                   double* out_spacing /*[dev] or NULL*/, void* stream);

/* Lower and upper polariton branch per spectrum, from the peak lists of pst_find_peaks (same y, reverse, axis,
 * max_peaks): the two highest kept peaks, each refined by the closed-form three-point Lorentzian fit (1/y is a
 * parabola around a Lorentzian peak).  Stands in, on the device, for the per-angle two-Lorentzian fits of the
 * reference's notebooks/Polariton Peak Analysis.ipynb (cells 8-17: l1_center = LP, l2_center = UP; the notebook's
 * lmfit dependency is not part of the reference repository, so there is no numerical oracle for it -- the restated
 * formula lives in oracle/peaks_oracle.py).  out[s][6] = (x0_lower, x0_upper, hwhm_lower, hwhm_upper, A_lower,
 * A_upper) in units of `axis`; NaN for spectra with fewer than two peaks. */
int pst_polariton_branches(pst_handle_t h, const double* y /*[dev]*/, int n_outer, int n_pts, int n_inner, int reverse,
                           const double* axis /*[dev] n_pts*/, const int* peak_idx /*[dev]*/,
                           const double* peak_height /*[dev]*/, const int* peak_count /*[dev]*/, int max_peaks,
                           double* out /*[dev] n_outer*n_inner*6*/, void* stream);

/* Joint two-Lorentzian least-squares fit per spectrum: the fit the reference's notebooks/Polariton Peak Analysis.ipynb
 * runs per angle (cells 8-14: lmfit `LorentzianModel(prefix='l1_') + LorentzianModel(prefix='l2_')`, six parameters,
 * Levenberg-Marquardt) and whose two centers are the lower and upper polariton (cell 17).  lmfit's line shape:
 *   f(x; amplitude, center, sigma) = (amplitude / pi) sigma / ((x - center)^2 + sigma^2).
 * y, n_outer, n_pts, n_inner, reverse, axis as in pst_find_peaks; seed[s][6] = the rows pst_polariton_branches wrote
 * (centers, half widths, peak heights of the two highest peaks; a NaN / non-positive half width is replaced by sigma0);
 * the fit uses samples [i_lo, i_hi) of the scan order (the notebook's wavenumber mask).  Levenberg-Marquardt with
 * Marquardt scaling, at most max_iter evaluations of the misfit, one pass over the window per evaluation; it stops at an
 * accepted step that lowers chi^2 by <= tol chi^2 or moves no parameter by more than tol of its size (MINPACK's ftol /
 * xtol; lmfit's default is 1.5e-8; clamped to >= 1e-15).  Steps that would take a centre more than one window span
 * outside the window, or a width beyond four spans, are refused.
 * out[s][8] = (center_lower, center_upper, sigma_lower, sigma_upper, amplitude_lower, amplitude_upper, chi^2, evaluations;
 * negative evaluations = tolerance not met within max_iter).  NaN rows for spectra whose seed is NaN (fewer than two
 * peaks).  Pinned against scipy.optimize.curve_fit (MINPACK Levenberg-Marquardt, what lmfit calls) in tests/test_peaks.py. */
int pst_fit_two_lorentzians(pst_handle_t h, const double* y /*[dev]*/, int n_outer, int n_pts, int n_inner, int reverse,
                            const double* axis /*[dev] n_pts*/, const double* seed /*[dev] n_outer*n_inner*6*/, int i_lo,
                            int i_hi, double sigma0, double tol, int max_iter, double* out /*[dev] n_outer*n_inner*8*/,
                            void* stream);

/* ---- measurement helpers ------------------------------------------------------------------------------ */
/* Independent-DFMA micro-benchmark: the measured FP64 (non-tensor) peak used as the roofline denominator.
 * Runs `iters` dependent FMAs on 8 independent chains per thread over a full-chip grid, returns TFLOP/s. */
int pst_fp64_peak_probe(pst_handle_t h, int iters, double* tflops_out, void* stream);
/* Overwrite a scratch buffer larger than L2 (flushes L2 between timed iterations). */
int pst_flush_l2(pst_handle_t h, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PISTACHIO_B200_H */
