// CUDA kernels of the transfer-matrix hot path (sm_100a, fp64, FP64-pipe bound -- no tensor cores:
// a chain of 2x2 complex products is not a dense contraction).
//
//   resample_kernel      K0  dispersion-table lookup            (reference tm.py:117-170)
//   prep_kernel              per (material, lambda): q = (n w)(1/c), n, 1/n (tm.py:269-270); per theta: cos, 1/cos;
//                            lossy flags of materials, layers and layer types -- one launch
//   tmm_chain_kernel     K1  one thread per (lambda, theta) point, both polarisations in the thread,
//                            optical table + layer list staged in shared memory once per block
//                        K2  same kernel with SEG=true: the interior layers are cut into blockDim.y
//                            ordered segments, one warp row per segment builds the 2x2 product of its
//                            segment, and the segment products are combined in order through shared
//                            memory (associative, non-commutative scan over layers)
//   tmm_typed_kernel     K1  layer CSE: identical (material, thickness) layers share one cached matrix; run-length op
//                            interpreter; repeated lossless pairs as the power of their period (pst_chain.cuh)
//   tmm_cavity_kernel    K1  periodic mirror / spacer / periodic mirror as straight-line code (the bench kernel)
//   tmm_sweep_kernel     K3  thickness sweeps: prefix / suffix of the chain once per thread, one sincos (+ exp pair) and
//                            two 2-term sums per thickness and polarisation, streaming stores
//   tmm_snell_kernel         refraction mode (PST_FLAG_SNELL);  field_amplitudes_kernel: (A_j, B_j) per layer
//   tmm_chain_f32_kernel     optional fp32 chain (pst_f32.cuh)
//   find_peaks_kernel, polariton_branches_kernel, fit_two_lorentzians_kernel (pst_fit.cuh)
//                            fused spectral reductions of the notebooks (scipy find_peaks semantics; joint two-Lorentzian
//                            Levenberg-Marquardt fit)
//   layer_matrices_kernel    [D, P, D^-1, M] per wavelength          (tm.py:190-305)
//   propagation_kernel, wavenumbers_kernel, chain_matrices_kernel, fresnel_kernel, interface_matrix_kernel
//                            the remaining per-method entry points of the drop-in API (tm.py:172-234, 559-568, 730-751)
//   t_r_kernel               (T, R) from caller-supplied matrices    (tm.py:464-514)
//
// Model: with phi = kx d, a = n cos(theta) (s) or n / cos(theta) (p),
//   M_j = D_j P_j D_j^-1 = [[cos phi, -i sin phi / a], [-i a sin phi, cos phi]],
//   M   = D_0^-1 (M_1 ... M_{N-2}) D_{N-1},   t = 1/M00, r = M10/M00, R = |r|^2, T = Re(det M |t|^2).
// Only the columns of M that are asked for are propagated: v <- M_j v from the exit side.
#pragma once
#include <type_traits>

#include "pst_bulk.cuh"
#include "pst_chain.cuh"
#include "pst_fit.cuh"

namespace pst {

struct LayerEntry {  // 16 bytes: one LDS.128 / LDG.128 per layer
  double d;
  int mat;
  int type;  // index into the per-thread layer-matrix cache (typed path); -1 when unused
};

// one 128-bit load per layer entry (the compiler otherwise fetches d, mat and type with three loads where they are used)
__device__ __forceinline__ LayerEntry load_entry(const LayerEntry* p) {
  const int4 raw = *reinterpret_cast<const int4*>(p);
  LayerEntry e;
  e.d = __hiloint2double(raw.y, raw.x);
  e.mat = raw.z;
  e.type = raw.w;
  return e;
}

constexpr int kMaxTypes = 8;        // distinct (material, thickness) pairs the typed path caches per thread
constexpr int kTypeRecBytes = 80;   // one cached layer matrix: 10 doubles (complex form) or 5 (real form); the
                                    // power-of-two exponents of all types follow the records as ints
constexpr int kRenormEvery = 16;
// deep-stack kernel: layers whose transcendental parts share a basic block, and groups per unrolled loop trip
// (measured on BASELINE configs[3] with the range test hoisted out of the loops -- the kernel is bound by dependent-issue
// latency at 7 warps per scheduler, so independent sincos chains per basic block pay: 2/1 0.550 ms, 2/2 0.579, 2/4 0.560,
// 4/1 0.539, 4/2 0.526, 4/4 0.516, 8/1 0.527, 8/2 0.551; 4/2 at 80 registers 0.521, at 64 registers 0.531.
// Before the hoist: 1/1 0.672, 2/2 0.569, 4/1 0.567)
#ifndef PST_K2_GROUP
#define PST_K2_GROUP 4
#endif
#ifndef PST_K2_UNROLL
#define PST_K2_UNROLL 4
#endif
constexpr int kK2Group = PST_K2_GROUP, kK2Unroll = PST_K2_UNROLL;
#ifndef PST_K2_MAXREG
#define PST_K2_MAXREG 72
#endif
#ifndef PST_K1_MAXREG
#define PST_K1_MAXREG 80
#endif
constexpr int kMaxOps = 24;         // run-length ops the typed kernel takes by value (longer programs use K1 untyped)

struct GridParams {
  int n_wl, n_theta, n_mat, n_layers;
  int sweep_layer;   // -1: none
  int pol_first;     // NPOL == 1: 0 = s, 1 = p
  unsigned pol_mask; // requested polarisations (bit 0 = s, bit 1 = p); used by the fp32 kernel, which always steps both
  int tl_rows;       // wavelength rows of the shared-memory optical table (>= max wavelengths per block)
  int opt_row_bytes; // bytes per wavelength row in shared memory: n_mat*48 (+16 pad when n_mat is even)
  int il0;           // first wavelength index of this launch (host pipeline launches wavelength windows)
  unsigned flags;
  int use_bulk;      // stage the tables with the bulk-copy engine (cp.async.bulk + mbarrier)
  double d_max;      // largest |thickness| of the layer list (host; the swept values are looked at by the kernels)
  long long points;  // wavelengths of this launch * n_theta
  const double* opt;        // [n_wl][n_mat][6]: OptRow order qr nr inr ni qi ini
  const int* mat_flags;     // [n_mat] bit0: k != 0 at some wavelength
  const unsigned* lossy_word;  // bit t: layer type t absorbs (typed kernel); bit 31: some material absorbs
  const double2* trig;      // [n_theta] (cos, 1/cos)
  const LayerEntry* layers; // [n_layers]
  const double* sweep_d;    // [gridDim.y] or nullptr
  double* R;
  double* T;
  double* A;
  double* M;
  int plain_out;            // 1: R, T and A all present, local, no peers, no multicast, no duplication: stores need no tests
  int pol_dup;              // 1: one polarisation is computed and written to both polarisation planes
                            // (PST_FLAG_POL_DEGENERATE: in the reference's model s and p are the same function)
  int n_peers;              // > 0: results go to these peer-mapped buffers instead of R, T, A (pst_peer_outputs)
  double* peer_R[8];
  double* peer_T[8];
  double* peer_A[8];
};

// Result stores.  PST_FLAG_MULTICAST_OUT (bit 8): R, T, A are addresses in an NVLink multicast mapping (CUDA multicast
// object; torch symmetric memory's multicast_ptr) and every value is written once with multimem.st -- the NVSwitch
// replicates it into the same offset of every GPU bound to the mapping, so a sharded sweep leaves the ASSEMBLED
// arrays on all devices without a separate all-gather pass (pistachio_b200/dist.py: AssembledResult).
constexpr unsigned kFlagMulticastOut = 1u << 8;
template <bool STREAMING>
__device__ __forceinline__ void store_result(double* p, double v, bool multicast) {
  if (multicast) {
    asm volatile("multimem.st.weak.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
  } else if (STREAMING) {
    __stcs(p, v);
  } else {
    *p = v;
  }
}

// One spectral point's R, T, A to wherever the call wants them: the local arrays (plain or multicast stores) or, with
// pst_peer_outputs, one unicast copy per listed GPU over NVLink (the pointer lists sit in the constant bank).
template <bool STREAMING>
__device__ __forceinline__ void store_point_once(const GridParams& prm, size_t o, double R, double T, double A);
template <bool STREAMING>
__device__ __forceinline__ void store_point(const GridParams& prm, size_t o, double R, double T, double A) {
  if (prm.plain_out) {  // the common call: one uniform test instead of seven
    store_result<STREAMING>(prm.R + o, R, false);
    store_result<STREAMING>(prm.T + o, T, false);
    store_result<STREAMING>(prm.A + o, A, false);
    return;
  }
  store_point_once<STREAMING>(prm, o, R, T, A);
  if (prm.pol_dup) store_point_once<STREAMING>(prm, o + (size_t)prm.points, R, T, A);  // the other polarisation's plane
}
template <bool STREAMING>
__device__ __forceinline__ void store_point_once(const GridParams& prm, size_t o, double R, double T, double A) {
  if (prm.n_peers > 0) {
    for (int i = 0; i < prm.n_peers; ++i) {
      if (prm.peer_R[i]) __stcs(prm.peer_R[i] + o, R);
      if (prm.peer_T[i]) __stcs(prm.peer_T[i] + o, T);
      if (prm.peer_A[i]) __stcs(prm.peer_A[i] + o, A);
    }
    return;
  }
  const bool mc = prm.flags & kFlagMulticastOut;
  if (prm.R) store_result<STREAMING>(prm.R + o, R, mc);
  if (prm.T) store_result<STREAMING>(prm.T + o, T, mc);
  if (prm.A) store_result<STREAMING>(prm.A + o, A, mc);
}

// The typed kernel's program, passed BY VALUE (constant bank): nothing on its dependent path is a global load.
struct TypedPlan {
  int n_types, n_ops;
  int sweep_type;           // type whose thickness is sweep_d[blockIdx.y]; -1: none
  int mat_first, mat_last;  // materials of the incidence and the exit medium
  int type_stride;          // bytes of layer-matrix cache per thread (odd multiple of 16: conflict-free LDS.128)
  int type_mat[kMaxTypes];
  double type_d[kMaxTypes];
  int4 ops[kMaxOps];        // (type A, type B or -1, repeat count, 0), in application order (exit side first)
  const unsigned* lossy_types;  // device word (prep_kernel): bits 0..7 the material of type t absorbs somewhere on the
                                // grid, bits 8..19 grid-wide growth bound E (pst_chain.cuh growth_log2), bit 31 any material
};

// The cavity kernel's program (by value): periodic mirror / spacer / periodic mirror over one pair of layer types --
// [power(A, B, k1), single C, power(A, B | B, A, k2)] in application order (exit side first); a DBR microcavity.
// Staged record per wavelength (doubles; pairs that are used together sit 16-byte aligned -> one 128-bit LDS):
//    0 q_A      1 q_B     |  2 n_A      3 1/n_A   |  4 n_B      5 1/n_B   |  6 q_C.re   7 d_C
//    8 n_C.re   9 1/n_C.re| 10 n_C.im  11 q_C.im  | 12 1/n_C.im 13 -      | 14 n_exit.re 15 n_exit.im
//   16 1/n_in.re 17 1/n_in.im | absorbing mirror layers only: 18..20 n_A.im q_A.im 1/n_A.im, 21..23 the same of B
constexpr int kCavityRow = 24;
struct CavityPlan {
  int src[kCavityRow];      // offsets (in doubles) of those values inside the wavelength's block of the optical table;
                            // -1: not from the table (d_C, padding)
  double d[3];              // thicknesses of A, B, C
  int k1, k2;               // pair counts of the first and the second mirror applied
  int hb1, hb2;             // floor(log2 k1), floor(log2 k2)
  int mirrored;             // 1: the second mirror applies (B, A) -- the mirror image of the first
  int c_swept;              // 1: the spacer is the swept layer (thickness from sweep_d)
  unsigned lossy_a, lossy_ab, lossy_c;  // bits of type A, of types A | B and of type C in the device's lossy-type word
  int tiles;                // theta tiles of 128 angles per block (row-mode launch)
  const unsigned* lossy_types;
};

// ---------------------------------------------------------------------------------------------- K0
// One thread per (table, wavelength).  lower_bound on the sorted table, then scipy's two-term formula
// with every operation rounded separately (no FMA) -- bit-identical to interp1d's linear branch.
__global__ void resample_kernel(int n_tab, const int* __restrict__ tab_off, const double* __restrict__ tab_wl,
                                const double* __restrict__ tab_n, const double* __restrict__ tab_k,
                                const double* __restrict__ wl, int n_wl, double* __restrict__ out_n,
                                double* __restrict__ out_k) {
  extern __shared__ double s_x[];  // the table's wavelength column when it fits (binary search in smem)
  const int t = blockIdx.y;
  const int off = tab_off[t];
  const int rows = tab_off[t + 1] - off;
  const bool staged = rows * sizeof(double) <= 32 * 1024;
  if (staged) {
    for (int i = threadIdx.x; i < rows; i += blockDim.x) s_x[i] = tab_wl[off + i];
    __syncthreads();
  }
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_wl) return;
  const double x = wl[i];
  if (rows == 1) {
    out_n[(size_t)t * n_wl + i] = tab_n[off];
    out_k[(size_t)t * n_wl + i] = tab_k[off];
    return;
  }
  const double* xs = staged ? s_x : tab_wl + off;
  int lo = 0, hi = rows;  // first index with xs[idx] >= x  (np.searchsorted side='left'; NaN sorts last)
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (xs[mid] < x) lo = mid + 1; else hi = mid;
  }
  if (x != x) lo = rows;
  int ih = min(max(lo, 1), rows - 1);
  int il = ih - 1;
  const double x_lo = xs[il], x_hi = xs[ih];
  const double den = sub_rn(x_hi, x_lo);
  const double w_hi = div_rn(sub_rn(x, x_lo), den);
  const double w_lo = div_rn(sub_rn(x_hi, x), den);
  out_n[(size_t)t * n_wl + i] = add_rn(mul_rn(w_hi, tab_n[off + ih]), mul_rn(w_lo, tab_n[off + il]));
  out_k[(size_t)t * n_wl + i] = add_rn(mul_rn(w_hi, tab_k[off + ih]), mul_rn(w_lo, tab_k[off + il]));
}

// ------------------------------------------------------------------------------------------- prep
// One launch prepares everything the chain kernels read (it used to be a memset and three launches, 12 us of a 275 us
// bench step):
//   blocks [0, n_opt_blocks)   optical table row entries q, n, 1/n per (material, wavelength) and the materials'
//                              lossy flags (k != 0 somewhere on the grid; warp-aggregated atomicOr)
//   the remaining blocks       (cos, 1/cos) per angle
//   the block that finishes last (ticket counter) derives what depends on the complete flags: the layer list with the
//   material's lossy flag in its `type` field (untyped kernels need no second lookup and can stage the list with a
//   plain bulk copy), the lossy bits of the typed kernel's layer types, and then clears flags and ticket for the next
//   call.
struct TypeMats {
  int n;
  int mat[kMaxTypes];
};
constexpr int kTrigPad = 1024;  // >= 128 * (largest cp.tiles)

__global__ void __launch_bounds__(256)
prep_kernel(const double* __restrict__ wl, const double* __restrict__ mat_n, const double* __restrict__ mat_k, int n_wl,
            int n_mat, double* __restrict__ opt, int* __restrict__ mat_flags, const double* __restrict__ theta,
            const double* __restrict__ cos_theta, int n_theta, double2* __restrict__ trig, int n_opt_blocks,
            const LayerEntry* __restrict__ layers, int n_layers, LayerEntry* __restrict__ layers_out, const TypeMats tm,
            unsigned* __restrict__ lossy_types, unsigned* __restrict__ ticket) {
  // ticket[0]: arrival counter; ticket[1], ticket[2]: largest high word (sign cleared) of the |n|, |1/n| components and
  // of 1/cos(theta) over the grid -- the grid-wide growth bound of pst_chain.cuh (growth_log2); zero between calls
  if ((int)blockIdx.x < n_opt_blocks) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < (long long)n_wl * n_mat) {
      const int m = (int)(idx / n_wl);
      const int i = (int)(idx - (long long)m * n_wl);
      const double nk = mat_k[idx];
      const OptRow r = make_opt_row(wl[i], mat_n[idx], nk);
      double2* o = reinterpret_cast<double2*>(opt + ((size_t)i * n_mat + m) * 6);
      o[0] = make_double2(r.qr, r.nr);
      o[1] = make_double2(r.inr, r.ni);
      o[2] = make_double2(r.qi, r.ini);
      const unsigned act = __activemask();
      {
        unsigned hw = (unsigned)hi_word(r.nr) & 0x7fffffffu;
        hw = max(hw, (unsigned)hi_word(r.ni) & 0x7fffffffu);
        hw = max(hw, (unsigned)hi_word(r.inr) & 0x7fffffffu);
        hw = max(hw, (unsigned)hi_word(r.ini) & 0x7fffffffu);
        hw = __reduce_max_sync(act, hw);
        if ((threadIdx.x & 31) == __ffs(act) - 1 && hw > __ldcg(ticket + 1)) atomicMax(ticket + 1, hw);
      }
      // the material absorbs (or amplifies) somewhere on the grid: one atomic per warp and material at most
      const int m0 = __shfl_sync(act, m, __ffs(act) - 1);
      const bool any_here = __any_sync(act, nk != 0.0 && m == m0);
      if (m == m0) {
        if (any_here && (threadIdx.x & 31) == __ffs(act) - 1) atomicOr(mat_flags + m, 1);
      } else if (nk != 0.0) {
        atomicOr(mat_flags + m, 1);  // warp straddling two materials
      }
    }
  } else {
    // the angle table is padded by kTrigPad copies of the last angle: the cavity kernel's tile loop reads whole tiles
    const int ip = ((int)blockIdx.x - n_opt_blocks) * blockDim.x + threadIdx.x;
    if (ip < n_theta + kTrigPad) {
      const int i = min(ip, n_theta - 1);
      const double c = cos_theta ? cos_theta[i] : cos(theta[i]);
      const double ic = 1.0 / c;
      trig[ip] = make_double2(c, ic);
      const unsigned act = __activemask();
      const unsigned hw = __reduce_max_sync(act, (unsigned)hi_word(ic) & 0x7fffffffu);
      if ((threadIdx.x & 31) == __ffs(act) - 1 && hw > __ldcg(ticket + 2)) atomicMax(ticket + 2, hw);
    }
  }
  // ---- the last block to arrive finishes the flag-dependent part
  __shared__ int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1u) ? 1 : 0;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (threadIdx.x == 0) {  // bit t: the material of layer type t absorbs (typed kernel); bit 31: any material does
    unsigned m = 0;
    for (int t = 0; t < tm.n; ++t) m |= (__ldcg(mat_flags + tm.mat[t]) != 0 ? 1u : 0u) << t;
    for (int i = 0; i < n_mat; ++i) m |= (__ldcg(mat_flags + i) != 0 ? 1u : 0u) << 31;
    const int e_n = (int)(__ldcg(ticket + 1) >> 20) - 1023, e_g = (int)(__ldcg(ticket + 2) >> 20) - 1023;
    const int E = min(max(e_n + e_g + 2, 0), 4095);
    *lossy_types = m | ((unsigned)E << 8);
  }
  for (int j = threadIdx.x; j < n_layers; j += blockDim.x) {
    LayerEntry e = layers[j];
    e.type = __ldcg(mat_flags + e.mat);
    layers_out[j] = e;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n_mat; i += blockDim.x) mat_flags[i] = 0;
  if (threadIdx.x == 0) { ticket[0] = 0u; ticket[1] = 0u; ticket[2] = 0u; }
}

// ------------------------------------------------------------------------------------ K1 / K2
// One launch covers `points` = (wavelengths of the launch) x n_theta spectral points (theta fastest) for gridDim.y
// thickness-sweep values; each thread owns one point and both polarisations.
//   STAGE_OPT / STAGE_LAYERS  optical table window and layer list staged in shared memory once per block
//   SEG    blockDim.y ordered layer segments per point, combined in order through shared memory (deep stacks)
template <int NPOL, int NCOL, bool STAGE_OPT, bool STAGE_LAYERS, bool SEG>
__global__ void __launch_bounds__(SEG ? 512 : 256)
__maxnreg__(SEG ? (NPOL == 1 ? PST_K2_MAXREG : 128) : (NCOL == 1 ? PST_K1_MAXREG : 255))  // deep-stack kernel, one polarisation: 72 registers = 28 warps per SM, what
tmm_chain_kernel(const GridParams prm) {          // BASELINE configs[3] (16 384 points x 8 segments) needs to be resident at once
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int NC = SEG ? 2 : NCOL;  // a segment product needs both columns
  const int PB = blockDim.x;          // points per block
  const int W = blockDim.y;           // ordered layer segments per point (1 unless SEG)
  const int tid = threadIdx.y * PB + threadIdx.x;
  const int nthr = PB * W;
  // a launch covers at most 2^31 - 1 points (host-enforced): 32-bit index arithmetic, no 64-bit divisions
  const unsigned npts = (unsigned)prm.points, nth = (unsigned)prm.n_theta;
  const unsigned p0 = blockIdx.x * (unsigned)PB;
  const unsigned p_last = min(p0 + (unsigned)PB, npts) - 1u;
  const int l_lo = prm.il0 + (int)(p0 / nth);
  const int TL = prm.il0 + (int)(p_last / nth) - l_lo + 1;
  const int L = prm.n_layers;

  // ---- shared-memory carve-up: [optical table][layer list][segment products]
  unsigned char* s_opt = smem_raw;
  size_t off = STAGE_OPT ? (size_t)prm.tl_rows * prm.opt_row_bytes : 0;
  LayerEntry* s_layers = reinterpret_cast<LayerEntry*>(smem_raw + off);
  off += STAGE_LAYERS ? (size_t)L * sizeof(LayerEntry) : 0;
  double* s_seg = reinterpret_cast<double*>(smem_raw + off);
  // SEG without a staged layer list (deep stacks: the list does not fit): every warp streams its segments' slices of the
  // list through a double-buffered shared-memory window with the bulk-copy engine (below)
  constexpr bool STREAM = SEG && !STAGE_LAYERS;
  constexpr int kStreamChunk = 32;  // layers per chunk and segment (512 bytes)
  off += SEG ? (size_t)W * NPOL * 9 * PB * sizeof(double) : 0;
  LayerEntry* s_stream = reinterpret_cast<LayerEntry*>(smem_raw + off);  // [warp][2][segments per warp][kStreamChunk]
  __shared__ __align__(8) uint64_t s_sbar[16][2];                        // one mbarrier per warp and buffer

  // the block's wavelength window of the [wl][mat][6] table is one contiguous global range; rows are re-pitched
  // to opt_row_bytes so that lanes on different wavelengths hit different 16-byte bank groups (pst_bulk.cuh)
  if constexpr (STAGE_OPT || STAGE_LAYERS)
    stage_block(prm.use_bulk != 0, s_opt,
                reinterpret_cast<const unsigned char*>(prm.opt + (size_t)l_lo * prm.n_mat * 6), STAGE_OPT ? TL : 0,
                prm.n_mat * 48, prm.opt_row_bytes, reinterpret_cast<unsigned char*>(s_layers),
                reinterpret_cast<const unsigned char*>(prm.layers), STAGE_LAYERS ? L * (int)sizeof(LayerEntry) : 0, tid,
                nthr);

  const unsigned p = min(p0 + threadIdx.x, npts - 1u);
  const bool valid = (p0 + threadIdx.x) < npts;
  const int il_loc = (int)(p / nth);
  const int it = (int)(p - (unsigned)il_loc * nth);
  const int il = prm.il0 + il_loc;
  const double2 tr = __ldg(prm.trig + it);
  const double ct = tr.x, ict = tr.y;
  const int sweep = blockIdx.y;
  const double d_sweep = prm.sweep_d ? __ldg(prm.sweep_d + sweep) : 0.0;
  // a staged layer list takes the block's swept thickness in place: the layer loops then test nothing per layer
  if constexpr (STAGE_LAYERS) {
    if (prm.sweep_d) {
      if (tid == 0) s_layers[prm.sweep_layer].d = d_sweep;
      __syncthreads();
    }
  }

  // optical-table accessor for this thread's wavelength: 48-byte rows, three 16-byte loads
  const unsigned char* optp = STAGE_OPT ? (s_opt + (size_t)(il - l_lo) * prm.opt_row_bytes)
                                        : reinterpret_cast<const unsigned char*>(prm.opt + (size_t)il * prm.n_mat * 6);
  const LayerEntry* lay = STAGE_LAYERS ? s_layers : prm.layers;
  auto opt_row = [&](int mat) {
    const double2* o = reinterpret_cast<const double2*>(optp + (size_t)mat * 48);
    const double2 a = o[0], b = o[1], c = o[2];
    OptRow r;
    r.qr = a.x; r.nr = a.y; r.inr = b.x; r.ni = b.y; r.qi = c.x; r.ini = c.y;
    return r;
  };
  auto opt_row_real = [&](int mat, double& qr, double& nr, double& inr) {
    const double2* o = reinterpret_cast<const double2*>(optp + (size_t)mat * 48);
    const double2 a = o[0];
    qr = a.x; nr = a.y; inr = o[1].x;
  };
  const OptRow exit_row = opt_row(lay[L - 1].mat);

  // Every phase of this thread is |q cos(theta) d| <= max|q| |cos| max|d|.  When that bound is inside sincos' fast range
  // for the whole warp, the layer loops below run without the per-layer range test (an FP64 compare, a branch and a
  // convergence barrier per layer); otherwise the tested loops run.  Same functions on the same arguments either way.
  bool warp_fast;
  {
    double qmax = 0.0;
    for (int m = 0; m < prm.n_mat; ++m) {
      const double2 a = reinterpret_cast<const double2*>(optp + (size_t)m * 48)[0];
      qmax = fmax(qmax, fabs(a.x));
    }
    const double bound = qmax * fabs(ct) * fmax(prm.d_max, fabs(d_sweep));
    warp_fast = __all_sync(0xffffffffu, bound < 536870912.0) != 0;  // 2^29: a factor two of slack for the roundings
  }

  // ---- the ordered chain over this thread's segment of interior layers
  const int Li = L - 2;
  const int seg = SEG ? threadIdx.y : 0;
  // balanced segment bounds without 64-bit division: the first Li % W segments get one extra layer
  const int seg_q = Li / W, seg_r = Li - seg_q * W;
  const int j_lo = 1 + seg * seg_q + min(seg, seg_r);
  const int j_hi = j_lo + seg_q + (seg < seg_r ? 1 : 0);  // exclusive

  ChainState<NPOL, NC> st;
  if constexpr (SEG) init_identity<NPOL>(st);
  else init_exit_columns<NPOL, NC>(st, exit_row.nr, exit_row.ni, ct, prm.pol_first);

  // Rescaling interval from the grid-wide growth bound (prep_kernel; the typed kernel's rule): one layer multiplies
  // max|v| by at most 2^lg, so 900 / lg layers after a rescale cannot leave the fp64 range -- ordinary stacks of a
  // hundred layers rescale once, at the end, instead of every kRenormEvery layers (rescaling is an exact power of two:
  // no result bit depends on where it happens).  Block-uniform: it sets trip counts.
  const unsigned growth_word = __ldg(prm.lossy_word);
  const int lg_layer = growth_log2(growth_E_of_word(growth_word), (growth_word >> 31) != 0u);
  const int every = max(kRenormEvery, min(1024, __float2int_rz(899.5f * __frcp_rn((float)lg_layer)))) & ~(kRenormEvery - 1);

  // ---- iteration over this thread's layers j_hi-1 .. j_lo: apply(entry, j) per layer, rescale() every `every` layers
  auto for_each_layer = [&](auto UNROLLED, auto&& apply, auto&& apply2, auto&& rescale) {
    if constexpr (!STREAM) {
      int j = j_hi - 1;
      while (j >= j_lo) {
        const int j_stop = max(j - every, j_lo - 1);
        for (; j - 1 > j_stop; j -= 2) apply2(lay + j, j, std::integral_constant<int, 2>{});  // layers j, j - 1
        if (j > j_stop) {
          apply(lay[j], j);
          --j;
        }
        rescale();
      }
    } else {
      // A warp holds SW = 32 / PB consecutive segments (lane = segment-in-warp * PB + point).  Chunk k of a segment is the
      // kStreamChunk layers below j_hi - k kStreamChunk: one contiguous, 16-byte aligned range of the global list -> one
      // cp.async.bulk per segment and chunk, completing on the warp's mbarrier of that buffer; chunk k + 1 is in flight
      // while chunk k is multiplied.  The per-layer global load (and its L2 round trip on the dependent path: 1.4
      // long-scoreboard stalls per issue in round 1's capture) becomes one broadcast LDS.128, and full batches of
      // kRenormEvery layers run unrolled: no loop counter, no index arithmetic per layer.
      const int lane = tid & 31, warp = tid >> 5;
      const int SWs = 32 / PB, segw = lane / PB;
      LayerEntry* wbuf = s_stream + (size_t)warp * 2 * SWs * kStreamChunk;
      if (lane == 0) {
        mbar_init(&s_sbar[warp][0], 1);
        mbar_init(&s_sbar[warp][1], 1);
      }
      __syncwarp();
      const int n_mine = j_hi - j_lo;
      const int n_chunks = (seg_q + (seg_r ? 1 : 0) + kStreamChunk - 1) / kStreamChunk;  // block-uniform upper bound
      const bool leader = (lane - segw * PB) == 0;
      auto issue = [&](int k) {
        // all segments of the warp copy [lo, hi) with lo = max(j_hi - (k + 1) CH, 0): the bytes are summed over the
        // warp's segment leaders (lanes with point index 0) for the transaction count
        const int hi = j_hi - k * kStreamChunk, lo = max(hi - kStreamChunk, 0);
        const int bytes = (leader && hi > lo) ? (hi - lo) * (int)sizeof(LayerEntry) : 0;
        const int total = __reduce_add_sync(0xffffffffu, bytes);
        if (lane == 0) mbar_expect_tx(&s_sbar[warp][k & 1], (uint32_t)total);
        __syncwarp();
        // the copy is ascending in j and lands at slots [CH - (hi - lo), CH): layer hi - 1 - i sits at slot CH - 1 - i
        if (bytes > 0)
          bulk_g2s(wbuf + ((size_t)(k & 1) * SWs + segw) * kStreamChunk + (kStreamChunk - (hi - lo)), prm.layers + lo,
                   (uint32_t)bytes, &s_sbar[warp][k & 1]);
      };
      issue(0);
      int since = 0;  // layers (rounded up to batches) applied since the last rescale
      for (int k = 0; k < n_chunks; ++k) {
        if (k + 1 < n_chunks) issue(k + 1);
        mbar_wait(&s_sbar[warp][k & 1], (uint32_t)((k >> 1) & 1));
        const LayerEntry* top = wbuf + ((size_t)(k & 1) * SWs + segw) * kStreamChunk + (kStreamChunk - 1);  // slot of layer j_top
        const int j_top = j_hi - 1 - k * kStreamChunk;
        const int left = n_mine - k * kStreamChunk;  // layers of this lane not yet applied (may be <= 0)
#pragma unroll
        for (int b0 = 0; b0 < kStreamChunk; b0 += kRenormEvery) {
          if (decltype(UNROLLED)::value && left >= b0 + kRenormEvery) {
#pragma unroll kK2Unroll
            for (int i = 0; i < kRenormEvery; i += kK2Group)  // kK2Group layers at a time: their sincos chains interleave
              apply2(top - (b0 + i), j_top - (b0 + i), std::integral_constant<int, kK2Group>{});
          } else if (left > b0) {
            const int end = min(left, b0 + kRenormEvery);
            for (int i = b0; i < end; ++i) apply(top[-i], j_top - i);
          }
          since += kRenormEvery;
          if (since >= every) { rescale(); since = 0; }  // block-uniform
        }
        __syncwarp();  // every lane is done with this buffer before chunk k + 2 overwrites it
      }
      rescale();  // the segment product leaves normalised: the ordered combine multiplies products
    }
  };

  bool done = false;
  if constexpr (SEG) {
    if (!(__ldg(prm.lossy_word) >> 31)) {
      // Every layer lossless: M_j = [[c, -i be], [-i al, c]] with real c, al, be, and a product of such matrices keeps
      // the pattern [[real, i real], [i real, real]] -- the segment product is four real numbers per polarisation
      // (column 0 = (a, i w), column 1 = (i u, b)), half the arithmetic of the general complex columns, bit for bit
      // the same values (the general path multiplies the structural zeros).
      double a[NPOL], w[NPOL], u[NPOL], b[NPOL];
      int ke[NPOL];
#pragma unroll
      for (int pp = 0; pp < NPOL; ++pp) { a[pp] = 1.0; w[pp] = 0.0; u[pp] = 0.0; b[pp] = 1.0; ke[pp] = 0; }
      double g[NPOL], gi[NPOL];
#pragma unroll
      for (int pp = 0; pp < NPOL; ++pp) {
        const bool is_s = pol_is_s<NPOL>(pp, prm.pol_first);
        g[pp] = is_s ? ct : ict;
        gi[pp] = is_s ? ict : ct;
      }
      const int sweep_j = prm.sweep_d ? prm.sweep_layer : -1;
      auto phase_of = [&](const LayerEntry& le, int j, double& nr, double& inr) {
        const double d = (!STAGE_LAYERS && j == sweep_j) ? d_sweep : le.d;
        double qr;
        opt_row_real(le.mat, qr, nr, inr);
        return layer_phase(qr, ct, d);
      };
      auto step = [&](double sx, double cx, double nr, double inr) {
        const RealLayer<NPOL> m = finish_real_layer_g<NPOL>(sx, cx, nr, inr, g, gi);
#pragma unroll
        for (int pp = 0; pp < NPOL; ++pp) {
          const double a0 = a[pp], w0 = w[pp], u0 = u[pp], b0 = b[pp];
          a[pp] = fma(m.c, a0, m.be[pp] * w0);
          w[pp] = fma(m.c, w0, -(m.al[pp] * a0));
          u[pp] = fma(m.c, u0, -(m.be[pp] * b0));
          b[pp] = fma(m.c, b0, m.al[pp] * u0);
        }
      };
      auto run = [&](auto NOTEST) {
      for_each_layer(
          std::true_type{},
          [&](const LayerEntry& le, int j) {
            double nr, inr, sx, cx;
            if constexpr (decltype(NOTEST)::value) sincos_core(phase_of(le, j, nr, inr), sx, cx);
            else sincos_fast(phase_of(le, j, nr, inr), sx, cx);
            step(sx, cx, nr, inr);
          },
          [&](const LayerEntry* first, int j_first, auto N_) {
            // the transcendental parts of N consecutive layers (first[0], first[-1], ...) in one basic block with one joint
            // range test (none when the warp's phase bound already guarantees the range), then the N ordered steps
            constexpr int N = decltype(N_)::value;
            double nr[N], inr[N], x[N], sx[N], cx[N];
            bool fast = true;
#pragma unroll
            for (int i = 0; i < N; ++i) {
              x[i] = phase_of(first[-i], j_first - i, nr[i], inr[i]);
              if constexpr (!decltype(NOTEST)::value) fast = fast && sincos_in_fast_range(x[i]);
            }
            if (fast) {
#pragma unroll
              for (int i = 0; i < N; ++i) sincos_core(x[i], sx[i], cx[i]);
            } else {
#pragma unroll
              for (int i = 0; i < N; ++i) sincos_fast(x[i], sx[i], cx[i]);
            }
#pragma unroll
            for (int i = 0; i < N; ++i) step(sx[i], cx[i], nr[i], inr[i]);
          },
          [&]() {
#pragma unroll
            for (int pp = 0; pp < NPOL; ++pp) {  // exact power-of-two rescale, as renormalise()
              unsigned mh = (unsigned)hi_word(a[pp]) & 0x7fffffffu;
              mh = max(mh, (unsigned)hi_word(w[pp]) & 0x7fffffffu);
              mh = max(mh, (unsigned)hi_word(u[pp]) & 0x7fffffffu);
              mh = max(mh, (unsigned)hi_word(b[pp]) & 0x7fffffffu);
              const int e = clamp_exp((int)(mh >> 20) - 1023);
              const double sc = from_words((1023 - e) << 20, 0);
              a[pp] *= sc; w[pp] *= sc; u[pp] *= sc; b[pp] *= sc;
              ke[pp] += e;
            }
          });
      };
      if (warp_fast) run(std::true_type{});
      else run(std::false_type{});
#pragma unroll
      for (int pp = 0; pp < NPOL; ++pp) {
        st.v[pp][0][0] = a[pp]; st.v[pp][0][1] = 0.0; st.v[pp][0][2] = 0.0; st.v[pp][0][3] = w[pp];
        st.v[pp][1][0] = 0.0; st.v[pp][1][1] = u[pp]; st.v[pp][1][2] = b[pp]; st.v[pp][1][3] = 0.0;
        st.kexp[pp] = ke[pp];
      }
      done = true;
    }
  }
  if (!done) {
    auto run = [&](auto NOTEST) {
    constexpr bool IN_RANGE = decltype(NOTEST)::value;  // the warp's phase bound covers every layer: no call in the loop
    auto apply_any = [&](const LayerEntry& le, int j) {
          const double d = (!STAGE_LAYERS && j == prm.sweep_layer) ? d_sweep : le.d;
          const int lossy = le.type;  // untyped launches get the flagged layer list (prep_kernel)
          if (lossy == 0) {
            double qr, nr, inr;
            opt_row_real(le.mat, qr, nr, inr);
            apply_real_layer<NPOL, NC>(st, build_real_layer<NPOL, IN_RANGE>(qr, nr, inr, d, ct, ict, prm.pol_first));
          } else {
            apply_cplx_layer<NPOL, NC>(st, build_cplx_layer<NPOL, IN_RANGE>(opt_row(le.mat), d, ct, ict, prm.pol_first));
          }
        };
    for_each_layer(std::false_type{}, apply_any,
                   [&](const LayerEntry* first, int j_first, auto N_) {
                     constexpr int N = decltype(N_)::value;
                     if constexpr (N == 2) {
                       // two lossless neighbours (block-uniform: the flag is the material's): their sincos chains share a
                       // basic block under one joint range test, then the two ordered steps
                       const LayerEntry le0 = load_entry(first), le1 = load_entry(first - 1);
                       if (le0.type == 0 && le1.type == 0) {
                         double q0, n0, in0, q1, n1, in1, sx0, cx0, sx1, cx1;
                         opt_row_real(le0.mat, q0, n0, in0);
                         opt_row_real(le1.mat, q1, n1, in1);
                         const double x0 = layer_phase(q0, ct, (!STAGE_LAYERS && j_first == prm.sweep_layer) ? d_sweep : le0.d);
                         const double x1 = layer_phase(q1, ct, (!STAGE_LAYERS && j_first - 1 == prm.sweep_layer) ? d_sweep : le1.d);
                         if (decltype(NOTEST)::value || (sincos_in_fast_range(x0) && sincos_in_fast_range(x1))) {
                           sincos_core(x0, sx0, cx0);
                           sincos_core(x1, sx1, cx1);
                         } else {
                           sincos_fast(x0, sx0, cx0);
                           sincos_fast(x1, sx1, cx1);
                         }
                         apply_real_layer<NPOL, NC>(st, finish_real_layer<NPOL>(sx0, cx0, n0, in0, ct, ict, prm.pol_first));
                         apply_real_layer<NPOL, NC>(st, finish_real_layer<NPOL>(sx1, cx1, n1, in1, ct, ict, prm.pol_first));
                         return;
                       }
                     }
                     for (int i = 0; i < N; ++i) apply_any(first[-i], j_first - i);
                   },
        [&]() { renormalise<NPOL, NC>(st); });
    };
    if (warp_fast) run(std::true_type{});
    else run(std::false_type{});
  }

  ChainState<NPOL, NCOL> fin;
  if constexpr (SEG) {
    // ---- ordered combine of the segment products S_0 S_1 ... S_{W-1} (associative, not commutative):
    // level 1, warp shuffles: a warp holds SW = 32 / PB consecutive segments of the same PB points (lane =
    // segment-in-warp * PB + point); a log2(SW)-step ordered tree multiplies neighbours, P_seg <- P_seg . P_{seg+s}
    const int SW = 32 / PB;
    for (int s = 1; s < SW; s <<= 1) {
      ChainState<NPOL, 2> other;
#pragma unroll
      for (int pp = 0; pp < NPOL; ++pp) {
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int k = 0; k < 4; ++k) other.v[pp][c][k] = __shfl_xor_sync(0xffffffffu, st.v[pp][c][k], s * PB);
        other.kexp[pp] = __shfl_xor_sync(0xffffffffu, st.kexp[pp], s * PB);
      }
      if (((seg % SW) & (2 * s - 1)) == 0) {  // this lane holds the left factor: apply it to the partner's columns
#pragma unroll
        for (int pp = 0; pp < NPOL; ++pp) {
          double left[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) left[k] = st.v[pp][k >> 2][k & 3];
          apply_segment<NPOL, 2>(other, pp, left, st.kexp[pp]);
        }
        st = other;
        renormalise<NPOL, 2>(st);
      }
    }
    // level 2, block: the first lane group of every warp publishes its product in shared memory and the seg-0 row
    // applies them, last to first, to the exit columns
    constexpr int REC = 9;  // 8 doubles + exponent per (warp product, polarisation); lanes contiguous
    const int WR = W / SW;  // products left after the shuffle level
    if (seg % SW == 0) {
      double* mine = s_seg + ((size_t)((seg / SW) * NPOL) * REC) * PB + threadIdx.x;
#pragma unroll
      for (int pp = 0; pp < NPOL; ++pp) {
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int k = 0; k < 4; ++k) mine[((size_t)pp * REC + c * 4 + k) * PB] = st.v[pp][c][k];
        mine[((size_t)pp * REC + 8) * PB] = (double)st.kexp[pp];
      }
    }
    __syncthreads();
    if (threadIdx.y != 0) return;
    init_exit_columns<NPOL, NCOL>(fin, exit_row.nr, exit_row.ni, ct, prm.pol_first);
    for (int w = WR - 1; w >= 0; --w) {
      const double* rec = s_seg + ((size_t)(w * NPOL) * REC) * PB + threadIdx.x;
#pragma unroll
      for (int pp = 0; pp < NPOL; ++pp) {
        double s[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) s[k] = rec[((size_t)pp * REC + k) * PB];
        apply_segment<NPOL, NCOL>(fin, pp, s, (int)rec[((size_t)pp * REC + 8) * PB]);
      }
      renormalise<NPOL, NCOL>(fin);
    }
  } else {
    fin = st;
  }
  if (!valid) return;

  // ---- epilogue
  const OptRow in_row = opt_row(lay[0].mat);
  const size_t plane = (size_t)prm.points;
#pragma unroll
  for (int pp = 0; pp < NPOL; ++pp) {
    const PointOut r = finish_point<NPOL, NCOL>(fin, pp, in_row.inr, in_row.ini, exit_row.nr, exit_row.ni, ict,
                                                prm.pol_first, prm.flags);
    const size_t o = ((size_t)sweep * (NPOL << prm.pol_dup) + pp) * plane + (size_t)p;
    store_point<false>(prm, o, r.R, r.T, r.A);
    if constexpr (NCOL == 2) {
      if (prm.M) {
        double4* mo = reinterpret_cast<double4*>(prm.M + o * 8);
        mo[0] = make_double4(r.M[0], r.M[1], r.M[2], r.M[3]);
        mo[1] = make_double4(r.M[4], r.M[5], r.M[6], r.M[7]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------ K1 typed (layer CSE)
// Layers with identical (material, thickness) share one layer matrix -- pure common-subexpression elimination: the
// applied matrices are bit-identical to the untyped path's.  Every distinct matrix is built once per thread into a
// thread-private shared-memory record; the host compiles the layer sequence into run-length ops "apply (A, B) k
// times" (periodic stacks: quarter-wave mirrors); the op loop keeps A and B in registers, and a lossless pair
// repeated k >= 3 times is applied as the closed-form power of its period (pst_chain.cuh).
// The whole program (types, ops) arrives by value in the constant bank and the 48-byte optical rows are read straight
// through the read-only path (a warp shares its wavelength, so they are broadcast loads, prefetched one type ahead):
// no staging, no barrier, no global load on the dependent path.  (History in profiles/: v2 read the cached matrix
// from shared memory per layer and was LSU-bound; a per-layer switch over register-resident types was latency-bound;
// v7 staged tables with a bulk copy and fetched ops/flags from global memory -- 18 % of its stall samples.)
// ROW: the launch is a (theta tile, wavelength, sweep) grid -- a block shares its wavelength, so the wavelength index
// and every table address are block-uniform (no integer division, nothing to spill); used when theta tiles of 128
// waste little.  Otherwise blocks tile the flattened (wavelength, theta) index and blockIdx.y is the sweep index.
template <int NPOL, int NCOL, int MINB, bool ROW>
__global__ void __launch_bounds__(128, MINB)
tmm_typed_kernel(const GridParams prm, const TypedPlan tp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int NC = NCOL;
  const unsigned npts = (unsigned)prm.points, nth = (unsigned)prm.n_theta;
  const unsigned sweep_idx = ROW ? blockIdx.z : blockIdx.y;
  bool valid;
  unsigned p;
  int il_loc, it;
  if constexpr (ROW) {
    const unsigned itr = blockIdx.x * 128u + threadIdx.x;
    valid = itr < nth;
    it = (int)(valid ? itr : nth - 1u);  // tail lanes shadow the last angle (warp-wide reductions below)
    il_loc = (int)blockIdx.y;
    p = blockIdx.y * nth + (unsigned)it;
  } else {
    const unsigned praw = blockIdx.x * 128u + threadIdx.x;
    valid = praw < npts;
    p = valid ? praw : npts - 1u;  // tail lanes shadow the last point
    il_loc = (int)(p / nth);
    it = (int)(p - (unsigned)il_loc * nth);
  }
  const int il = prm.il0 + il_loc;
  const double2* optp = reinterpret_cast<const double2*>(prm.opt + (size_t)il * prm.n_mat * 6);  // rows of 3 double2
  const double2 tr = __ldg(prm.trig + it);
  const unsigned lossy_mask = __ldg(tp.lossy_types);  // bits 0..7: lossy layer types, bits 8..19: growth bound, bit 31
  const double ct = tr.x, ict = tr.y;
  const int row_last = tp.mat_last * 3;
  const double2 ex0 = __ldg(optp + row_last), ex1 = __ldg(optp + row_last + 1);
  const double nfr = ex0.y, nfi = ex1.y;
  const double2* in_row = optp + tp.mat_first * 3;  // incidence medium: 1/n for the epilogue
  double2 in1, in2;

  ChainState<NPOL, NC> st;
  init_exit_columns<NPOL, NC>(st, nfr, nfi, ct, prm.pol_first);

  {
    // ---- build each distinct layer matrix once
    unsigned char* my_types = smem_raw + (size_t)threadIdx.x * tp.type_stride;
    int* my_sc = reinterpret_cast<int*>(my_types + tp.n_types * kTypeRecBytes);
    const double d_sweep = prm.sweep_d ? __ldg(prm.sweep_d + sweep_idx) : 0.0;
    {
      double2 ra = __ldg(optp + tp.type_mat[0] * 3);
      for (int t = 0; t < tp.n_types; ++t) {
        const double2* row = optp + tp.type_mat[t] * 3;
        const double2 a = ra, b = __ldg(row + 1);  // (q.re, n.re), (1/n).re, n.im): same cache line as a
        if (t + 1 < tp.n_types) ra = __ldg(optp + tp.type_mat[t + 1] * 3);  // next type's line is in flight during the build
        const double d = (t == tp.sweep_type) ? d_sweep : tp.type_d[t];
        const bool lossy = (lossy_mask >> t) & 1u;
        TypeRegs<NPOL> tmp;
        if (!lossy) {
          tmp.set_real(build_real_layer<NPOL>(a.x, a.y, b.x, d, ct, ict, prm.pol_first));
        } else {
          const double2 c = __ldg(row + 2);
          OptRow o;
          o.qr = a.x; o.nr = a.y; o.inr = b.x; o.ni = b.y; o.qi = c.x; o.ini = c.y;
          tmp.set_cplx(build_cplx_layer<NPOL>(o, d, ct, ict, prm.pol_first));
        }
        double2* rec = reinterpret_cast<double2*>(my_types + t * kTypeRecBytes);
        if (!lossy) {
          rec[0] = make_double2(tmp.a[0], tmp.a[1]);
          rec[1] = make_double2(tmp.a[2], tmp.a[3]);
          rec[2] = make_double2(tmp.a[4], tmp.a[5]);
          rec[3] = make_double2(tmp.a[6], 0.0);
        } else {
  #pragma unroll
          for (int k = 0; k < 5; ++k) rec[k] = make_double2(tmp.a[2 * k], tmp.a[2 * k + 1]);
          my_sc[t] = tmp.sc;
        }
      }
    }
    auto load_type = [&](int t, bool lossy, TypeRegs<NPOL>& out) {
      const double2* rec = reinterpret_cast<const double2*>(my_types + t * kTypeRecBytes);
      if (!lossy) {
  #pragma unroll
        for (int k = 0; k < 4; ++k) { const double2 v = rec[k]; out.a[2 * k] = v.x; out.a[2 * k + 1] = v.y; }
        out.sc = 0;
      } else {
  #pragma unroll
        for (int k = 0; k < 5; ++k) { const double2 v = rec[k]; out.a[2 * k] = v.x; out.a[2 * k + 1] = v.y; }
        out.sc = my_sc[t];
      }
    };

    // Rescaling interval from the grid-wide growth bound (prep_kernel): after a rescale max|v| < 2 and n layers multiply
    // it by at most G^n, so n <= 900 / log2 G layers cannot overflow.  Block-uniform: it sets loop trip counts.
    // Ordinary mirrors need no rescale inside the chain at all.
    const int lg = growth_log2(growth_E_of_word(lossy_mask), (lossy_mask & 0xffu) != 0u);
    const int every = max(2, min(1024, __float2int_rz(899.5f * __frcp_rn((float)lg)))) & ~1;  // <= 900 / lg, no division

    PairPower S;
    S.t = 1.0; S.s = 0.0; S.ks = 0;
    int s_lo = -1, s_hi = -1, s_k = -1;  // the period S belongs to
    int s_first = -1;                   // the layer type applied first when S was computed (orientation of its sums)
    int since = 0;                      // layers applied since the last rescale
    for (int op = 0; op < tp.n_ops; ++op) {
      const int4 o = tp.ops[op];
      TypeRegs<NPOL> A, B;
      const bool la = (lossy_mask >> o.x) & 1u;
      load_type(o.x, la, A);
      if (o.y < 0) {
        for (int k = 0; k < o.z; ++k) {
          apply_type<NPOL, NC>(st, A, la);
          if (++since >= every) { renormalise<NPOL, NC>(st); since = 0; }
        }
        continue;
      }
      const bool lb = (lossy_mask >> o.y) & 1u;
      load_type(o.y, lb, B);
      if (!la && !lb && o.z >= 3 && !(prm.flags & 128u)) {
        // a lossless pair repeated k times: closed-form power of the period (binary powering in the ring spanned by I and
        // the traceless part of the period, pst_chain.cuh).  |U^k entries| < 2^(kPowerNoRescaleLog2 + 1) (coefficients
        // not rescaled) or < 2^4 (rescaled); rescale the state first only if that times its current bound 2 G^since
        // could leave the fp64 range
        if (since * lg > 1000 - kPowerNoRescaleLog2 - 20) renormalise<NPOL, NC>(st);
        const int t_lo = min(o.x, o.y), t_hi = max(o.x, o.y);
        Period U;
        if (t_lo != s_lo || t_hi != s_hi || o.z != s_k) {  // else: same or mirrored period, same (x, s2), same power
          period_form<NPOL>(A, B, false, U);
          pair_power_coeffs(U, o.z, lg, S);
          s_lo = t_lo; s_hi = t_hi; s_k = o.z;
          s_first = o.x;
        } else {
          period_form<NPOL>(A, B, o.x != s_first, U);  // a mirror image rounds like the period it mirrors
        }
        PairMatrix<NPOL> P;
        pair_power_matrix<NPOL>(U, S, ct, ict, prm.pol_first, P);
        apply_pair_matrix<NPOL, NC>(st, P);
        renormalise<NPOL, NC>(st);
        since = 0;
        continue;
      }
      int k = o.z;
      while (k > 0) {
        const int n = min(k, max((every - since) / 2, 1));
        if (!la && !lb) pair_loop<NPOL, NC, false, false>(st, A, B, n);
        else if (!la) pair_loop<NPOL, NC, false, true>(st, A, B, n);
        else if (!lb) pair_loop<NPOL, NC, true, false>(st, A, B, n);
        else pair_loop<NPOL, NC, true, true>(st, A, B, n);
        k -= n;
        since += 2 * n;
        if (since >= every) { renormalise<NPOL, NC>(st); since = 0; }
      }
    }
    if (since > 0) renormalise<NPOL, NC>(st);
    in1 = __ldg(in_row + 1);
    in2 = __ldg(in_row + 2);
  }
  if (!valid) return;

  // ---- epilogue
  const size_t plane = (size_t)prm.points;
#pragma unroll
  for (int pp = 0; pp < NPOL; ++pp) {
    const PointOut r = finish_point<NPOL, NCOL>(st, pp, in1.x, in2.y, nfr, nfi, ict, prm.pol_first, prm.flags);
    const size_t o = ((size_t)sweep_idx * (NPOL << prm.pol_dup) + pp) * plane + (size_t)p;
    store_point<false>(prm, o, r.R, r.T, r.A);
    if constexpr (NCOL == 2) {
      if (prm.M) {
        double4* mo = reinterpret_cast<double4*>(prm.M + o * 8);
        mo[0] = make_double4(r.M[0], r.M[1], r.M[2], r.M[3]);
        mo[1] = make_double4(r.M[4], r.M[5], r.M[6], r.M[7]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------ K1 cavity (straight line)
// The commonest typed program -- periodic mirror / spacer / periodic mirror over one pair of layer types (a DBR
// microcavity, BASELINE configs[1]) -- as straight-line code: no type records, no op loop.  It evaluates the same
// functions in the same order as tmm_typed_kernel's interpreter, so the results are bit-identical
// (PST_FLAG_NO_SHAPE_FASTPATH routes the same stack through the interpreter; tested).
//
// Row-mode launch (ROW): a block owns ONE wavelength and cp.tiles consecutive tiles of 128 angles.  Everything the
// program needs from the optical table -- the rows of the three layer types, the exit index, the incidence 1/n, the
// spacer thickness: 24 doubles -- is block-uniform; 24 threads fetch it into shared memory once, and every tile reads
// it back with broadcast 128-bit LDS at the point of use (asm volatile: the compiler must not hoist 40 registers' worth
// of loads out of the tile loop -- that was what spilled in round 1's tile loop).  The only per-thread global load of
// a tile is its (cos, 1/cos) pair, requested one tile ahead.  So the L2 round trip that every warp of the round-1
// kernel paid before its first instruction (13 global loads per thread, 12 % of a warp's life) is paid once per block.
//
// Register diet (80 -> 6 blocks/SM): the mirror layers are only ever needed through their polarisation-independent parts
// (cos phi, n sin phi, sin phi / n: pst_chain.cuh BareLayer) because the period is formed from those; the two mirror
// sincos chains interleave, the period is powered and turned into the matrix P (6 doubles) before the spacer is even
// built, and the spacer's matrix dies before the second application of P.  What is block-uniform and decided by
// device data (absorbing spacer? rescaled powering?) selects one of four compiled tile loops up front, so the loop
// bodies test nothing.
//
// Absorbing mirror layers (cold, out of line): the same kernel applies the pairs one by one -- no closed form exists.
template <int NPOL>
__device__ __noinline__ void cavity_lossy_mirrors(const double* rv, double dA, double dB, int k1, int k2, bool mirrored,
                                                  bool lossyA, bool lossyB, bool lossyC, double ct, double ict, int pol_first,
                                                  unsigned flags, double* res /* [NPOL][3] = R, T, A */) {
  // self-contained (own state, own epilogue): the hot path's state must never have its address taken
  ChainState<NPOL, 1> st;
  init_exit_columns<NPOL, 1>(st, rv[14], rv[15], ct, pol_first);
  TypeRegs<NPOL> A, B, C;
  auto build = [&](const OptRow& o, double d, bool lossy, TypeRegs<NPOL>& out) {
    if (!lossy) out.set_real(build_real_layer<NPOL>(o.qr, o.nr, o.inr, d, ct, ict, pol_first));
    else out.set_cplx(build_cplx_layer<NPOL>(o, d, ct, ict, pol_first));
  };
  build(OptRow{rv[0], rv[2], rv[3], rv[18], rv[19], rv[20]}, dA, lossyA, A);
  build(OptRow{rv[1], rv[4], rv[5], rv[21], rv[22], rv[23]}, dB, lossyB, B);
  build(OptRow{rv[6], rv[8], rv[9], rv[10], rv[11], rv[12]}, rv[7], lossyC, C);
  for (int i = 0; i < k1; ++i) {
    apply_type<NPOL, 1>(st, A, lossyA);
    apply_type<NPOL, 1>(st, B, lossyB);
    renormalise<NPOL, 1>(st);
  }
  apply_type<NPOL, 1>(st, C, lossyC);
  renormalise<NPOL, 1>(st);
  for (int i = 0; i < k2; ++i) {
    apply_type<NPOL, 1>(st, mirrored ? B : A, mirrored ? lossyB : lossyA);
    apply_type<NPOL, 1>(st, mirrored ? A : B, mirrored ? lossyA : lossyB);
    renormalise<NPOL, 1>(st);
  }
#pragma unroll
  for (int pp = 0; pp < NPOL; ++pp) {
    const PointOut r = finish_point<NPOL, 1>(st, pp, rv[16], rv[17], rv[14], rv[15], ict, pol_first, flags);
    res[3 * pp] = r.R; res[3 * pp + 1] = r.T; res[3 * pp + 2] = r.A;
  }
}

// MODE: 0 lossless spacer, 1 absorbing spacer, 2 absorbing mirror layers (cold);  RESC: rescaled powering;
// MIRRORED: the second mirror is the mirror image of the first (only looked at when the pair counts agree)
// NOTEST: the block's three phases are bounded inside sincos' fast range for every angle (|q d| < 2^28 with |cos| <= 1;
// decided once per block from the staged row), so the tile loop carries no range tests and no call to the out-of-line
// Payne-Hanek path: a possible call costs a convergence barrier (BSSY/BSYNC) around each sincos and pins the compiler's
// schedule there (measured: 506 -> 495 instructions per point, 0.1927 -> 0.1869 ms)
template <int NPOL, bool ROW, int MODE, bool RESC, bool MIRRORED, bool NOTEST = false>
__device__ __forceinline__ void cavity_tiles(const GridParams& prm, const CavityPlan& cp, unsigned s_base, unsigned lossy_mask,
                                             bool mid_rescale, bool calm_skips_rescale) {
  const unsigned npts = (unsigned)prm.points, nth = (unsigned)prm.n_theta;
  const unsigned sweep_idx = ROW ? blockIdx.z : blockIdx.y;
  const int n_tiles = ROW ? cp.tiles : 1;
  const unsigned tile0 = ROW ? blockIdx.x * (unsigned)n_tiles : 0u;
  const double* optr = nullptr;  // !ROW: this thread's block of the optical table
  unsigned p = 0;
  int it = 0;
  bool valid = true;
  if constexpr (!ROW) {
    const unsigned praw = blockIdx.x * 128u + threadIdx.x;
    valid = praw < npts;
    p = valid ? praw : npts - 1u;  // tail lanes shadow the last point
    const int il_loc = (int)(p / nth);
    it = (int)(p - (unsigned)il_loc * nth);
    optr = prm.opt + (size_t)(prm.il0 + il_loc) * prm.n_mat * 6;
  }
  // values i, i + 1 of the staged record (i even)
  auto val2 = [&](auto I) -> double2 {
    constexpr int i = decltype(I)::value;
    if constexpr (ROW) {
      double2 v;
      asm volatile("ld.shared.v2.f64 {%0, %1}, [%2+%3];" : "=d"(v.x), "=d"(v.y) : "r"(s_base), "n"(i * 8));
      return v;
    } else {
      double2 v;
      v.x = (i == 6 || cp.src[i] >= 0) ? __ldg(optr + cp.src[i]) : 0.0;
      if (i + 1 == 7) v.y = cp.c_swept ? __ldg(prm.sweep_d + sweep_idx) : cp.d[2];
      else v.y = cp.src[i + 1] >= 0 ? __ldg(optr + cp.src[i + 1]) : 0.0;
      return v;
    }
  };
#define PST_V2(i) val2(std::integral_constant<int, i>{})
  const size_t plane = (size_t)prm.points;
  const size_t out_base = (size_t)sweep_idx * (size_t)(NPOL << prm.pol_dup) * plane;
  // per-thread cursors, advanced by one tile (128 angles) per trip: the angle table is padded (kTrigPad), so tiles are
  // read whole and only the stores are predicated
  unsigned itr = ROW ? tile0 * 128u + threadIdx.x : 0u;
  const double2* trp = prm.trig + (ROW ? itr : (unsigned)it);
  size_t o = out_base + (ROW ? (size_t)blockIdx.y * nth + itr : (size_t)p);
  double2 tr_next = __ldg(trp);
  for (int tile = 0; tile < n_tiles; ++tile, itr += 128u, o += 128u) {
    if constexpr (ROW) {
      if (itr - threadIdx.x >= nth) break;  // block-uniform: ragged last block
      valid = itr < nth;
    }
    const double ct = tr_next.x, ict = tr_next.y;
    if constexpr (ROW) {
      trp += 128;
      if (tile + 1 < n_tiles) tr_next = __ldg(trp);
    }
    double outR[NPOL], outT[NPOL], outA[NPOL];
    if constexpr (MODE != 2) {
      const double2 nf = PST_V2(14);
      ChainState<NPOL, 1> st;
      init_exit_columns<NPOL, 1>(st, nf.x, nf.y, ct, prm.pol_first);
      // ---- first mirror: the two sincos chains interleave in one basic block
      PairMatrix<NPOL> P;
      Period U;
      bool calm = false;
      {
        const double2 q = PST_V2(0);
        const double xA = layer_phase(q.x, ct, cp.d[0]), xB = layer_phase(q.y, ct, cp.d[1]);
        double sA, cA, sB, cB;
        if (NOTEST || (sincos_in_fast_range(xA) && sincos_in_fast_range(xB))) {
          sincos_core(xA, sA, cA);
          sincos_core(xB, sB, cB);
        } else {
          sincos_fast(xA, sA, cA);
          sincos_fast(xB, sB, cB);
        }
        const double2 na = PST_V2(2), nb = PST_V2(4);
        const BareLayer A = finish_bare_layer(sA, cA, na.x, na.y), B = finish_bare_layer(sB, cB, nb.x, nb.y);
        period_form(A, B, false, U);
        PairPower S;
        if constexpr (RESC) {
          pair_power_coeffs_t<true>(U, cp.k1, cp.hb1, S);
        } else {
          if (!pair_power_small(U, cp.k1, S)) pair_power_coeffs_t<false>(U, cp.k1, cp.hb1, S);
        }
        pair_power_matrix<NPOL>(U, S, ct, ict, prm.pol_first, P);
        // |x| < 2 on every lane of the warp (tested on the high word; NaN fails): the Bloch eigenvalue is below 2 + sqrt 3
        if constexpr (!RESC) calm = __all_sync(0xffffffffu, ((unsigned)hi_word(U.x) & 0x7fffffffu) < 0x40000000u) != 0;
        // a second mirror with another pair count has a power of its own (see the interpreter's op loop); its period
        // rounds in the order its own first layer dictates and is formed now, while A and B are live
        if (cp.k2 != cp.k1 && cp.mirrored) period_form(B, A, false, U);
      }
      apply_pair_matrix<NPOL, 1, true>(st, P);
      if (mid_rescale) renormalise<NPOL, 1>(st);
      // ---- spacer
      {
        const double2 qd = PST_V2(6), nc = PST_V2(8);
        double sC, cC;
        if constexpr (NOTEST) sincos_core(layer_phase(qd.x, ct, qd.y), sC, cC);
        else sincos_fast(layer_phase(qd.x, ct, qd.y), sC, cC);
        if constexpr (MODE == 0) {
          apply_real_layer<NPOL, 1>(st, finish_real_layer<NPOL>(sC, cC, nc.x, nc.y, ct, ict, prm.pol_first));
        } else {
          const double2 c2 = PST_V2(10), c3 = PST_V2(12);
          const OptRow o = {qd.x, nc.x, nc.y, c2.x, c2.y, c3.x};
          apply_cplx_layer<NPOL, 1>(st, finish_cplx_layer<NPOL>(sC, cC, layer_phase(o.qi, ct, qd.y), o, ct, ict, prm.pol_first));
        }
      }
      if (mid_rescale) renormalise<NPOL, 1>(st);
      // ---- second mirror: the same matrix (exchanged diagonal for the mirror image: w -> -w) when the pair counts agree
      if (cp.k2 == cp.k1) {
        if constexpr (MIRRORED) { const double t = P.p00; P.p00 = P.p11; P.p11 = t; }
      } else {
        PairPower S2;  // powered only now: cheaper than holding a second matrix across the spacer
        pair_power_coeffs_t<RESC>(U, cp.k2, cp.hb2, S2);
        pair_power_matrix<NPOL>(U, S2, ct, ict, prm.pol_first, P);
      }
      apply_pair_matrix<NPOL, 1>(st, P);
      // The epilogue squares the entries of M: the state is rescaled first unless the whole product provably stays far
      // inside the fp64 range (warp-uniform; rescaling is exact, so the bits do not depend on the decision)
      if (!(calm_skips_rescale && calm)) renormalise<NPOL, 1>(st);
      {
        const double2 in0 = PST_V2(16), nf2 = PST_V2(14);
#pragma unroll
        for (int pp = 0; pp < NPOL; ++pp) {
          const PointOut r = finish_point<NPOL, 1>(st, pp, in0.x, in0.y, nf2.x, nf2.y, ict, prm.pol_first, prm.flags);
          outR[pp] = r.R; outT[pp] = r.T; outA[pp] = r.A;
        }
      }
    } else {
      double rv[kCavityRow], res[3 * NPOL];
      {
        const double2 v0 = PST_V2(0), v1 = PST_V2(2), v2 = PST_V2(4), v3 = PST_V2(6), v4 = PST_V2(8), v5 = PST_V2(10);
        const double2 v6 = PST_V2(12), v7 = PST_V2(14), v8 = PST_V2(16), v9 = PST_V2(18), v10 = PST_V2(20), v11 = PST_V2(22);
        const double2 all[12] = {v0, v1, v2, v3, v4, v5, v6, v7, v8, v9, v10, v11};
        for (int i = 0; i < 12; ++i) { rv[2 * i] = all[i].x; rv[2 * i + 1] = all[i].y; }
      }
      cavity_lossy_mirrors<NPOL>(rv, cp.d[0], cp.d[1], cp.k1, cp.k2, cp.mirrored != 0, (lossy_mask & cp.lossy_a) != 0u,
                                 (lossy_mask & (cp.lossy_ab & ~cp.lossy_a)) != 0u, (lossy_mask & cp.lossy_c) != 0u, ct, ict,
                                 prm.pol_first, prm.flags, res);
#pragma unroll
      for (int pp = 0; pp < NPOL; ++pp) { outR[pp] = res[3 * pp]; outT[pp] = res[3 * pp + 1]; outA[pp] = res[3 * pp + 2]; }
    }
    if (valid) {
      if (prm.plain_out) {  // the common call: three local arrays, no tests per value
        double* pR = prm.R + o;
        double* pT = prm.T + o;
        double* pA = prm.A + o;
#pragma unroll
        for (int pp = 0; pp < NPOL; ++pp) {
          pR[(size_t)pp * plane] = outR[pp];
          pT[(size_t)pp * plane] = outT[pp];
          pA[(size_t)pp * plane] = outA[pp];
        }
      } else {
#pragma unroll
        for (int pp = 0; pp < NPOL; ++pp) store_point<false>(prm, o + (size_t)pp * plane, outR[pp], outT[pp], outA[pp]);
      }
    }
  }
#undef PST_V2
}

template <int NPOL, int MINB, bool ROW>
__global__ void __launch_bounds__(128, MINB)
tmm_cavity_kernel(const GridParams prm, const CavityPlan cp) {
  __shared__ __align__(16) double s_row[kCavityRow];
  const unsigned lossy_mask = __ldg(cp.lossy_types);
  unsigned s_base = 0;
  if constexpr (ROW) {
    if (threadIdx.x < kCavityRow) {
      const int src = cp.src[threadIdx.x];
      double v = 0.0;
      if (src >= 0) v = __ldg(prm.opt + (size_t)(prm.il0 + (int)blockIdx.y) * prm.n_mat * 6 + src);
      else if (threadIdx.x == 7) v = cp.c_swept ? __ldg(prm.sweep_d + blockIdx.z) : cp.d[2];
      s_row[threadIdx.x] = v;
    }
    s_base = (unsigned)__cvta_generic_to_shared(s_row);
    asm volatile("" : "+r"(s_base));  // opaque: held in a register instead of being rebuilt from the window base per load
    __syncthreads();
  }
  const bool lossyAB = (lossy_mask & cp.lossy_ab) != 0u;
  const bool lossyC = (lossy_mask & cp.lossy_c) != 0u;
  // Rescaling schedule from the grid-wide growth bound (prep_kernel; block-uniform).  One power grows the state by
  // < 2^(kPowerNoRescaleLog2 + 1) (rescaled coefficients: < 2^4), the spacer by 2^lg: up to lg = 100 the three ops fit the
  // fp64 range without a rescale in between.  The interpreter rescales after every power; rescaling is exact, so the
  // bits are the same.
  const int lg = growth_log2(growth_E_of_word(lossy_mask), lossyC || lossyAB);
  const bool mid_rescale = lg > 100;
  const bool resc = pair_power_needs_rescale(max(cp.k1, cp.k2), lg);
  const bool mir = cp.mirrored != 0;
  // Growth of the whole program when |x| < 2 at a point (x = half the period's trace): U^k = t I + s W with
  // |t| <= mu^k, |s| <= k mu^(k-1), mu = |x| + sqrt(x^2 - 1) < 2^2, entries of W below 2^(2 lg)  ->  one mirror grows the
  // state by < 2^(2 k + hb + 2 lg + 4), the spacer by 2^lg, exit column and D_0^-1 by 2^E each: when the sum of the
  // exponents stays below 500 the squares |M00|^2, |M10|^2 cannot leave the fp64 range and the final rescaling is skipped
  // for warps whose 32 points all have |x| < 2 (every pass band and every stop band of ordinary mirrors).
  const int E = growth_E_of_word(lossy_mask);
  const bool calm_skips = !mid_rescale && 2 * (cp.k1 + cp.k2) + cp.hb1 + cp.hb2 + 5 * lg + 2 * E + 9 <= 500;
#define PST_CT(MODE, RESC, MIR) cavity_tiles<NPOL, ROW, MODE, RESC, MIR>(prm, cp, s_base, lossy_mask, mid_rescale, calm_skips)
#define PST_CTN(MODE, MIR) cavity_tiles<NPOL, ROW, MODE, false, MIR, true>(prm, cp, s_base, lossy_mask, mid_rescale, calm_skips)
  bool notest = false;
  if constexpr (ROW) {  // block-uniform: the staged row holds q_A, q_B, q_C.re and d_C of this wavelength
    const double bound = fmax(fmax(fabs(s_row[0] * cp.d[0]), fabs(s_row[1] * cp.d[1])), fabs(s_row[6] * s_row[7]));
    notest = bound < 268435456.0;  // 2^28: |cos(theta)| <= 1, a factor four of slack; NaN fails
  }
  if (lossyAB) PST_CT(2, false, false);
  else if (!resc && notest) {
    if (!lossyC) { if (mir) PST_CTN(0, true); else PST_CTN(0, false); }
    else { if (mir) PST_CTN(1, true); else PST_CTN(1, false); }
  } else if (!resc) {
    if (!lossyC) { if (mir) PST_CT(0, false, true); else PST_CT(0, false, false); }
    else { if (mir) PST_CT(1, false, true); else PST_CT(1, false, false); }
  } else {
    if (!lossyC) { if (mir) PST_CT(0, true, true); else PST_CT(0, true, false); }
    else { if (mir) PST_CT(1, true, true); else PST_CT(1, true, false); }
  }
#undef PST_CT
#undef PST_CTN
}

// ------------------------------------------------------------------------------------ refraction mode
// PST_FLAG_SNELL: one thread per spectral point, both polarisations, a plain loop over the
// layers with the refracted angle of every layer (pst_chain.cuh, build_snell_layer).  Every layer goes through the
// complex form (beyond the critical angle even a lossless layer has a complex phase).  Tables are read through the
// read-only path: a warp shares its wavelength when n_theta >= 32, so rows and layer entries are broadcast loads.
template <int NPOL, int NCOL>
__global__ void __launch_bounds__(128, 4)
tmm_snell_kernel(const GridParams prm) {
  const unsigned npts = (unsigned)prm.points, nth = (unsigned)prm.n_theta;
  const unsigned praw = blockIdx.x * 128u + threadIdx.x;
  if (praw >= npts) return;
  const unsigned p = praw;
  const int il_loc = (int)(p / nth);
  const int it = (int)(p - (unsigned)il_loc * nth);
  const double2 tr = __ldg(prm.trig + it);
  const double ct = tr.x, ict = tr.y;
  const int sweep = blockIdx.y;
  const double d_sweep = prm.sweep_d ? __ldg(prm.sweep_d + sweep) : 0.0;
  const double2* optp = reinterpret_cast<const double2*>(prm.opt + (size_t)(prm.il0 + il_loc) * prm.n_mat * 6);
  auto opt_row = [&](int mat) {
    const double2 a = __ldg(optp + mat * 3), b = __ldg(optp + mat * 3 + 1), c = __ldg(optp + mat * 3 + 2);
    OptRow r;
    r.qr = a.x; r.nr = a.y; r.inr = b.x; r.ni = b.y; r.qi = c.x; r.ini = c.y;
    return r;
  };
  const int L = prm.n_layers;
  const OptRow in_row = opt_row(prm.layers[0].mat), exit_row = opt_row(prm.layers[L - 1].mat);
  const cplx s2 = snell_s2(in_row.nr, in_row.ni, ct);
  const SnellAngle gf = snell_angle(exit_row, s2);

  ChainState<NPOL, NCOL> st;
  init_exit_columns_snell<NPOL, NCOL>(st, exit_row.nr, exit_row.ni, gf, prm.pol_first);
  int j = L - 2;
  while (j >= 1) {
    const int j_stop = max(j - kRenormEvery, 0);
    for (; j > j_stop; --j) {
      const int4 raw = __ldg(reinterpret_cast<const int4*>(prm.layers + j));  // {d, mat, type}
      const double d = (j == prm.sweep_layer) ? d_sweep : __hiloint2double(raw.y, raw.x);
      apply_cplx_layer<NPOL, NCOL>(st, build_snell_layer<NPOL>(opt_row(raw.z), d, s2, prm.pol_first));
    }
    renormalise<NPOL, NCOL>(st);
  }

  const double det_re = snell_det_re(gf, in_row.inr, in_row.ini, ict);
  const size_t plane = (size_t)prm.points;
#pragma unroll
  for (int pp = 0; pp < NPOL; ++pp) {
    const PointOut r = finish_point_det<NPOL, NCOL>(st, pp, in_row.inr, in_row.ini, det_re, ict, prm.pol_first, prm.flags);
    const size_t o = ((size_t)sweep * NPOL + pp) * plane + (size_t)p;
    store_point<false>(prm, o, r.R, r.T, r.A);
    if constexpr (NCOL == 2) {
      if (prm.M) {
        double4* mo = reinterpret_cast<double4*>(prm.M + o * 8);
        mo[0] = make_double4(r.M[0], r.M[1], r.M[2], r.M[3]);
        mo[1] = make_double4(r.M[4], r.M[5], r.M[6], r.M[7]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------ per-layer field amplitudes
// pst_field_amplitudes (the reference's YAML carries unused A0/B0 keys, fabry-perot.yaml:13-14).
// The state of the backward chain after layers N-2 .. j is the tangential field pair at the left boundary of layer j for
// a unit transmitted amplitude; D_j^-1 turns it into the forward/backward amplitudes (A_j, B_j) of layer j there
// (tm.py:190-215 for D_j).  Results are normalised to the incident amplitude: layer 0 gets (1, r), the exit medium
// (t, 0).  One thread per (point, polarisation) runs the chain twice -- once for the incident amplitude and its
// power-of-two exponent, once to emit the normalised pairs -- so nothing unnormalised (it can exceed the fp64 range in
// a stop band) is ever stored.  amp[pol][wl][theta][layer][2] complex128.
template <bool SNELL>
__global__ void __launch_bounds__(128)
field_amplitudes_kernel(const GridParams prm, int pol_a, int pol_b, double* __restrict__ amp) {
  const unsigned npts = (unsigned)prm.points, nth = (unsigned)prm.n_theta;
  const unsigned p = blockIdx.x * 128u + threadIdx.x;
  if (p >= npts) return;
  const int pol = blockIdx.y == 0 ? pol_a : pol_b;  // 0 = s, 1 = p
  const bool is_s = pol == 0;
  const int il_loc = (int)(p / nth);
  const int it = (int)(p - (unsigned)il_loc * nth);
  const double2 tr = __ldg(prm.trig + it);
  const double ct = tr.x, ict = tr.y;
  const double2* optp = reinterpret_cast<const double2*>(prm.opt + (size_t)(prm.il0 + il_loc) * prm.n_mat * 6);
  auto opt_row = [&](int mat) {
    const double2 a = __ldg(optp + mat * 3), b = __ldg(optp + mat * 3 + 1), c = __ldg(optp + mat * 3 + 2);
    OptRow r;
    r.qr = a.x; r.nr = a.y; r.inr = b.x; r.ni = b.y; r.qi = c.x; r.ini = c.y;
    return r;
  };
  const int L = prm.n_layers;
  const OptRow in_row = opt_row(prm.layers[0].mat), exit_row = opt_row(prm.layers[L - 1].mat);
  const cplx s2 = SNELL ? snell_s2(in_row.nr, in_row.ni, ct) : cplx{0.0, 0.0};
  // rows of 2 D_j^-1 = [[w0, w1], [w0, -w1]]:  s: w0 = 1, w1 = 1/(n cos)   p: w0 = 1/cos, w1 = 1/n
  auto dinv = [&](const OptRow& o, cplx& w0, cplx& w1) {
    const cplx inv_n = {o.inr, o.ini};
    if constexpr (SNELL) {
      const SnellAngle g = snell_angle(o, s2);
      w0 = is_s ? cplx{1.0, 0.0} : g.ic;
      w1 = is_s ? cmul(inv_n, g.ic) : inv_n;
    } else {
      w0 = is_s ? cplx{1.0, 0.0} : cplx{ict, 0.0};
      w1 = is_s ? cscale(inv_n, ict) : inv_n;
    }
  };
  auto step = [&](ChainState<1, 1>& st, int j) {
    const int4 raw = __ldg(reinterpret_cast<const int4*>(prm.layers + j));  // {d, mat, lossy flag}
    const double d = __hiloint2double(raw.y, raw.x);
    const OptRow o = opt_row(raw.z);
    if constexpr (SNELL) apply_cplx_layer<1, 1>(st, build_snell_layer<1>(o, d, s2, pol));
    else apply_layer<1, 1>(st, o, raw.w == 0, d, ct, ict, pol);
    renormalise<1, 1>(st);
  };
  auto init = [&](ChainState<1, 1>& st) {
    if constexpr (SNELL) init_exit_columns_snell<1, 1>(st, exit_row.nr, exit_row.ni, snell_angle(exit_row, s2), pol);
    else init_exit_columns<1, 1>(st, exit_row.nr, exit_row.ni, ct, pol);
  };
  auto amplitudes = [&](const ChainState<1, 1>& st, const OptRow& o, cplx& A, cplx& B) {
    cplx w0, w1;
    dinv(o, w0, w1);
    const cplx a = cscale(cmul(w0, {st.v[0][0][0], st.v[0][0][1]}), 0.5);
    const cplx b = cscale(cmul(w1, {st.v[0][0][2], st.v[0][0][3]}), 0.5);
    A = cadd(a, b);
    B = csub(a, b);
  };

  ChainState<1, 1> st;
  init(st);
  for (int j = L - 2; j >= 1; --j) step(st, j);
  cplx A0, B0;
  amplitudes(st, in_row, A0, B0);
  const cplx inv_A0 = crecip(A0);  // times 2^-kexp0
  const int kexp0 = st.kexp[0];

  double4* out = reinterpret_cast<double4*>(amp) + ((size_t)blockIdx.y * npts + p) * (size_t)L;
  auto emit = [&](int j, cplx A, cplx B, int kexp) {
    const cplx a = cmul(A, inv_A0), b = cmul(B, inv_A0);
    const int e = kexp - kexp0;
    out[j] = make_double4(scale2(a.re, e), scale2(a.im, e), scale2(b.re, e), scale2(b.im, e));
  };
  emit(0, A0, B0, kexp0);
  init(st);
  {
    cplx A, B;
    amplitudes(st, exit_row, A, B);
    emit(L - 1, A, {0.0, 0.0}, 0);  // B is zero by construction (rounding residue dropped)
  }
  for (int j = L - 2; j >= 1; --j) {
    step(st, j);
    const int4 raw = __ldg(reinterpret_cast<const int4*>(prm.layers + j));
    cplx A, B;
    amplitudes(st, opt_row(raw.z), A, B);
    emit(j, A, B, st.kexp[0]);
  }
}

// Per-thread constants of a thickness sweep (see tmm_sweep_kernel) and its inner loop.
template <int NPOL>
struct SweepConsts {
  double qcr, qci, det_re;
  cplx al0[NPOL], be0[NPOL], al1[NPOL], be1[NPOL];
  int kfix[NPOL];
};

// FAST: 0 = general stores (peers, multicast, plane duplication, any other subset); otherwise the mask of LOCAL arrays
// the loop writes with streaming stores and untested cursors -- 7 = R, T, A; 3 = R, T (what tm.py returns); 2 = T only
// (sweeps feeding the fused peak search: through the general path a T-only chunk ran 30 % SLOWER than writing all three)
template <int NPOL, bool LOSSLESS, int FAST>
__device__ __forceinline__ void sweep_loop(const GridParams& prm, const SweepConsts<NPOL>& k, int s_begin, int s_end,
                                           unsigned p) {
  const size_t plane = (size_t)prm.points;
  const size_t o0 = (size_t)s_begin * NPOL * plane + (size_t)p;
  double* pR = (FAST & 1) ? prm.R + o0 : nullptr;
  double* pT = (FAST & 2) ? prm.T + o0 : nullptr;
  double* pA = (FAST & 4) ? prm.A + o0 : nullptr;
  double d = __ldg(prm.sweep_d + s_begin);
  for (int sw = s_begin; sw < s_end; ++sw) {
    const double d_next = __ldg(prm.sweep_d + min(sw + 1, s_end - 1));  // next thickness in flight during this one
    double sx, cx;
    sincos_fast(mul_rn(k.qcr, d), sx, cx);
    cplx cphi = {cx, 0.0}, sphi = {sx, 0.0};
    int sc = 0;
    if constexpr (!LOSSLESS) {
      double ch, sh;
      sc = cosh_sinh_scaled(mul_rn(k.qci, d), ch, sh);  // cos, sin of (x + i y), both times 2^sc
      cphi = {cx * ch, -(sx * sh)};
      sphi = {sx * ch, cx * sh};
    }
#pragma unroll
    for (int pp = 0; pp < NPOL; ++pp) {
      cplx m00, m10;
      if constexpr (LOSSLESS) {
        m00 = {fma(cx, k.al0[pp].re, sx * k.be0[pp].re), fma(cx, k.al0[pp].im, sx * k.be0[pp].im)};
        m10 = {fma(cx, k.al1[pp].re, sx * k.be1[pp].re), fma(cx, k.al1[pp].im, sx * k.be1[pp].im)};
      } else {
        // cos(phi) a + sin(phi) b as one FMA chain per component (8 FP64 instructions per entry instead of the 10 of two
        // complex products and a sum): cos(phi) = (P1, -P2), sin(phi) = (P3, P4)
        const double P1 = cphi.re, P2 = -cphi.im, P3 = sphi.re, P4 = sphi.im;
        auto mix = [&](const cplx& a, const cplx& b) -> cplx {
          return {fma(P1, a.re, fma(P2, a.im, fma(P3, b.re, -(P4 * b.im)))),
                  fma(P1, a.im, fma(P3, b.im, fma(P4, b.re, -(P2 * a.re))))};
        };
        m00 = mix(k.al0[pp], k.be0[pp]);
        m10 = mix(k.al1[pp], k.be1[pp]);
      }
      const double den = fma(m00.re, m00.re, m00.im * m00.im);
      const double num = fma(m10.re, m10.re, m10.im * m10.im);
      const double iden = recip_nr(den);
      const double R = num * iden;
      const double T = scale_sq(k.det_re * iden, -(k.kfix[pp] + sc));
      if constexpr (FAST != 0) {
        if constexpr ((FAST & 1) != 0) { __stcs(pR, R); pR += plane; }
        if constexpr ((FAST & 2) != 0) { __stcs(pT, T); pT += plane; }
        if constexpr ((FAST & 4) != 0) { __stcs(pA, (1.0 - R) - T); pA += plane; }
      } else {
        store_point<true>(prm, ((size_t)sw * (NPOL << prm.pol_dup) + pp) * plane + (size_t)p, R, T, (1.0 - R) - T);
      }
    }
    d = d_next;
  }
}

// ------------------------------------------------------------------------------------ K3: thickness sweep
// One thread per (wavelength, theta) point loops over a chunk of sweep thicknesses.  Everything that does not depend
// on the swept layer is computed once per thread:
//   suffix  u0 = (M_{s+1} ... M_{N-2}) D_f[:, 0]                         (state entering the swept layer)
//   prefix  rows of D_0^-1 (M_1 ... M_{s-1}):  r0, r1  with  M00 = r0 . u, M10 = r1 . u
// and each thickness costs one layer-matrix build (shared by both polarisations), one layer step and two 2-vector
// dot products per polarisation.  Results of one chunk iteration are 32 consecutive doubles per warp and array.
template <int NPOL, bool STAGE_OPT, bool STAGE_LAYERS>
__global__ void __launch_bounds__(128, 4)
tmm_sweep_kernel(const GridParams prm, int n_sweep, int sweeps_per_block) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int PB = blockDim.x;
  const int tid = threadIdx.x;
  const unsigned npts = (unsigned)prm.points, nth = (unsigned)prm.n_theta;  // < 2^31 points per launch (host-enforced)
  const unsigned p0 = blockIdx.x * (unsigned)PB;
  const unsigned p_last = min(p0 + (unsigned)PB, npts) - 1u;
  const int l_lo = prm.il0 + (int)(p0 / nth);
  const int TL = prm.il0 + (int)(p_last / nth) - l_lo + 1;
  const int L = prm.n_layers;
  unsigned char* s_opt = smem_raw;
  const size_t off = STAGE_OPT ? (size_t)prm.tl_rows * prm.opt_row_bytes : 0;
  LayerEntry* s_layers = reinterpret_cast<LayerEntry*>(smem_raw + off);
  if constexpr (STAGE_OPT || STAGE_LAYERS)
    stage_block(prm.use_bulk != 0, s_opt,
                reinterpret_cast<const unsigned char*>(prm.opt + (size_t)l_lo * prm.n_mat * 6), STAGE_OPT ? TL : 0,
                prm.n_mat * 48, prm.opt_row_bytes, reinterpret_cast<unsigned char*>(s_layers),
                reinterpret_cast<const unsigned char*>(prm.layers), STAGE_LAYERS ? L * (int)sizeof(LayerEntry) : 0, tid,
                PB);
  if (p0 + tid >= npts) return;
  const unsigned p = p0 + tid;
  const int il_loc = (int)(p / nth);
  const int it = (int)(p - (unsigned)il_loc * nth);
  const int il = prm.il0 + il_loc;
  const double2 tr = __ldg(prm.trig + it);
  const double ct = tr.x, ict = tr.y;
  const unsigned char* optp = STAGE_OPT ? (s_opt + (size_t)(il - l_lo) * prm.opt_row_bytes)
                                        : reinterpret_cast<const unsigned char*>(prm.opt + (size_t)il * prm.n_mat * 6);
  const LayerEntry* lay = STAGE_LAYERS ? s_layers : prm.layers;
  auto opt_row = [&](int mat) {
    const double2* o = reinterpret_cast<const double2*>(optp + (size_t)mat * 48);
    const double2 a = o[0], b = o[1], c = o[2];
    OptRow r;
    r.qr = a.x; r.nr = a.y; r.inr = b.x; r.ni = b.y; r.qi = c.x; r.ini = c.y;
    return r;
  };
  auto lossy_of = [&](const LayerEntry& le) { return le.type; };  // flagged layer list (prep_kernel)
  const int js = prm.sweep_layer;
  const OptRow exit_row = opt_row(lay[L - 1].mat);
  const OptRow in_row = opt_row(lay[0].mat);

  // ---- suffix: exit column through the layers behind the swept one
  ChainState<NPOL, 1> suf;
  init_exit_columns<NPOL, 1>(suf, exit_row.nr, exit_row.ni, ct, prm.pol_first);
  {
    int since = 0;
    for (int j = L - 2; j > js; --j) {
      const LayerEntry le = lay[j];
      apply_layer<NPOL, 1>(suf, opt_row(le.mat), lossy_of(le) == 0, le.d, ct, ict, prm.pol_first);
      if (++since == kRenormEvery) { renormalise<NPOL, 1>(suf); since = 0; }
    }
    renormalise<NPOL, 1>(suf);
  }
  // ---- prefix: both columns of M_1 ... M_{s-1}, then the rows of D_0^-1 times it
  ChainState<NPOL, 2> pre;
  init_identity<NPOL>(pre);
  {
    int since = 0;
    for (int j = js - 1; j >= 1; --j) {
      const LayerEntry le = lay[j];
      apply_layer<NPOL, 2>(pre, opt_row(le.mat), lossy_of(le) == 0, le.d, ct, ict, prm.pol_first);
      if (++since == kRenormEvery) { renormalise<NPOL, 2>(pre); since = 0; }
    }
    renormalise<NPOL, 2>(pre);
  }
  cplx r0[NPOL][2], r1[NPOL][2];
#pragma unroll
  for (int pp = 0; pp < NPOL; ++pp) {
    const bool is_s = pol_is_s<NPOL>(pp, prm.pol_first);
    const double w0 = is_s ? 1.0 : ict;
    const cplx w1 = is_s ? cplx{in_row.inr * ict, in_row.ini * ict} : cplx{in_row.inr, in_row.ini};
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const double* v = pre.v[pp][c];
      const cplx a = {0.5 * w0 * v[0], 0.5 * w0 * v[1]};
      const cplx b = cscale(cmul(w1, {v[2], v[3]}), 0.5);
      r0[pp][c] = cadd(a, b);
      r1[pp][c] = csub(a, b);
    }
  }
  const double det_re = fma(exit_row.nr, in_row.inr, -(exit_row.ni * in_row.ini));
  const LayerEntry ls = lay[js];
  const OptRow srow = opt_row(ls.mat);
  const bool s_lossless = lossy_of(ls) == 0;

  // The swept layer is M_s = cos(phi) I + sin(phi) J with J = [[0, -i/a], [-i a, 0]], and a = n g does not depend on
  // the thickness.  With u the suffix vector and r0, r1 the prefix rows,
  //   M00 = r0 . (M_s u) = cos(phi) al0 + sin(phi) be0,   al0 = r0 . u,  be0 = r0 . (J u)      (M10 likewise with r1)
  // so one thickness costs one complex sincos (shared by both polarisations) and two 2-term complex sums per
  // polarisation; al, be are per-thread constants of the sweep.
  cplx al0[NPOL], be0[NPOL], al1[NPOL], be1[NPOL];
  int kfix[NPOL];
#pragma unroll
  for (int pp = 0; pp < NPOL; ++pp) {
    const bool is_s = pol_is_s<NPOL>(pp, prm.pol_first);
    const double g = is_s ? ct : ict, gi = is_s ? ict : ct;
    const cplx u0 = {suf.v[pp][0][0], suf.v[pp][0][1]}, u1 = {suf.v[pp][0][2], suf.v[pp][0][3]};
    const cplx a = {srow.nr * g, srow.ni * g}, ia = {srow.inr * gi, srow.ini * gi};
    const cplx t0 = cmul(ia, u1), t1 = cmul(a, u0);
    const cplx w0 = {t0.im, -t0.re}, w1 = {t1.im, -t1.re};  // J u = (-i u1 / a, -i a u0)
    al0[pp] = cadd(cmul(r0[pp][0], u0), cmul(r0[pp][1], u1));
    al1[pp] = cadd(cmul(r1[pp][0], u0), cmul(r1[pp][1], u1));
    be0[pp] = cadd(cmul(r0[pp][0], w0), cmul(r0[pp][1], w1));
    be1[pp] = cadd(cmul(r1[pp][0], w0), cmul(r1[pp][1], w1));
    kfix[pp] = suf.kexp[pp] + pre.kexp[pp];
  }
  const double qcr = mul_rn(srow.qr, ct), qci = mul_rn(srow.qi, ct);  // first rounding of layer_phase()

  // ---- the sweep loop, specialised on the swept layer's kind and on the output mode so that neither is tested per
  // thickness (FAST: R, T and A all go to local arrays with streaming stores)
  const int s_begin = blockIdx.y * sweeps_per_block;
  const int s_end = min(s_begin + sweeps_per_block, n_sweep);
  SweepConsts<NPOL> kc;
  kc.qcr = qcr; kc.qci = qci; kc.det_re = det_re;
#pragma unroll
  for (int pp = 0; pp < NPOL; ++pp) {
    kc.al0[pp] = al0[pp]; kc.al1[pp] = al1[pp]; kc.be0[pp] = be0[pp]; kc.be1[pp] = be1[pp]; kc.kfix[pp] = kfix[pp];
  }
  const bool local = prm.n_peers == 0 && !(prm.flags & kFlagMulticastOut) && !prm.pol_dup;
  const int mask = local ? ((prm.R ? 1 : 0) | (prm.T ? 2 : 0) | (prm.A ? 4 : 0)) : 0;
#define PST_SL(LOSSLESS)                                                              \
  if (mask == 7) sweep_loop<NPOL, LOSSLESS, 7>(prm, kc, s_begin, s_end, p);           \
  else if (mask == 3) sweep_loop<NPOL, LOSSLESS, 3>(prm, kc, s_begin, s_end, p);      \
  else if (mask == 2) sweep_loop<NPOL, LOSSLESS, 2>(prm, kc, s_begin, s_end, p);      \
  else sweep_loop<NPOL, LOSSLESS, 0>(prm, kc, s_begin, s_end, p);
  if (s_lossless) { PST_SL(true) } else { PST_SL(false) }
#undef PST_SL
}

// ----------------------------------------------------------------------- per-layer matrices (API parity)
// out[i] = [D, P, D^-1, M] (each 2x2 complex128, row major) for wavelength i.  Closed forms of what the
// reference obtains with np.linalg.inv and multi_dot (tm.py:272-275).
__global__ void layer_matrices_kernel(const double* __restrict__ n_re, const double* __restrict__ n_im,
                                      const double* __restrict__ wl, int n_wl, double ct, unsigned pol, double d,
                                      double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_wl) return;
  const double nr = n_re[i], ni = n_im[i];
  const double omega = div_rn(kTwoPiC, wl[i]);
  const double kxr = mul_rn(mul_rn(mul_rn(nr, omega), kInvC0), ct);
  const double kxi = mul_rn(mul_rn(mul_rn(ni, omega), kInvC0), ct);
  const double x = mul_rn(kxr, d), y = mul_rn(kxi, d);
  double sx, cx;
  sincos(x, &sx, &cx);
  const double ep = exp(y), em = exp(-y);
  const bool is_s = pol == 1u;
  const double ict = 1.0 / ct;
  // D
  const cplx d0 = is_s ? cplx{1.0, 0.0} : cplx{ct, 0.0};
  const cplx d1 = is_s ? cplx{nr * ct, ni * ct} : cplx{nr, ni};
  // D^-1 = 1/2 [[1/d0, 1/d1], [1/d0, -1/d1]]
  const cplx h0 = {0.5 / d0.re, 0.0};
  const cplx h1 = cscale(crecip(d1), 0.5);
  // P = diag(e^{y}(cos x - i sin x), e^{-y}(cos x + i sin x))                    (tm.py:234)
  const cplx p0 = {ep * cx, -(ep * sx)};
  const cplx p1 = {em * cx, em * sx};
  // M = [[cos phi, -i sin phi / a], [-i a sin phi, cos phi]], a = d1/d0
  const double ch = 0.5 * (ep + em), sh = 0.5 * (ep - em);
  const cplx cphi = {cx * ch, -(sx * sh)};
  const cplx sphi = {sx * ch, cx * sh};
  const cplx a = is_s ? d1 : cplx{nr * ict, ni * ict};
  const cplx as = cmul(a, sphi);
  const cplx sa = cmul(sphi, crecip(a));
  double* o = out + (size_t)i * 32;
  // D
  o[0] = d0.re; o[1] = 0.0; o[2] = d0.re; o[3] = 0.0;
  o[4] = d1.re; o[5] = d1.im; o[6] = -d1.re; o[7] = -d1.im;
  // P
  o[8] = p0.re; o[9] = p0.im; o[10] = 0.0; o[11] = 0.0;
  o[12] = 0.0; o[13] = 0.0; o[14] = p1.re; o[15] = p1.im;
  // D^-1
  o[16] = h0.re; o[17] = 0.0; o[18] = h1.re; o[19] = h1.im;
  o[20] = h0.re; o[21] = 0.0; o[22] = -h1.re; o[23] = -h1.im;
  // M
  o[24] = cphi.re; o[25] = cphi.im; o[26] = sa.im; o[27] = -sa.re;
  o[28] = as.im; o[29] = -as.re; o[30] = cphi.re; o[31] = cphi.im;
}

__global__ void propagation_kernel(const double* __restrict__ kx, int n, double dr, double di,
                                   double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // -1j*kx*d = (b, -a)(dr + i di) in numpy's complex-product order: (b dr + a di) + i (b di - a dr)   (tm.py:234)
  const double a = kx[2 * i], b = kx[2 * i + 1];
  const double y = add_rn(mul_rn(b, dr), mul_rn(a, di));
  const double x = sub_rn(mul_rn(a, dr), mul_rn(b, di));
  double sx, cx;
  sincos(x, &sx, &cx);
  const double ep = exp(y), em = exp(-y);
  double* o = out + (size_t)i * 8;
  o[0] = ep * cx; o[1] = -(ep * sx); o[2] = 0.0; o[3] = 0.0;
  o[4] = 0.0; o[5] = 0.0; o[6] = em * cx; o[7] = em * sx;
}

__global__ void wavenumbers_kernel(const double* __restrict__ n_re, const double* __restrict__ n_im,
                                   const double* __restrict__ wl, int n_wl, double ct, double stheta,
                                   double* __restrict__ kx, double* __restrict__ kz) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_wl) return;
  const double omega = div_rn(kTwoPiC, wl[i]);
  const double qr = mul_rn(mul_rn(n_re[i], omega), kInvC0), qi = mul_rn(mul_rn(n_im[i], omega), kInvC0);
  kx[2 * i] = mul_rn(qr, ct); kx[2 * i + 1] = mul_rn(qi, ct);
  kz[2 * i] = mul_rn(qr, stheta); kz[2 * i + 1] = mul_rn(qi, stheta);
}

// (T, R) from arbitrary matrices: t = 1/M00, r = M10/M00, T = Re(det M |t|^2)      (tm.py:482-486)
__global__ void t_r_kernel(const double* __restrict__ M, long long n, double* __restrict__ T, double* __restrict__ R) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double4 r0 = reinterpret_cast<const double4*>(M)[2 * i];      // M00, M01
  const double4 r1 = reinterpret_cast<const double4*>(M)[2 * i + 1];  // M10, M11
  const cplx m00 = {r0.x, r0.y}, m01 = {r0.z, r0.w}, m10 = {r1.x, r1.y}, m11 = {r1.z, r1.w};
  // scale by a power of two so that |M00|^2 cannot overflow/underflow (exact)
  const int e = max(min(exponent_of(fmax(fabs(m00.re), fabs(m00.im))), 1000), -1000);
  const double s = pow2i(-e);
  const cplx a = cscale(m00, s), b = cscale(m10, s), c = cscale(m01, s), d = cscale(m11, s);
  const double den = fma(a.re, a.re, a.im * a.im);
  const cplx det = csub(cmul(a, d), cmul(c, b));
  T[i] = det.re / den;
  R[i] = fma(b.re, b.re, b.im * b.im) / den;
}

// Ordered chain over explicit layer matrices, one thread per wavelength             (tm.py:559-568)
struct mat2 {
  cplx a, b, c, d;  // [[a, b], [c, d]]
};
__device__ __forceinline__ mat2 load_mat2(const double* p) {
  const double4 r0 = reinterpret_cast<const double4*>(p)[0], r1 = reinterpret_cast<const double4*>(p)[1];
  return {{r0.x, r0.y}, {r0.z, r0.w}, {r1.x, r1.y}, {r1.z, r1.w}};
}
__device__ __forceinline__ mat2 matmul2(const mat2& x, const mat2& y) {
  return {cadd(cmul(x.a, y.a), cmul(x.b, y.c)), cadd(cmul(x.a, y.b), cmul(x.b, y.d)),
          cadd(cmul(x.c, y.a), cmul(x.d, y.c)), cadd(cmul(x.c, y.b), cmul(x.d, y.d))};
}
__global__ void chain_matrices_kernel(const double* __restrict__ first_inv, const double* __restrict__ interior,
                                      const double* __restrict__ last_d, int n_int, int n_wl, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_wl) return;
  mat2 acc = {{1.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}, {1.0, 0.0}};
  for (int j = n_int - 1; j >= 0; --j) acc = matmul2(load_mat2(interior + ((size_t)j * n_wl + i) * 8), acc);
  const mat2 m = matmul2(load_mat2(first_inv + (size_t)i * 8), matmul2(acc, load_mat2(last_d + (size_t)i * 8)));
  double4* o = reinterpret_cast<double4*>(out + (size_t)i * 8);
  o[0] = make_double4(m.a.re, m.a.im, m.b.re, m.b.im);
  o[1] = make_double4(m.c.re, m.c.im, m.d.re, m.d.im);
}

// Interface matrices (1/t) [[1, r], [r, 1]] from given coefficients                 (tm.py:746-751)
__global__ void interface_matrix_kernel(const double* __restrict__ r, const double* __restrict__ t, long long n,
                                        double* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const cplx it = crecip({t[2 * i], t[2 * i + 1]});
  const cplx ro = cmul({r[2 * i], r[2 * i + 1]}, it);
  double* m = out + i * 8;
  m[0] = it.re; m[1] = it.im; m[2] = ro.re; m[3] = ro.im;
  m[4] = ro.re; m[5] = ro.im; m[6] = it.re; m[7] = it.im;
}

// Fresnel coefficients / interface matrices                                         (tm.py:730-751)
__global__ void fresnel_kernel(const double* __restrict__ n1, const double* __restrict__ n2,
                               const double* __restrict__ k1, const double* __restrict__ k2, long long n,
                               double* __restrict__ rt, double* __restrict__ D) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const cplx a = {n1[2 * i], n1[2 * i + 1]}, b = {n2[2 * i], n2[2 * i + 1]};
  const cplx k1x = {k1[2 * i], k1[2 * i + 1]}, k2x = {k2[2 * i], k2[2 * i + 1]};
  const cplx isum = crecip(cadd(k1x, k2x));
  const cplx rs = cmul(csub(k1x, k2x), isum);
  const cplx ts = cmul(cscale(k1x, 2.0), isum);
  const cplx a2k2 = cmul(cmul(a, a), k2x), b2k1 = cmul(cmul(b, b), k1x);
  const cplx ip = crecip(cadd(a2k2, b2k1));
  const cplx rp = cmul(csub(a2k2, b2k1), ip);
  const cplx tp = cmul(cscale(a2k2, 2.0), ip);
  double* o = rt + i * 8;
  o[0] = rs.re; o[1] = rs.im; o[2] = ts.re; o[3] = ts.im; o[4] = rp.re; o[5] = rp.im; o[6] = tp.re; o[7] = tp.im;
  if (D) {
    const cplx r[2] = {rs, rp}, t[2] = {ts, tp};
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const cplx it = crecip(t[q]);
      const cplx ro = cmul(r[q], it);
      double* m = D + (i * 2 + q) * 8;
      m[0] = it.re; m[1] = it.im; m[2] = ro.re; m[3] = ro.im;
      m[4] = ro.re; m[5] = ro.im; m[6] = it.re; m[7] = it.im;
    }
  }
}

// ------------------------------------------------------------------------------ fused spectral reduction
// Peak search along the wavelength axis of a result array laid out [outer][point][inner] (the grid layout with
// outer = (sweep, pol), point = wavelength, inner = theta): one thread per spectrum, lanes on adjacent `inner` so every
// step of the scan is one coalesced load.  Semantics of scipy.signal.find_peaks(y, height=, distance=) as used by the
// reference's "Cavity Length Determiner" notebook (cell 3); see oracle/peaks_oracle.py for the restated algorithm.
// When `reverse` is set the spectrum is scanned from the last sample to the first (ascending wavenumber for an
// ascending wavelength grid) and indices refer to that reversed order.
constexpr int kPeakCap = 128;
__global__ void find_peaks_kernel(const double* __restrict__ y, int n_outer, int n_pts, int n_inner, double height,
                                  int use_height, int distance, int max_peaks, int reverse,
                                  const double* __restrict__ axis, int* __restrict__ out_idx,
                                  double* __restrict__ out_height, int* __restrict__ out_count,
                                  double* __restrict__ out_spacing) {
  const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= (long long)n_outer * n_inner) return;
  const int outer = (int)(s / n_inner), inner = (int)(s - (long long)outer * n_inner);
  const double* base = y + ((size_t)outer * n_pts) * n_inner + inner;
  auto at = [&](int i) { return __ldg(base + (size_t)(reverse ? n_pts - 1 - i : i) * n_inner); };
  int idx[kPeakCap];
  double hh[kPeakCap];
  unsigned char keep[kPeakCap];
  int n = 0;
  bool overflow = false;
  // ---- local maxima with plateau midpoints, height filter.  One-sample-at-a-time state machine (a rise opens a
  // candidate plateau, a strict descent closes it as a peak at the plateau's midpoint; plateaus touching either end
  // never close, as in scipy's _local_maxima_1d), fed from batches of independent loads so that the scan is
  // bandwidth- rather than latency-bound.  emit(index, height) is called per peak that passes the height filter.
  auto scan = [&](auto&& emit) {
    if (n_pts < 3) return;
    constexpr int kBatch = 16;
    double prev = at(0);
    bool rising = false;
    int start = 0;
    // one sample of the state machine, branch-free except for the (rare) emission of a peak:
    //   x > prev            opens a plateau at i
    //   x == prev           keeps the state
    //   x < prev (or NaN)   closes it; a descent from an open plateau is a peak at the plateau's midpoint
    auto feed = [&](int i, double x) {
      const bool gt = x > prev, lt = x < prev, ne = x != prev;
      if (lt && rising && (!use_height || prev >= height)) emit((start + i - 1) / 2, prev);
      rising = gt || (rising && !ne);
      start = gt ? i : start;
      prev = x;
    };
    int i0 = 1;
    for (; i0 + kBatch <= n_pts; i0 += kBatch) {
      double buf[kBatch];
#pragma unroll
      for (int k = 0; k < kBatch; ++k) buf[k] = at(i0 + k);
#pragma unroll
      for (int k = 0; k < kBatch; ++k) feed(i0 + k, buf[k]);
    }
    for (; i0 < n_pts; ++i0) feed(i0, at(i0));
  };
  // max_peaks bounds the peaks REPORTED; the candidates a distance condition thins out may be more (up to kPeakCap here)
  const int cand_cap = distance > 1 ? kPeakCap : max_peaks;
  scan([&](int i, double h) {
    if (n < cand_cap) { idx[n] = i; hh[n] = h; keep[n] = 1; ++n; }
    else overflow = true;
  });
  if (overflow && distance > 1) {
    // More candidates than the buffer holds (a rippled spectrum) and a distance condition that may leave only a few:
    // exact selection without a candidate list.  scipy visits candidates from the highest down (ties: the later one
    // first, as below) and a visited candidate that is still alive removes its neighbours closer than `distance`; so
    // the kept set is built by repeatedly taking the highest candidate not within `distance` of one already kept --
    // one re-scan of the spectrum (L2-resident) per kept peak.  Cold path: only the overflowing threads run it.
    n = 0;
    overflow = false;
    for (;;) {
      int bi = -1;
      double bh = 0.0;
      scan([&](int i, double h) {
        if (bi >= 0 && h < bh) return;
        for (int k = 0; k < n; ++k)
          if (abs(idx[k] - i) < distance) return;  // removed by (or is) a kept peak
        bi = i;
        bh = h;
      });
      if (bi < 0) break;
      if (n == max_peaks) { overflow = true; break; }
      int k = n++;  // kept peaks stay sorted by position
      for (; k > 0 && idx[k - 1] > bi; --k) { idx[k] = idx[k - 1]; hh[k] = hh[k - 1]; }
      idx[k] = bi;
      hh[k] = bh;
    }
    for (int k = 0; k < n; ++k) keep[k] = 1;
  } else if (distance > 1 && n > 1) {
    // ---- distance condition: from the highest peak down, a kept peak removes neighbours closer than `distance`
    unsigned char done[kPeakCap];
    for (int k = 0; k < n; ++k) done[k] = 0;
    for (int it = 0; it < n; ++it) {
      int j = -1;
      double best = 0.0;
      for (int k = 0; k < n; ++k)
        if (!done[k] && keep[k] && (j < 0 || hh[k] >= best)) { j = k; best = hh[k]; }
      if (j < 0) break;
      done[j] = 1;
      for (int k = j - 1; k >= 0 && idx[j] - idx[k] < distance; --k) keep[k] = 0;
      for (int k = j + 1; k < n && idx[k] - idx[j] < distance; ++k) keep[k] = 0;
    }
  }
  int m = 0, first = -1, last = -1;
  for (int k = 0; k < n; ++k)
    if (keep[k]) {
      if (m == max_peaks) { overflow = true; break; }
      out_idx[s * max_peaks + m] = idx[k];
      if (out_height) out_height[s * max_peaks + m] = hh[k];
      if (first < 0) first = idx[k];
      last = idx[k];
      ++m;
    }
  out_count[s] = overflow ? -1 : m;
  if (out_spacing) {
    double sp = __longlong_as_double(0x7ff8000000000000LL);  // NaN for fewer than two peaks
    if (axis && m >= 2) {
      const double a0 = __ldg(axis + (reverse ? n_pts - 1 - first : first));
      const double a1 = __ldg(axis + (reverse ? n_pts - 1 - last : last));
      sp = (a1 - a0) / (double)(m - 1);  // mean of consecutive differences telescopes
    }
    out_spacing[s] = sp;
  }
}

// Lower / upper polariton branch of every spectrum (the reference's "Polariton Peak Analysis"
// notebook fits two Lorentzians per angle and reads off their centres, cells 8-17).  Here the two HIGHEST peaks that
// pst_find_peaks kept are each refined by the exact three-point Lorentzian: for y = A / (1 + ((x - x0)/g)^2) the
// reciprocal 1/y is a parabola in x, so the peak sample and its two neighbours give x0, g (half width at half maximum)
// and A in closed form, on a non-uniform axis (wavenumbers of a wavelength grid) through divided differences.
// out[s][6] = (x0_lower, x0_upper, g_lower, g_upper, A_lower, A_upper), NaN when the spectrum has fewer than two
// peaks; "lower"/"upper" by position on `axis`.
__global__ void polariton_branches_kernel(const double* __restrict__ y, int n_outer, int n_pts, int n_inner, int reverse,
                                          const double* __restrict__ axis, const int* __restrict__ peak_idx,
                                          const double* __restrict__ peak_height, const int* __restrict__ peak_count,
                                          int max_peaks, double* __restrict__ out) {
  const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= (long long)n_outer * n_inner) return;
  const int outer = (int)(s / n_inner), inner = (int)(s - (long long)outer * n_inner);
  const double* base = y + ((size_t)outer * n_pts) * n_inner + inner;
  auto at = [&](int i) { return __ldg(base + (size_t)(reverse ? n_pts - 1 - i : i) * n_inner); };
  auto ax = [&](int i) { return __ldg(axis + (reverse ? n_pts - 1 - i : i)); };
  const double nan = __longlong_as_double(0x7ff8000000000000LL);
  double* o = out + s * 6;
  const int n = peak_count[s];
  if (n < 2) {
    for (int k = 0; k < 6; ++k) o[k] = nan;
    return;
  }
  // the two highest kept peaks (ties: the earlier one)
  int j1 = 0, j2 = -1;
  for (int k = 1; k < n; ++k)
    if (peak_height[s * max_peaks + k] > peak_height[s * max_peaks + j1]) j1 = k;
  for (int k = 0; k < n; ++k)
    if (k != j1 && (j2 < 0 || peak_height[s * max_peaks + k] > peak_height[s * max_peaks + j2])) j2 = k;
  double res[2][3];
  int pos[2] = {peak_idx[s * max_peaks + j1], peak_idx[s * max_peaks + j2]};
  for (int b = 0; b < 2; ++b) {
    const int i = pos[b];  // 1 <= i <= n_pts - 2: a local maximum is never an end sample
    const double x0 = ax(i - 1), x1 = ax(i), x2 = ax(i + 1);
    const double z0 = 1.0 / at(i - 1), z1 = 1.0 / at(i), z2 = 1.0 / at(i + 1);
    const double s01 = (z1 - z0) / (x1 - x0), s12 = (z2 - z1) / (x2 - x1);
    const double a = (s12 - s01) / (x2 - x0);                 // curvature of 1/y
    const double b1 = s01 + a * (x1 - x0);                    // slope of the parabola at x1
    if (a > 0.0 && isfinite(a) && isfinite(b1)) {
      const double du = -b1 / (2.0 * a);                      // centre relative to x1
      const double zmin = z1 - b1 * b1 / (4.0 * a);
      res[b][0] = x1 + du;
      res[b][1] = zmin > 0.0 ? sqrt(zmin / a) : nan;
      res[b][2] = zmin > 0.0 ? 1.0 / zmin : nan;
    } else {  // not Lorentzian-like around the sample (flat top, non-positive neighbours): report the sample itself
      res[b][0] = x1;
      res[b][1] = nan;
      res[b][2] = at(i);
    }
  }
  const int lo = res[0][0] <= res[1][0] ? 0 : 1, hi = 1 - lo;
  o[0] = res[lo][0]; o[1] = res[hi][0];
  o[2] = res[lo][1]; o[3] = res[hi][1];
  o[4] = res[lo][2]; o[5] = res[hi][2];
}

// Joint two-Lorentzian fit per spectrum (pst_fit.cuh; the reference's "Polariton Peak Analysis" notebook, cells 8-17): one
// thread per spectrum, lanes on adjacent `inner` (theta) so that every pass over the window is coalesced loads; seeds are
// the rows pst_polariton_branches wrote.  The window [i_lo, i_hi) counts samples in scan order (`reverse` as in
// find_peaks_kernel).  out[s][8] = (center_lower, center_upper, sigma_lower, sigma_upper, amplitude_lower,
// amplitude_upper, chi^2, evaluations) in lmfit's LorentzianModel convention.
__global__ void __launch_bounds__(128)
fit_two_lorentzians_kernel(const double* __restrict__ y, int n_outer, int n_pts, int n_inner, int reverse,
                           const double* __restrict__ axis, const double* __restrict__ seed, int i_lo, int i_hi, double sigma0,
                           double tol, int max_iter, double* __restrict__ out) {
  const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= (long long)n_outer * n_inner) return;
  const int outer = (int)(s / n_inner), inner = (int)(s - (long long)outer * n_inner);
  const int m0 = reverse ? n_pts - 1 - i_lo : i_lo;  // memory index of the window's first sample
  FitProblem pb;
  pb.y = y + ((size_t)outer * n_pts + m0) * n_inner + inner;
  pb.y_stride = reverse ? -(long long)n_inner : (long long)n_inner;
  pb.x = axis + m0;
  pb.x_stride = reverse ? -1 : 1;
  pb.n = i_hi - i_lo;
  double sd[6], res[8];
#pragma unroll
  for (int k = 0; k < 6; ++k) sd[k] = seed[s * 6 + k];
  fit_two_lorentzians(pb, sd, sigma0, tol, max_iter, res);
#pragma unroll
  for (int k = 0; k < 8; ++k) out[s * 8 + k] = res[k];
}

// ------------------------------------------------------------------------------ measurement helpers
// 8 independent dependent-FMA chains per thread: the non-tensor FP64 peak of the chip.
__global__ void dfma_probe_kernel(int iters, double seed, double* __restrict__ sink) {
  double a0 = seed, a1 = seed + 1, a2 = seed + 2, a3 = seed + 3, a4 = seed + 4, a5 = seed + 5, a6 = seed + 6, a7 = seed + 7;
  const double m = 0.9999999, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
      a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
  }
  const double s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
  if (s == 123.456) sink[0] = s;  // never true; keeps the chains alive
}

}  // namespace pst

#include "pst_f32.cuh"
