// C ABI of libpistachio_b200.so (declared in include/pistachio_b200.h).
// Host-side responsibilities only: argument validation, scratch memory, kernel-variant choice,
// launch geometry, and the host-buffer pipeline.  No CPU implementation of any numeric step lives here.
#include "../../include/pistachio_b200.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "pst_kernels.cuh"

using namespace pst;

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

#define PST_CUDA(call)                                                                              \
  do {                                                                                              \
    cudaError_t e__ = (call);                                                                       \
    if (e__ != cudaSuccess)                                                                         \
      return fail(PST_E_CUDA, std::string(#call) + ": " + cudaGetErrorName(e__) + " - " + cudaGetErrorString(e__)); \
  } while (0)

#define PST_TRY(expr)          \
  do {                         \
    int rc__ = (expr);         \
    if (rc__ != PST_OK) return rc__; \
  } while (0)

struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    ok = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;  // elements
  int ensure(size_t n) {
    if (n <= cap) return PST_OK;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = std::max(n, (size_t)256);
    cudaError_t e = cudaMalloc(&p, want * sizeof(T));
    if (e != cudaSuccess) return fail(PST_E_NOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    cap = want;
    return PST_OK;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

}  // namespace

struct pst_handle_s {
  int device = 0;
  int sm_count = 148;
  size_t smem_optin = 0;
  int64_t launches = 0;
  DevBuf<double> opt;
  DevBuf<double2> trig;
  DevBuf<LayerEntry> layers_dev;
  DevBuf<LayerEntry> layers_flagged;    // same list with `type` = the material's lossy flag (untyped kernels)
  std::vector<LayerEntry> layers_host;  // what layers_dev currently holds
  DevBuf<int> taboff;
  DevBuf<int> mat_flags;
  DevBuf<unsigned> prep_ticket;
  int n_types = 0;                 // layer-type analysis of the current layer list (sync_layers)
  int type_layer[kMaxTypes] = {0};
  std::vector<int4> ops_host;      // run-length ops of the typed path (passed by value: TypedPlan)
  DevBuf<unsigned> lossy_types;    // bit t: the material of layer type t absorbs (prep_kernel)
  DevBuf<unsigned char> l2buf;
  DevBuf<double> sink;
  // host-buffer pipeline (pst_eval_grid_host)
  cudaStream_t s_compute = nullptr, s_copy = nullptr;
  cudaEvent_t ev_done[2] = {nullptr, nullptr}, ev_copied[2] = {nullptr, nullptr};
  DevBuf<double> hp_in;       // wl | theta | cos | mat_n | mat_k | sweep_d
  DevBuf<double> hp_out[2];   // per-chunk R|T|A (|M)
  // Ordering of the scratch above across streams: every entry point that touches it first makes its stream wait for
  // the event recorded behind the previous such call (scratch_acquire) and records the event again behind its own
  // work (scratch_release).  Calls on one stream pay one event record; a call on another stream -- or
  // pst_eval_grid_host, which runs on internal streams -- queues behind the work that still reads the tables.
  cudaEvent_t ev_scratch = nullptr;
  cudaStream_t scratch_stream = nullptr;
  bool scratch_busy = false;
};

namespace {
int scratch_acquire(pst_handle_t h, cudaStream_t st) {
  if (!h->ev_scratch) PST_CUDA(cudaEventCreateWithFlags(&h->ev_scratch, cudaEventDisableTiming));
  if (h->scratch_busy && h->scratch_stream != st) PST_CUDA(cudaStreamWaitEvent(st, h->ev_scratch, 0));
  return PST_OK;
}
int scratch_release(pst_handle_t h, cudaStream_t st) {
  PST_CUDA(cudaEventRecord(h->ev_scratch, st));
  h->scratch_stream = st;
  h->scratch_busy = true;
  return PST_OK;
}
}  // namespace

// ------------------------------------------------------------------------------------------------ lifetime
extern "C" int pst_abi_version(void) { return PST_ABI_VERSION; }
extern "C" const char* pst_last_error(void) { return g_err.c_str(); }

extern "C" int pst_create(int device, pst_handle_t* out) {
  if (!out) return fail(PST_E_INVALID, "pst_create: out is NULL");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail(PST_E_NODEVICE, std::string("pst_create: no CUDA device visible (") + cudaGetErrorString(e) +
                                    "); this library has no CPU fallback");
  if (device < 0 || device >= count) return fail(PST_E_INVALID, "pst_create: device ordinal out of range");
  cudaDeviceProp prop;
  PST_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(PST_E_NODEVICE, std::string("pst_create: device '") + prop.name + "' is sm_" + std::to_string(prop.major) +
                                    std::to_string(prop.minor) + "; this build carries sm_100a code only");
  pst_handle_s* h = new (std::nothrow) pst_handle_s();
  if (!h) return fail(PST_E_NOMEM, "pst_create: out of host memory");
  h->device = device;
  h->sm_count = prop.multiProcessorCount;
  h->smem_optin = prop.sharedMemPerBlockOptin;
  *out = h;
  return PST_OK;
}

extern "C" int pst_destroy(pst_handle_t h) {
  if (!h) return PST_OK;
  DeviceGuard g(h->device);
  cudaDeviceSynchronize();
  h->opt.release(); h->trig.release(); h->layers_dev.release(); h->layers_flagged.release(); h->taboff.release(); h->mat_flags.release(); h->prep_ticket.release(); h->lossy_types.release(); h->l2buf.release();
  h->sink.release(); h->hp_in.release(); h->hp_out[0].release(); h->hp_out[1].release();
  for (int i = 0; i < 2; ++i) {
    if (h->ev_done[i]) cudaEventDestroy(h->ev_done[i]);
    if (h->ev_copied[i]) cudaEventDestroy(h->ev_copied[i]);
  }
  if (h->s_compute) cudaStreamDestroy(h->s_compute);
  if (h->s_copy) cudaStreamDestroy(h->s_copy);
  if (h->ev_scratch) cudaEventDestroy(h->ev_scratch);
  delete h;
  return PST_OK;
}

extern "C" int64_t pst_launch_count(pst_handle_t h) { return h ? h->launches : -1; }

// ------------------------------------------------------------------------------------------------ K0
extern "C" int pst_resample_tables(pst_handle_t h, int n_tab, const int* tab_off, const double* tab_wl,
                                   const double* tab_n, const double* tab_k, const double* wl, int n_wl,
                                   double* out_n, double* out_k, void* stream) {
  if (!h) return fail(PST_E_INVALID, "pst_resample_tables: NULL handle");
  if (n_tab <= 0 || n_wl <= 0) return fail(PST_E_INVALID, "pst_resample_tables: n_tab and n_wl must be positive");
  if (!tab_off || !tab_wl || !tab_n || !tab_k || !wl || !out_n || !out_k)
    return fail(PST_E_INVALID, "pst_resample_tables: NULL pointer argument");
  if (n_tab > 65535) return fail(PST_E_INVALID, "pst_resample_tables: at most 65535 tables per call");
  int max_rows = 0;
  for (int t = 0; t < n_tab; ++t) {
    int rows = tab_off[t + 1] - tab_off[t];
    if (rows < 1) return fail(PST_E_INVALID, "pst_resample_tables: every table needs at least one row");
    max_rows = std::max(max_rows, rows);
  }
  DeviceGuard g(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  PST_TRY(scratch_acquire(h, st));
  PST_TRY(h->taboff.ensure(n_tab + 1));
  PST_CUDA(cudaMemcpyAsync(h->taboff.p, tab_off, (n_tab + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
  const size_t smem = (size_t)max_rows * sizeof(double) <= 32 * 1024 ? (size_t)max_rows * sizeof(double) : 0;
  dim3 grid((n_wl + 255) / 256, n_tab);
  resample_kernel<<<grid, 256, smem, st>>>(n_tab, h->taboff.p, tab_wl, tab_n, tab_k, wl, n_wl, out_n, out_k);
  PST_CUDA(cudaGetLastError());
  h->launches += 1;
  return scratch_release(h, st);
}

// ------------------------------------------------------------------------------------------------ grid
namespace {

using ChainKernel = void (*)(const GridParams);

template <int P, int C>
ChainKernel chain_variant(bool so, bool sl, bool seg) {
#define PST_V(O, L, S) \
  if (so == O && sl == L && seg == S) return tmm_chain_kernel<P, C, O, L, S>;
#define PST_V2(O, L) PST_V(O, L, false) PST_V(O, L, true)
  PST_V2(true, true) PST_V2(true, false) PST_V2(false, true) PST_V2(false, false)
#undef PST_V2
#undef PST_V
  return nullptr;
}

ChainKernel pick_chain(int npol, int ncol, bool so, bool sl, bool seg) {
  if (npol == 1 && ncol == 1) return chain_variant<1, 1>(so, sl, seg);
  if (npol == 1 && ncol == 2) return chain_variant<1, 2>(so, sl, seg);
  if (npol == 2 && ncol == 1) return chain_variant<2, 1>(so, sl, seg);
  if (npol == 2 && ncol == 2) return chain_variant<2, 2>(so, sl, seg);
  return nullptr;
}

using CavityKernel = void (*)(const GridParams, const CavityPlan);
CavityKernel pick_cavity(int npol, bool row) {
  // 2 polarisations: 80 registers, 6 blocks/SM; 1 polarisation: 64 registers, 8 blocks/SM
#ifndef PST_CAV_MINB2
#define PST_CAV_MINB2 6
#endif
  if (npol == 2) return row ? tmm_cavity_kernel<2, PST_CAV_MINB2, true> : tmm_cavity_kernel<2, PST_CAV_MINB2, false>;
  if (npol == 1) return row ? tmm_cavity_kernel<1, 8, true> : tmm_cavity_kernel<1, 8, false>;
  return nullptr;
}

using TypedKernel = void (*)(const GridParams, const TypedPlan);
TypedKernel pick_typed(int npol, int ncol, bool row) {
#define PST_T(P, C, B) if (npol == P && ncol == C) return row ? tmm_typed_kernel<P, C, B, true> : tmm_typed_kernel<P, C, B, false>;
  PST_T(1, 1, 5)
  PST_T(1, 2, 4)
  PST_T(2, 1, 6)  // 80 registers, 6 blocks/SM: +5 % over 5 blocks now that a symmetric cavity keeps one period matrix live
  PST_T(2, 2, 3)
#undef PST_T
  return nullptr;
}

int validate_grid(const pst_grid_args* a, const char* who) {
  std::string w(who);
  if (!a) return fail(PST_E_INVALID, w + ": args is NULL");
  if (a->struct_size != sizeof(pst_grid_args))
    return fail(PST_E_INVALID, w + ": struct_size mismatch (header/library ABI skew)");
  if (a->n_wl <= 0 || a->n_theta <= 0 || a->n_mat <= 0) return fail(PST_E_INVALID, w + ": n_wl, n_theta, n_mat must be positive");
  if (a->n_layers < 2) return fail(PST_E_INVALID, w + ": a structure needs at least an incidence and an exit layer (n_layers >= 2)");
  if (a->pol_mask == 0 || (a->pol_mask & ~3u)) return fail(PST_E_INVALID, w + ": pol_mask must be PST_POL_S, PST_POL_P or both");
  if (!a->wl || !a->mat_n || !a->mat_k || !a->layer_mat || !a->layer_d) return fail(PST_E_INVALID, w + ": NULL input pointer");
  if (!a->theta && !a->cos_theta) return fail(PST_E_INVALID, w + ": theta and cos_theta are both NULL");
  if (a->peers) {
    const pst_peer_outputs* po = a->peers;
    if (po->n_peers < 1 || po->n_peers > PST_MAX_PEERS) return fail(PST_E_INVALID, w + ": peers->n_peers must be 1..PST_MAX_PEERS");
    if (a->R || a->T || a->A || a->M) return fail(PST_E_INVALID, w + ": with peers, R, T, A and M must be NULL (results go to the peer buffers)");
    if (a->flags & PST_FLAG_MULTICAST_OUT) return fail(PST_E_INVALID, w + ": peers and PST_FLAG_MULTICAST_OUT exclude each other");
    bool any = false;
    for (int i = 0; i < po->n_peers; ++i) any = any || po->R[i] || po->T[i] || po->A[i];
    if (!any) return fail(PST_E_INVALID, w + ": no output requested");
  } else if (!a->R && !a->T && !a->A && !a->M) {
    return fail(PST_E_INVALID, w + ": no output requested");
  }
  if (a->sweep_layer >= 0) {
    if (a->sweep_layer < 1 || a->sweep_layer > a->n_layers - 2)
      return fail(PST_E_INVALID, w + ": sweep_layer must be an interior layer");
    if (a->n_sweep <= 0 || !a->sweep_d) return fail(PST_E_INVALID, w + ": sweep needs n_sweep > 0 and sweep_d");
  } else if (a->n_sweep != 1) {
    return fail(PST_E_INVALID, w + ": n_sweep must be 1 without a sweep_layer");
  }
  if ((a->flags & PST_FLAG_FORCE_POINT_KERNEL) && (a->flags & PST_FLAG_FORCE_DEEP_KERNEL))
    return fail(PST_E_INVALID, w + ": contradictory kernel flags");
  if ((a->flags & PST_FLAG_MULTICAST_OUT) && a->M) return fail(PST_E_INVALID, w + ": PST_FLAG_MULTICAST_OUT does not cover the matrix output M");
  if ((a->flags & PST_FLAG_POL_DEGENERATE) && (a->M || a->pol_mask != (PST_POL_S | PST_POL_P)))
    return fail(PST_E_INVALID, w + ": PST_FLAG_POL_DEGENERATE needs pol_mask = S | P and no matrix output");
  if ((a->flags & PST_FLAG_FP32) && a->M) return fail(PST_E_INVALID, w + ": PST_FLAG_FP32 does not cover the matrix output M");
  if ((a->flags & PST_FLAG_FP32) && a->n_layers - 2 > PST_FP32_MAX_LAYERS)
    return fail(PST_E_INVALID, w + ": PST_FLAG_FP32 is limited to PST_FP32_MAX_LAYERS interior layers (its error grows with the "
                                   "layer count and would pass 1e-5 of full scale); use the fp64 default");
  if ((a->flags & PST_FLAG_SNELL) && (a->flags & (PST_FLAG_FP32 | PST_FLAG_POL_DEGENERATE | PST_FLAG_FORCE_DEEP_KERNEL)))
    return fail(PST_E_INVALID, w + ": PST_FLAG_SNELL excludes the fp32, polarisation-degenerate and deep-kernel flags");
  if ((a->flags & PST_FLAG_NUMERIC_DET) && !a->M) return fail(PST_E_INVALID, w + ": PST_FLAG_NUMERIC_DET needs the matrix output M");
  for (int j = 0; j < a->n_layers; ++j)
    if (a->layer_mat[j] < 0 || a->layer_mat[j] >= a->n_mat) return fail(PST_E_INVALID, w + ": layer_mat entry out of range");
  if ((long long)a->n_wl * a->n_theta > 0x7fffffffLL)
    return fail(PST_E_INVALID, w + ": n_wl * n_theta must stay below 2^31 per call (kernels index points with 32 bits)");
  return PST_OK;
}

// Layer list -> device, with the layer-type analysis of the typed path: interior layers with the same
// (material, thickness) get the same type id; the swept layer is a type of its own.
int sync_layers(pst_handle_t h, const pst_grid_args* a, cudaStream_t st) {
  std::vector<LayerEntry> cur(a->n_layers);
  h->n_types = 0;
  bool overflow = false;
  for (int j = 0; j < a->n_layers; ++j) {
    cur[j] = LayerEntry{a->layer_d[j], a->layer_mat[j], -1};
    if (j == 0 || j == a->n_layers - 1 || overflow) continue;
    int t = -1;
    if (j != a->sweep_layer) {
      for (int u = 0; u < h->n_types; ++u) {
        const int r = h->type_layer[u];
        if (r != a->sweep_layer && cur[r].mat == cur[j].mat && std::memcmp(&cur[r].d, &cur[j].d, sizeof(double)) == 0) {
          t = u;
          break;
        }
      }
    }
    if (t < 0) {
      if (h->n_types == kMaxTypes) {
        overflow = true;
        continue;
      }
      t = h->n_types++;
      h->type_layer[t] = j;
    }
    cur[j].type = t;
  }
  if (overflow) {
    h->n_types = 0;
    for (auto& e : cur) e.type = -1;
  }
  const bool same = cur.size() == h->layers_host.size() &&
                    std::memcmp(cur.data(), h->layers_host.data(), cur.size() * sizeof(LayerEntry)) == 0;
  if (same) return PST_OK;
  // run-length ops over the type sequence in application order (last interior layer first):
  // (A, B, k) = "apply A then B, k times"; (A, -1, k) = "apply A k times"
  std::vector<int4> ops;
  if (h->n_types > 0) {
    std::vector<int> seq;
    for (int j = a->n_layers - 2; j >= 1; --j) seq.push_back(cur[j].type);
    const size_t n = seq.size();
    size_t i = 0;
    auto repeats = [&](size_t at) {  // how often the pair (seq[at], seq[at + 1]) repeats from `at` on
      int k = 1;
      while (at + 2 * (size_t)k + 1 < n && seq[at + 2 * k] == seq[at] && seq[at + 2 * k + 1] == seq[at + 1]) ++k;
      return k;
    };
    while (i < n) {
      if (i + 1 < n) {
        const int ta = seq[i], tb = seq[i + 1];
        const int k = repeats(i);
        if (k == 1 && i + 2 < n && repeats(i + 1) >= 2) {
          // a lone layer in front of a periodic run (the spacer of a cavity): keep the run whole
          ops.push_back(make_int4(ta, -1, 1, 0));
          i += 1;
          continue;
        }
        ops.push_back(make_int4(ta, tb, k, 0));
        i += 2 * (size_t)k;
      } else {
        ops.push_back(make_int4(seq[i], -1, 1, 0));
        i += 1;
      }
    }
  }
  h->ops_host.swap(ops);
  PST_TRY(h->layers_dev.ensure(cur.size()));
  PST_CUDA(cudaMemcpyAsync(h->layers_dev.p, cur.data(), cur.size() * sizeof(LayerEntry), cudaMemcpyHostToDevice, st));
  h->layers_host.swap(cur);
  return PST_OK;
}

// opt table + material flags + trig table for the whole grid (device pointers)
int run_prep(pst_handle_t h, const pst_grid_args* a, const double* wl, const double* theta, const double* cos_theta,
             const double* mat_n, const double* mat_k, cudaStream_t st) {
  const long long nm = (long long)a->n_wl * a->n_mat;
  PST_TRY(h->opt.ensure((size_t)nm * 6));
  PST_TRY(h->trig.ensure((size_t)a->n_theta + kTrigPad));  // padded with copies of the last angle (prep_kernel)
  PST_TRY(h->layers_flagged.ensure(a->n_layers));
  PST_TRY(h->lossy_types.ensure(1));
  // flags and ticket are zero between calls (the kernel's last block clears them); zero them when (re)allocated
  const int* flags_before = h->mat_flags.p;
  PST_TRY(h->mat_flags.ensure((size_t)a->n_mat + 1));  // [cap - 1] is unused; the ticket has its own buffer
  if (h->mat_flags.p != flags_before) PST_CUDA(cudaMemsetAsync(h->mat_flags.p, 0, h->mat_flags.cap * sizeof(int), st));
  const unsigned* ticket_before = h->prep_ticket.p;
  PST_TRY(h->prep_ticket.ensure(1));
  if (h->prep_ticket.p != ticket_before) PST_CUDA(cudaMemsetAsync(h->prep_ticket.p, 0, h->prep_ticket.cap * sizeof(unsigned), st));
  TypeMats tm;
  tm.n = h->n_types;
  for (int t = 0; t < kMaxTypes; ++t) tm.mat[t] = t < h->n_types ? a->layer_mat[h->type_layer[t]] : 0;
  const int n_opt_blocks = (int)((nm + 255) / 256);
  const int n_trig_blocks = (a->n_theta + kTrigPad + 255) / 256;
  prep_kernel<<<(unsigned)(n_opt_blocks + n_trig_blocks), 256, 0, st>>>(
      wl, mat_n, mat_k, a->n_wl, a->n_mat, h->opt.p, h->mat_flags.p, theta, cos_theta, a->n_theta, h->trig.p, n_opt_blocks,
      h->layers_dev.p, a->n_layers, h->layers_flagged.p, tm, h->lossy_types.p, h->prep_ticket.p);
  PST_CUDA(cudaGetLastError());
  h->launches += 1;
  return PST_OK;
}

// plain_out: the stores of a launch need no per-value tests (all three arrays present and local)
inline void set_plain_out(GridParams& prm) {
  prm.plain_out = (prm.R && prm.T && prm.A && prm.n_peers == 0 && !(prm.flags & PST_FLAG_MULTICAST_OUT) && !prm.pol_dup) ? 1 : 0;
}

struct ChainPlan {
  int npol, npol_out, ncol, W, PB, tl_rows, opt_row_bytes, type_stride;
  bool so, sl, seg, typed, row, cavity;
  size_t smem;
  ChainKernel fn;
  TypedKernel fn_typed;
  CavityKernel fn_cavity;
  CavityPlan cp;
};

// Program shape "periodic mirror / spacer / periodic mirror over the same pair of layer types" (a DBR microcavity):
// fills the cavity kernel's plan from the run-length ops of sync_layers.
bool cavity_shape(pst_handle_t h, const pst_grid_args* a, CavityPlan* cp) {
  if (h->ops_host.size() != 3) return false;
  const int4 o0 = h->ops_host[0], o1 = h->ops_host[1], o2 = h->ops_host[2];
  const bool pair_ok = o0.y >= 0 && o0.z >= 3 && o2.y >= 0 && o2.z >= 3 && o0.x != o0.y &&
                       ((o2.x == o0.x && o2.y == o0.y) || (o2.x == o0.y && o2.y == o0.x));
  const bool spacer_ok = o1.y < 0 && o1.z == 1 && o1.x != o0.x && o1.x != o0.y;
  if (!pair_ok || !spacer_ok) return false;
  int sweep_type = -1;
  for (int t = 0; t < h->n_types; ++t)
    if (h->type_layer[t] == a->sweep_layer) sweep_type = t;
  if (sweep_type == o0.x || sweep_type == o0.y) return false;  // a swept mirror layer is not a periodic mirror
  std::memset(cp, 0, sizeof(*cp));
  // staged record layout: see CavityPlan (pst_kernels.cuh); optical row of a material: qr nr inr ni qi ini
  const int jA = h->type_layer[o0.x], jB = h->type_layer[o0.y], jC = h->type_layer[o1.x];
  const int mA = a->layer_mat[jA] * 6, mB = a->layer_mat[jB] * 6, mC = a->layer_mat[jC] * 6;
  const int m_last = a->layer_mat[a->n_layers - 1] * 6, m_first = a->layer_mat[0] * 6;
  const int src[kCavityRow] = {mA + 0, mB + 0, mA + 1, mA + 2, mB + 1, mB + 2, mC + 0, -1,     mC + 1,     mC + 2,      mC + 3, mC + 4,
                               mC + 5, -1,     m_last + 1, m_last + 3,  m_first + 2, m_first + 5, mA + 3, mA + 4, mA + 5, mB + 3, mB + 4, mB + 5};
  for (int i = 0; i < kCavityRow; ++i) cp->src[i] = src[i];
  cp->d[0] = a->layer_d[jA];
  cp->d[1] = a->layer_d[jB];
  cp->d[2] = a->layer_d[jC];
  cp->k1 = o0.z;
  cp->k2 = o2.z;
  cp->hb1 = log2_floor(cp->k1);
  cp->hb2 = log2_floor(cp->k2);
  cp->mirrored = (o2.x == o0.x) ? 0 : 1;
  cp->c_swept = (o1.x == sweep_type) ? 1 : 0;
  cp->lossy_a = 1u << o0.x;
  cp->lossy_ab = (1u << o0.x) | (1u << o0.y);
  cp->lossy_c = 1u << o1.x;
  cp->tiles = 1;
  cp->lossy_types = h->lossy_types.p;
  return true;
}

int plan_chain(pst_handle_t h, const pst_grid_args* a, int n_wl_launch, ChainPlan* pl) {
  pl->npol_out = __builtin_popcount(a->pol_mask);
  pl->npol = (a->flags & PST_FLAG_POL_DEGENERATE) ? 1 : pl->npol_out;  // polarisations the kernel steps
  pl->ncol = a->M ? 2 : 1;
  const int Li = a->n_layers - 2;
  const long long threads1 = (long long)n_wl_launch * a->n_theta * a->n_sweep;
  // A stack that compiles into a short run-length program (<= 8 layer types, <= 24 ops: periodic mirrors of any depth)
  // goes to the typed point kernel however deep it is: its layer matrices are built once per thread and repeated
  // lossless pairs cost two FMAs per period (closed-form power), against one sincos per layer in the layer-parallel
  // kernel -- 10 000 periodic layers on 16 384 points take tens of microseconds instead of 0.65 ms.
  const bool typed_ok = !(a->flags & PST_FLAG_NO_LAYER_CSE) && h->n_types > 0 && Li >= 2 * h->n_types &&
                        (int)h->ops_host.size() <= kMaxOps;
  int W = 1;
  if (a->flags & PST_FLAG_FORCE_DEEP_KERNEL) {
    W = 8;
  } else if (!(a->flags & PST_FLAG_FORCE_POINT_KERNEL) && !typed_ok) {
    const long long want = (long long)h->sm_count * 512;
    if (threads1 < want && Li >= 128) {
      while (W < 16 && threads1 * W < want && Li / (2 * W) >= 32) W *= 2;
    }
  }
  W = std::max(1, std::min(W, std::max(Li, 1)));
  pl->W = W;
  pl->seg = W > 1;
  pl->PB = pl->seg ? 32 : 128;
  if (pl->seg) {
    // deep stacks have few points: pick the points-per-block that balances the blocks over the SMs
    // (148 SMs; e.g. 16 384 points / 32 = 512 blocks would leave SMs with 4 vs 3 blocks, 86 % efficiency)
    const long long pts = (long long)n_wl_launch * a->n_theta;
    double best = -1.0;
    for (int pb : {32, 16, 8}) {
      if (W % (32 / pb) != 0) continue;  // a warp holds 32/pb whole segments (shuffle level of the combine)
      const long long nb = ((pts + pb - 1) / pb) * a->n_sweep;
      const double eff = (double)nb / ((double)h->sm_count * (double)((nb + h->sm_count - 1) / h->sm_count));
      if (eff > best + 0.03) {
        best = eff;
        pl->PB = pb;
      }
    }
  }
  const int nthr = pl->PB * W;
  // typed path (single-segment kernel only): worth it when layer matrices repeat
  pl->typed = !pl->seg && typed_ok;
  pl->type_stride = 0;
  pl->row = false;
  pl->cavity = false;
  pl->fn_typed = nullptr;
  pl->fn_cavity = nullptr;
  if (pl->typed) {
    int units = (h->n_types * (kTypeRecBytes + 4) + 15) / 16;  // records + one exponent int per type
    if (units % 2 == 0) units += 1;
    pl->type_stride = units * 16;
    pl->so = pl->sl = false;
    pl->smem = (size_t)pl->PB * pl->type_stride;
    pl->tl_rows = pl->opt_row_bytes = 0;
    // row mode: one wavelength per block row when 128-wide theta tiles waste < 7 % of the lanes
    const int tiles = (a->n_theta + 127) / 128;
    pl->row = n_wl_launch <= 65535 && a->n_theta >= 120 && tiles * 128 * 100 <= a->n_theta * 107;
    pl->fn = nullptr;
    if (pl->ncol == 1 && !(a->flags & (PST_FLAG_NO_PERIODIC_POWER | PST_FLAG_NO_SHAPE_FASTPATH)) && cavity_shape(h, a, &pl->cp)) {
      // straight-line cavity kernel; in row mode a block walks several theta tiles of its wavelength so that the
      // block-uniform table rows are fetched once -- as many tiles as still leave >= 16 waves of blocks
      pl->cavity = true;
      pl->smem = 0;
      pl->fn_cavity = pick_cavity(pl->npol, pl->row);
      if (!pl->fn_cavity) return fail(PST_E_INVALID, "pst_eval_grid: no cavity kernel variant for this configuration");
      if (pl->row) {
        const long long slots = (long long)h->sm_count * (pl->npol == 2 ? PST_CAV_MINB2 : 8);
        for (int t : {8, 4, 2}) {
          if (t <= tiles && (long long)((tiles + t - 1) / t) * n_wl_launch * a->n_sweep >= 16 * slots) {
            pl->cp.tiles = t;
            break;
          }
        }
      }
      return PST_OK;
    }
    pl->fn_typed = pick_typed(pl->npol, pl->ncol, pl->row);
    if (!pl->fn_typed) return fail(PST_E_INVALID, "pst_eval_grid: no typed kernel variant for this configuration");
    if (pl->smem > 48 * 1024)
      PST_CUDA(cudaFuncSetAttribute((const void*)pl->fn_typed, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl->smem));
    return PST_OK;
  }
  pl->tl_rows = std::min(n_wl_launch, (pl->PB - 1) / a->n_theta + 2);
  pl->opt_row_bytes = a->n_mat * 48 + ((a->n_mat % 2 == 0) ? 16 : 0);
  const size_t budget = std::min<size_t>(h->smem_optin, 200 * 1024);
  const size_t seg_bytes = pl->seg ? (size_t)W * pl->npol * 9 * pl->PB * sizeof(double) : 0;
  const size_t type_bytes = 0;
  (void)nthr;
  const size_t lay_bytes = (size_t)a->n_layers * sizeof(LayerEntry);
  const size_t opt_bytes = (size_t)pl->tl_rows * pl->opt_row_bytes;
  pl->sl = lay_bytes <= 64 * 1024 && seg_bytes + type_bytes + lay_bytes <= budget;
  // deep stacks whose layer list does not fit: every warp streams its segments' slices through a double-buffered window
  // of 32 layers per segment (tmm_chain_kernel, STREAM): 2 x 32 x 16 bytes per segment
  const size_t stream_bytes = (pl->seg && !pl->sl) ? (size_t)W * 1024 : 0;
  size_t used = seg_bytes + type_bytes + stream_bytes + (pl->sl ? lay_bytes : 0);
  pl->so = !(a->flags & PST_FLAG_NO_SMEM_STAGING) && used + opt_bytes <= std::min<size_t>(budget, std::max<size_t>(96 * 1024, used + 16 * 1024));
  used += pl->so ? opt_bytes : 0;
  if (used > budget) return fail(PST_E_INVALID, "pst_eval_grid: shared-memory plan exceeds the device limit");
  pl->smem = used;
  pl->fn = pick_chain(pl->npol, pl->ncol, pl->so, pl->sl, pl->seg);
  if (!pl->fn) return fail(PST_E_INVALID, "pst_eval_grid: no kernel variant for this configuration");
  if (used > 48 * 1024) PST_CUDA(cudaFuncSetAttribute((const void*)pl->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget));
  return PST_OK;
}

// Launch the chain kernel for wavelengths [il0, il0 + n_wl_launch) of the prepared grid.  Output pointers address
// the first element of that wavelength window; a (sweep, pol) plane holds n_wl_launch * n_theta elements.
int launch_chain(pst_handle_t h, const pst_grid_args* a, const ChainPlan& pl, int il0, int n_wl_launch,
                 const double* sweep_d_dev, double* R, double* T, double* A, double* M, cudaStream_t st) {
  GridParams prm;
  std::memset(&prm, 0, sizeof(prm));
  prm.n_wl = a->n_wl;
  prm.n_theta = a->n_theta;
  prm.n_mat = a->n_mat;
  prm.n_layers = a->n_layers;
  prm.sweep_layer = a->sweep_layer;
  prm.pol_first = (a->pol_mask & PST_POL_S) ? 0 : 1;
  prm.tl_rows = pl.tl_rows;
  prm.opt_row_bytes = pl.opt_row_bytes;
  prm.il0 = il0;
  prm.flags = a->flags;
  prm.points = (long long)n_wl_launch * a->n_theta;
  prm.opt = h->opt.p;
  prm.mat_flags = h->mat_flags.p;
  prm.lossy_word = h->lossy_types.p;
  prm.trig = h->trig.p;
  prm.layers = h->layers_flagged.p;  // `type` field = the material's lossy flag
  for (int j = 0; j < a->n_layers; ++j) {
    const double v = std::fabs(a->layer_d[j]);
    if (!(v <= prm.d_max)) prm.d_max = (v == v) ? v : INFINITY;  // a NaN thickness: the kernels keep their range tests
  }
  TypedPlan tp;
  std::memset(&tp, 0, sizeof(tp));
  if (pl.typed) {
    tp.n_types = h->n_types;
    tp.n_ops = (int)h->ops_host.size();
    tp.sweep_type = -1;
    tp.mat_first = a->layer_mat[0];
    tp.mat_last = a->layer_mat[a->n_layers - 1];
    tp.type_stride = pl.type_stride;
    for (int t = 0; t < h->n_types; ++t) {
      const int j = h->type_layer[t];
      tp.type_mat[t] = a->layer_mat[j];
      tp.type_d[t] = a->layer_d[j];
      if (j == a->sweep_layer) tp.sweep_type = t;
    }
    for (int i = 0; i < tp.n_ops; ++i) tp.ops[i] = h->ops_host[i];
    tp.lossy_types = h->lossy_types.p;
  }
  prm.use_bulk = (a->flags & PST_FLAG_NO_BULK_COPY) ? 0 : 1;
  const long long nblk = (prm.points + pl.PB - 1) / pl.PB;
  if (nblk > 0x7fffffffLL) return fail(PST_E_INVALID, "pst_eval_grid: too many blocks");
  prm.pol_dup = (a->flags & PST_FLAG_POL_DEGENERATE) ? 1 : 0;
  const size_t per_sweep = (size_t)pl.npol_out * (size_t)prm.points;
  for (int s0 = 0; s0 < a->n_sweep; s0 += 65535) {
    const int ns = std::min(65535, a->n_sweep - s0);
    prm.sweep_d = sweep_d_dev ? sweep_d_dev + s0 : nullptr;
    prm.R = R ? R + (size_t)s0 * per_sweep : nullptr;
    prm.T = T ? T + (size_t)s0 * per_sweep : nullptr;
    prm.A = A ? A + (size_t)s0 * per_sweep : nullptr;
    prm.M = M ? M + (size_t)s0 * per_sweep * 8 : nullptr;
    prm.n_peers = a->peers ? a->peers->n_peers : 0;
    for (int i = 0; i < prm.n_peers; ++i) {
      prm.peer_R[i] = a->peers->R[i] ? a->peers->R[i] + (size_t)s0 * per_sweep : nullptr;
      prm.peer_T[i] = a->peers->T[i] ? a->peers->T[i] + (size_t)s0 * per_sweep : nullptr;
      prm.peer_A[i] = a->peers->A[i] ? a->peers->A[i] + (size_t)s0 * per_sweep : nullptr;
    }
    set_plain_out(prm);
    dim3 grid((unsigned)nblk, (unsigned)ns), block(pl.PB, pl.W);
    if (pl.typed && pl.row) {
      const int tiles = (a->n_theta + 127) / 128, per = pl.cavity ? pl.cp.tiles : 1;
      grid = dim3((unsigned)((tiles + per - 1) / per), (unsigned)n_wl_launch, (unsigned)ns);
    }
    if (pl.cavity) pl.fn_cavity<<<grid, block, 0, st>>>(prm, pl.cp);
    else if (pl.typed) pl.fn_typed<<<grid, block, pl.smem, st>>>(prm, tp);
    else pl.fn<<<grid, block, pl.smem, st>>>(prm);
    PST_CUDA(cudaGetLastError());
    h->launches += 1;
  }
  return PST_OK;
}

// fp32 mode (PST_FLAG_FP32): one thread per point, both polarisations packed, layer loop without CSE.
int launch_chain_f32(pst_handle_t h, const pst_grid_args* a, int il0, int n_wl_launch, const double* sweep_d_dev,
                     double* R, double* T, double* A, cudaStream_t st) {
  const int PB = 128;
  const int npol = __builtin_popcount(a->pol_mask);
  GridParams prm;
  std::memset(&prm, 0, sizeof(prm));
  prm.n_wl = a->n_wl;
  prm.n_theta = a->n_theta;
  prm.n_mat = a->n_mat;
  prm.n_layers = a->n_layers;
  prm.sweep_layer = a->sweep_layer;
  prm.pol_first = 0;
  prm.pol_mask = a->pol_mask;
  prm.tl_rows = std::min(n_wl_launch, (PB - 1) / a->n_theta + 2);
  prm.opt_row_bytes = a->n_mat * 48 + ((a->n_mat % 2 == 0) ? 16 : 0);
  prm.il0 = il0;
  prm.flags = a->flags;
  prm.points = (long long)n_wl_launch * a->n_theta;
  prm.opt = h->opt.p;
  prm.mat_flags = h->mat_flags.p;
  prm.lossy_word = h->lossy_types.p;
  prm.trig = h->trig.p;
  prm.layers = h->layers_flagged.p;
  prm.use_bulk = (a->flags & PST_FLAG_NO_BULK_COPY) ? 0 : 1;
  const size_t smem = (size_t)prm.tl_rows * prm.opt_row_bytes + (size_t)a->n_layers * sizeof(LayerEntry);
  const bool stage = !(a->flags & PST_FLAG_NO_SMEM_STAGING) && smem <= 64 * 1024;
  auto fn = stage ? tmm_chain_f32_kernel<true> : tmm_chain_f32_kernel<false>;
  if (stage && smem > 48 * 1024) PST_CUDA(cudaFuncSetAttribute((const void*)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  const long long nblk = (prm.points + PB - 1) / PB;
  if (nblk > 0x7fffffffLL) return fail(PST_E_INVALID, "pst_eval_grid: too many blocks");
  const size_t per_sweep = (size_t)npol * (size_t)prm.points;
  for (int s0 = 0; s0 < a->n_sweep; s0 += 65535) {
    const int ns = std::min(65535, a->n_sweep - s0);
    prm.sweep_d = sweep_d_dev ? sweep_d_dev + s0 : nullptr;
    prm.R = R ? R + (size_t)s0 * per_sweep : nullptr;
    prm.T = T ? T + (size_t)s0 * per_sweep : nullptr;
    prm.A = A ? A + (size_t)s0 * per_sweep : nullptr;
    prm.n_peers = a->peers ? a->peers->n_peers : 0;
    for (int i = 0; i < prm.n_peers; ++i) {
      prm.peer_R[i] = a->peers->R[i] ? a->peers->R[i] + (size_t)s0 * per_sweep : nullptr;
      prm.peer_T[i] = a->peers->T[i] ? a->peers->T[i] + (size_t)s0 * per_sweep : nullptr;
      prm.peer_A[i] = a->peers->A[i] ? a->peers->A[i] + (size_t)s0 * per_sweep : nullptr;
    }
    set_plain_out(prm);
    fn<<<dim3((unsigned)nblk, (unsigned)ns), PB, stage ? smem : 0, st>>>(prm);
    PST_CUDA(cudaGetLastError());
    h->launches += 1;
  }
  return PST_OK;
}

// Refraction mode (PST_FLAG_SNELL): one thread per point, plain layer loop with per-layer refracted angles.
int launch_snell(pst_handle_t h, const pst_grid_args* a, int il0, int n_wl_launch, const double* sweep_d_dev, double* R,
                 double* T, double* A, double* M, cudaStream_t st) {
  const int npol = __builtin_popcount(a->pol_mask);
  GridParams prm;
  std::memset(&prm, 0, sizeof(prm));
  prm.n_wl = a->n_wl;
  prm.n_theta = a->n_theta;
  prm.n_mat = a->n_mat;
  prm.n_layers = a->n_layers;
  prm.sweep_layer = a->sweep_layer;
  prm.pol_first = (a->pol_mask & PST_POL_S) ? 0 : 1;
  prm.pol_mask = a->pol_mask;
  prm.il0 = il0;
  prm.flags = a->flags;
  prm.points = (long long)n_wl_launch * a->n_theta;
  prm.opt = h->opt.p;
  prm.mat_flags = h->mat_flags.p;
  prm.lossy_word = h->lossy_types.p;
  prm.trig = h->trig.p;
  prm.layers = h->layers_flagged.p;
  using SnellKernel = void (*)(const GridParams);
  SnellKernel fn = nullptr;
  if (npol == 1) fn = M ? tmm_snell_kernel<1, 2> : tmm_snell_kernel<1, 1>;
  if (npol == 2) fn = M ? tmm_snell_kernel<2, 2> : tmm_snell_kernel<2, 1>;
  if (!fn) return fail(PST_E_INVALID, "pst_eval_grid: PST_FLAG_SNELL needs one or two polarisations");
  const long long nblk = (prm.points + 127) / 128;
  if (nblk > 0x7fffffffLL) return fail(PST_E_INVALID, "pst_eval_grid: too many blocks");
  const size_t per_sweep = (size_t)npol * (size_t)prm.points;
  for (int s0 = 0; s0 < a->n_sweep; s0 += 65535) {
    const int ns = std::min(65535, a->n_sweep - s0);
    prm.sweep_d = sweep_d_dev ? sweep_d_dev + s0 : nullptr;
    prm.R = R ? R + (size_t)s0 * per_sweep : nullptr;
    prm.T = T ? T + (size_t)s0 * per_sweep : nullptr;
    prm.A = A ? A + (size_t)s0 * per_sweep : nullptr;
    prm.M = M ? M + (size_t)s0 * per_sweep * 8 : nullptr;
    prm.n_peers = a->peers ? a->peers->n_peers : 0;
    for (int i = 0; i < prm.n_peers; ++i) {
      prm.peer_R[i] = a->peers->R[i] ? a->peers->R[i] + (size_t)s0 * per_sweep : nullptr;
      prm.peer_T[i] = a->peers->T[i] ? a->peers->T[i] + (size_t)s0 * per_sweep : nullptr;
      prm.peer_A[i] = a->peers->A[i] ? a->peers->A[i] + (size_t)s0 * per_sweep : nullptr;
    }
    set_plain_out(prm);
    fn<<<dim3((unsigned)nblk, (unsigned)ns), 128, 0, st>>>(prm);
    PST_CUDA(cudaGetLastError());
    h->launches += 1;
  }
  return PST_OK;
}

using SweepKernel = void (*)(const GridParams, int, int);
SweepKernel pick_sweep(int npol, bool so, bool sl) {
#define PST_S(P, O, L) if (npol == P && so == O && sl == L) return tmm_sweep_kernel<P, O, L>;
  PST_S(1, true, true) PST_S(1, true, false) PST_S(1, false, true) PST_S(1, false, false)
  PST_S(2, true, true) PST_S(2, true, false) PST_S(2, false, true) PST_S(2, false, false)
#undef PST_S
  return nullptr;
}

bool use_sweep_kernel(const pst_grid_args* a) {
  return a->sweep_layer >= 0 && !a->M && a->n_sweep >= 4 && !(a->flags & PST_FLAG_NO_SWEEP_REUSE) &&
         !(a->flags & PST_FLAG_FORCE_DEEP_KERNEL);
}

// K3 launch for wavelengths [il0, il0 + n_wl_launch): outputs address the window's first element.
int launch_sweep(pst_handle_t h, const pst_grid_args* a, int il0, int n_wl_launch, const double* sweep_d_dev, double* R,
                 double* T, double* A, cudaStream_t st) {
  const int npol = (a->flags & PST_FLAG_POL_DEGENERATE) ? 1 : __builtin_popcount(a->pol_mask);
  const int PB = 128;
  GridParams prm;
  std::memset(&prm, 0, sizeof(prm));
  prm.n_wl = a->n_wl;
  prm.n_theta = a->n_theta;
  prm.n_mat = a->n_mat;
  prm.n_layers = a->n_layers;
  prm.sweep_layer = a->sweep_layer;
  prm.pol_first = (a->pol_mask & PST_POL_S) ? 0 : 1;
  prm.pol_dup = (a->flags & PST_FLAG_POL_DEGENERATE) ? 1 : 0;
  prm.tl_rows = std::min(n_wl_launch, (PB - 1) / a->n_theta + 2);
  prm.opt_row_bytes = a->n_mat * 48 + ((a->n_mat % 2 == 0) ? 16 : 0);
  prm.il0 = il0;
  prm.flags = a->flags;
  prm.points = (long long)n_wl_launch * a->n_theta;
  prm.opt = h->opt.p;
  prm.mat_flags = h->mat_flags.p;
  prm.lossy_word = h->lossy_types.p;
  prm.trig = h->trig.p;
  prm.layers = h->layers_flagged.p;
  prm.use_bulk = (a->flags & PST_FLAG_NO_BULK_COPY) ? 0 : 1;
  prm.sweep_d = sweep_d_dev;
  prm.R = R; prm.T = T; prm.A = A;
  prm.n_peers = a->peers ? a->peers->n_peers : 0;
  for (int i = 0; i < prm.n_peers; ++i) {
    prm.peer_R[i] = a->peers->R[i];
    prm.peer_T[i] = a->peers->T[i];
    prm.peer_A[i] = a->peers->A[i];
  }
  const size_t budget = std::min<size_t>(h->smem_optin, 200 * 1024);
  const size_t lay_bytes = (size_t)a->n_layers * sizeof(LayerEntry);
  const size_t opt_bytes = (size_t)prm.tl_rows * prm.opt_row_bytes;
  const bool sl = lay_bytes <= 32 * 1024;
  const bool so = !(a->flags & PST_FLAG_NO_SMEM_STAGING) && opt_bytes + (sl ? lay_bytes : 0) <= 64 * 1024;
  const size_t smem = (so ? opt_bytes : 0) + (sl ? lay_bytes : 0);
  SweepKernel fn = pick_sweep(npol, so, sl);
  if (!fn) return fail(PST_E_INVALID, "pst_eval_grid: no sweep kernel variant");
  if (smem > 48 * 1024) PST_CUDA(cudaFuncSetAttribute((const void*)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget));
  const long long nblk = (prm.points + PB - 1) / PB;
  if (nblk > 0x7fffffffLL) return fail(PST_E_INVALID, "pst_eval_grid: too many blocks");
  // sweeps per block: amortise the prefix/suffix work, but keep at least ~8 blocks per SM in flight
  int spb = 64;
  while (spb > 8 && nblk * ((a->n_sweep + spb - 1) / spb) < (long long)h->sm_count * 8) spb /= 2;
  const long long gy = (a->n_sweep + spb - 1) / spb;
  if (gy > 65535) spb = (int)((a->n_sweep + 65534) / 65535);
  dim3 grid((unsigned)nblk, (unsigned)((a->n_sweep + spb - 1) / spb));
  fn<<<grid, PB, smem, st>>>(prm, a->n_sweep, spb);
  PST_CUDA(cudaGetLastError());
  h->launches += 1;
  return PST_OK;
}

}  // namespace

extern "C" int pst_eval_grid(pst_handle_t h, const pst_grid_args* a, void* stream) {
  if (!h) return fail(PST_E_INVALID, "pst_eval_grid: NULL handle");
  PST_TRY(validate_grid(a, "pst_eval_grid"));
  DeviceGuard g(h->device);
  if (!g.ok) return fail(PST_E_CUDA, "pst_eval_grid: cudaSetDevice failed");
  cudaStream_t st = (cudaStream_t)stream;
  PST_TRY(scratch_acquire(h, st));
  PST_TRY(sync_layers(h, a, st));
  PST_TRY(run_prep(h, a, a->wl, a->theta, a->cos_theta, a->mat_n, a->mat_k, st));
  if (a->flags & PST_FLAG_SNELL) {
    PST_TRY(launch_snell(h, a, 0, a->n_wl, a->sweep_d, a->R, a->T, a->A, a->M, st));
  } else if (a->flags & PST_FLAG_FP32) {
    PST_TRY(launch_chain_f32(h, a, 0, a->n_wl, a->sweep_d, a->R, a->T, a->A, st));
  } else if (use_sweep_kernel(a)) {
    PST_TRY(launch_sweep(h, a, 0, a->n_wl, a->sweep_d, a->R, a->T, a->A, st));
  } else {
    ChainPlan pl;
    PST_TRY(plan_chain(h, a, a->n_wl, &pl));
    PST_TRY(launch_chain(h, a, pl, 0, a->n_wl, a->sweep_d, a->R, a->T, a->A, a->M, st));
  }
  return scratch_release(h, st);
}

extern "C" int pst_eval_grid_host(pst_handle_t h, const pst_grid_args* a) {
  if (!h) return fail(PST_E_INVALID, "pst_eval_grid_host: NULL handle");
  PST_TRY(validate_grid(a, "pst_eval_grid_host"));
  if (a->flags & PST_FLAG_MULTICAST_OUT) return fail(PST_E_INVALID, "pst_eval_grid_host: PST_FLAG_MULTICAST_OUT needs device (multicast) output pointers");
  if (a->peers) return fail(PST_E_INVALID, "pst_eval_grid_host: peer outputs are device pointers (pst_eval_grid only)");
  DeviceGuard g(h->device);
  if (!g.ok) return fail(PST_E_CUDA, "pst_eval_grid_host: cudaSetDevice failed");
  if (!h->s_compute) {
    PST_CUDA(cudaStreamCreateWithFlags(&h->s_compute, cudaStreamNonBlocking));
    PST_CUDA(cudaStreamCreateWithFlags(&h->s_copy, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      PST_CUDA(cudaEventCreateWithFlags(&h->ev_done[i], cudaEventDisableTiming));
      PST_CUDA(cudaEventCreateWithFlags(&h->ev_copied[i], cudaEventDisableTiming));
    }
  }
  cudaStream_t sc = h->s_compute, sd = h->s_copy;
  PST_TRY(scratch_acquire(h, sc));  // work of an earlier pst_eval_grid on the caller's stream may still read the tables
  const int npol = __builtin_popcount(a->pol_mask);
  const size_t nwl = a->n_wl, nth = a->n_theta, nm = a->n_mat, nsw = a->n_sweep;
  // ---- inputs: one device block, host -> device on the compute stream
  const size_t in_elems = nwl + 2 * nth + 2 * nm * nwl + nsw;
  PST_TRY(h->hp_in.ensure(in_elems));
  double* d_wl = h->hp_in.p;
  double* d_th = d_wl + nwl;
  double* d_ct = d_th + nth;
  double* d_n = d_ct + nth;
  double* d_k = d_n + nm * nwl;
  double* d_sw = d_k + nm * nwl;
  PST_CUDA(cudaMemcpyAsync(d_wl, a->wl, nwl * 8, cudaMemcpyHostToDevice, sc));
  if (a->theta) PST_CUDA(cudaMemcpyAsync(d_th, a->theta, nth * 8, cudaMemcpyHostToDevice, sc));
  if (a->cos_theta) PST_CUDA(cudaMemcpyAsync(d_ct, a->cos_theta, nth * 8, cudaMemcpyHostToDevice, sc));
  PST_CUDA(cudaMemcpyAsync(d_n, a->mat_n, nm * nwl * 8, cudaMemcpyHostToDevice, sc));
  PST_CUDA(cudaMemcpyAsync(d_k, a->mat_k, nm * nwl * 8, cudaMemcpyHostToDevice, sc));
  if (a->sweep_d) PST_CUDA(cudaMemcpyAsync(d_sw, a->sweep_d, nsw * 8, cudaMemcpyHostToDevice, sc));
  PST_TRY(sync_layers(h, a, sc));
  PST_TRY(run_prep(h, a, d_wl, a->theta ? d_th : nullptr, a->cos_theta ? d_ct : nullptr, d_n, d_k, sc));

  // ---- wavelength chunks: kernel on s_compute, device -> host on s_copy, two output buffers in flight
  const int n_out = (a->R ? 1 : 0) + (a->T ? 1 : 0) + (a->A ? 1 : 0) + (a->M ? 8 : 0);
  const size_t row_elems = (size_t)nsw * npol * nth;               // result elements per wavelength and real output
  const size_t target = (size_t)48 << 20;                          // ~48 MiB of results per chunk
  size_t chunk = std::max<size_t>(1, target / std::max<size_t>(1, row_elems * n_out * 8));
  chunk = std::min(chunk, nwl);
  if (chunk < nwl) chunk = std::max<size_t>(1, (nwl + ((nwl + chunk - 1) / chunk) - 1) / ((nwl + chunk - 1) / chunk));
  for (int b = 0; b < 2; ++b) PST_TRY(h->hp_out[b].ensure(row_elems * n_out * chunk));
  int nchunk = 0;
  for (size_t l0 = 0; l0 < nwl; l0 += chunk, ++nchunk) {
    const int b = nchunk & 1;
    const size_t cw = std::min(chunk, nwl - l0);
    const size_t plane_c = cw * nth;           // chunk-local (sweep, pol) plane
    const size_t plane_f = nwl * nth;          // plane in the caller's arrays
    const size_t planes = nsw * npol;
    double* base = h->hp_out[b].p;
    double* dR = a->R ? base : nullptr; if (a->R) base += planes * plane_c;
    double* dT = a->T ? base : nullptr; if (a->T) base += planes * plane_c;
    double* dA = a->A ? base : nullptr; if (a->A) base += planes * plane_c;
    double* dM = a->M ? base : nullptr;
    if (nchunk >= 2) PST_CUDA(cudaStreamWaitEvent(sc, h->ev_copied[b], 0));
    ChainPlan pl;
    pst_grid_args loc = *a;
    loc.R = dR; loc.T = dT; loc.A = dA; loc.M = dM;
    if (loc.flags & PST_FLAG_SNELL) {
      PST_TRY(launch_snell(h, &loc, (int)l0, (int)cw, a->sweep_d ? d_sw : nullptr, dR, dT, dA, dM, sc));
    } else if (loc.flags & PST_FLAG_FP32) {
      PST_TRY(launch_chain_f32(h, &loc, (int)l0, (int)cw, a->sweep_d ? d_sw : nullptr, dR, dT, dA, sc));
    } else if (use_sweep_kernel(&loc)) {
      PST_TRY(launch_sweep(h, &loc, (int)l0, (int)cw, d_sw, dR, dT, dA, sc));
    } else {
      PST_TRY(plan_chain(h, &loc, (int)cw, &pl));
      PST_TRY(launch_chain(h, &loc, pl, (int)l0, (int)cw, a->sweep_d ? d_sw : nullptr, dR, dT, dA, dM, sc));
    }
    PST_CUDA(cudaEventRecord(h->ev_done[b], sc));
    PST_CUDA(cudaStreamWaitEvent(sd, h->ev_done[b], 0));
    auto copy_back = [&](double* host, const double* dev, size_t width) -> int {
      // planes rows of `width*cw*nth` doubles; the caller's rows are plane_f*width apart
      PST_CUDA(cudaMemcpy2DAsync(host + l0 * nth * width, plane_f * width * 8, dev, plane_c * width * 8,
                                 plane_c * width * 8, planes, cudaMemcpyDeviceToHost, sd));
      return PST_OK;
    };
    if (a->R) PST_TRY(copy_back(a->R, dR, 1));
    if (a->T) PST_TRY(copy_back(a->T, dT, 1));
    if (a->A) PST_TRY(copy_back(a->A, dA, 1));
    if (a->M) PST_TRY(copy_back(a->M, dM, 8));
    PST_CUDA(cudaEventRecord(h->ev_copied[b], sd));
  }
  PST_CUDA(cudaStreamSynchronize(sd));
  PST_CUDA(cudaStreamSynchronize(sc));
  h->scratch_busy = false;  // everything this handle enqueued on its internal streams has completed
  return PST_OK;
}

// ------------------------------------------------------------------------------------------------ API-parity kernels
extern "C" int pst_layer_matrices(pst_handle_t h, const double* n_re, const double* n_im, const double* wl, int n_wl,
                                  double theta, double cos_theta, unsigned pol, double thickness, double* out,
                                  void* stream) {
  (void)theta;
  if (!h) return fail(PST_E_INVALID, "pst_layer_matrices: NULL handle");
  if (!n_re || !n_im || !wl || !out || n_wl <= 0) return fail(PST_E_INVALID, "pst_layer_matrices: bad argument");
  if (pol != PST_POL_S && pol != PST_POL_P)
    return fail(PST_E_INVALID, "pst_layer_matrices: Must specify 's-wave' or 'p-wave'.");
  DeviceGuard g(h->device);
  layer_matrices_kernel<<<(n_wl + 127) / 128, 128, 0, (cudaStream_t)stream>>>(n_re, n_im, wl, n_wl, cos_theta, pol,
                                                                             thickness, out);
  PST_CUDA(cudaGetLastError());
  h->launches += 1;
  return PST_OK;
}

extern "C" int pst_propagation_matrices(pst_handle_t h, const double* kx, int n, double d_re, double d_im, double* out,
                                        void* stream) {
  if (!h) return fail(PST_E_INVALID, "pst_propagation_matrices: NULL handle");
  if (!kx || !out || n <= 0) return fail(PST_E_INVALID, "pst_propagation_matrices: bad argument");
  DeviceGuard g(h->device);
  propagation_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(kx, n, d_re, d_im, out);
  PST_CUDA(cudaGetLastError());
  h->launches += 1;
  return PST_OK;
}

extern "C" int pst_wavenumbers(pst_handle_t h, const double* n_re, const double* n_im, const double* wl, int n_wl,
                               double cos_theta, double sin_theta, double* kx, double* kz, void* stream) {
  if (!h) return fail(PST_E_INVALID, "pst_wavenumbers: NULL handle");
  if (!n_re || !n_im || !wl || !kx || !kz || n_wl <= 0) return fail(PST_E_INVALID, "pst_wavenumbers: bad argument");
  DeviceGuard g(h->device);
  wavenumbers_kernel<<<(n_wl + 127) / 128, 128, 0, (cudaStream_t)stream>>>(n_re, n_im, wl, n_wl, cos_theta, sin_theta,
                                                                          kx, kz);
  PST_CUDA(cudaGetLastError());
  h->launches += 1;
  return PST_OK;
}

extern "C" int pst_t_r_from_matrices(pst_handle_t h, const double* M, int64_t n, double* T, double* R, void* stream) {
  if (!h) return fail(PST_E_INVALID, "pst_t_r_from_matrices: NULL handle");
  if (!M || !T || !R || n <= 0) return fail(PST_E_INVALID, "pst_t_r_from_matrices: bad argument");
  DeviceGuard g(h->device);
  t_r_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(M, n, T, R);
  PST_CUDA(cudaGetLastError());
  h->launches += 1;
  return PST_OK;
}

extern "C" int pst_chain_matrices(pst_handle_t h, const double* first_inv, const double* interior, const double* last_d,
                                  int n_interior, int n_wl, double* out, void* stream) {
  if (!h) return fail(PST_E_INVALID, "pst_chain_matrices: NULL handle");
  if (!first_inv || !last_d || !out || n_wl <= 0 || n_interior < 0 || (n_interior > 0 && !interior))
    return fail(PST_E_INVALID, "pst_chain_matrices: bad argument");
  DeviceGuard g(h->device);
  chain_matrices_kernel<<<(n_wl + 127) / 128, 128, 0, (cudaStream_t)stream>>>(first_inv, interior, last_d, n_interior,
                                                                             n_wl, out);
  PST_CUDA(cudaGetLastError());
  h->launches += 1;
  return PST_OK;
}

extern "C" int pst_fresnel(pst_handle_t h, const double* n1, const double* n2, const double* k1x, const double* k2x,
                           int64_t n, double* out_rt, double* out_D, void* stream) {
  if (!h) return fail(PST_E_INVALID, "pst_fresnel: NULL handle");
  if (!n1 || !n2 || !k1x || !k2x || !out_rt || n <= 0) return fail(PST_E_INVALID, "pst_fresnel: bad argument");
  DeviceGuard g(h->device);
  fresnel_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(n1, n2, k1x, k2x, n, out_rt, out_D);
  PST_CUDA(cudaGetLastError());
  h->launches += 1;
  return PST_OK;
}

extern "C" int pst_interface_matrices(pst_handle_t h, const double* r, const double* t, int64_t n, double* out,
                                      void* stream) {
  if (!h) return fail(PST_E_INVALID, "pst_interface_matrices: NULL handle");
  if (!r || !t || !out || n <= 0) return fail(PST_E_INVALID, "pst_interface_matrices: bad argument");
  DeviceGuard g(h->device);
  interface_matrix_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(r, t, n, out);
  PST_CUDA(cudaGetLastError());
  h->launches += 1;
  return PST_OK;
}

extern "C" int pst_field_amplitudes(pst_handle_t h, const pst_grid_args* a, double* amp, void* stream) {
  if (!h) return fail(PST_E_INVALID, "pst_field_amplitudes: NULL handle");
  if (!amp) return fail(PST_E_INVALID, "pst_field_amplitudes: NULL output");
  pst_grid_args loc;
  if (a) {
    loc = *a;
    loc.R = amp;  // validate_grid wants some output
    loc.T = loc.A = loc.M = nullptr;
    loc.peers = nullptr;
  }
  PST_TRY(validate_grid(a ? &loc : nullptr, "pst_field_amplitudes"));
  if (a->sweep_layer >= 0) return fail(PST_E_INVALID, "pst_field_amplitudes: thickness sweeps are not covered");
  if (a->flags & ~(unsigned)PST_FLAG_SNELL) return fail(PST_E_INVALID, "pst_field_amplitudes: only PST_FLAG_SNELL is accepted");
  DeviceGuard g(h->device);
  if (!g.ok) return fail(PST_E_CUDA, "pst_field_amplitudes: cudaSetDevice failed");
  cudaStream_t st = (cudaStream_t)stream;
  PST_TRY(scratch_acquire(h, st));
  PST_TRY(sync_layers(h, &loc, st));
  PST_TRY(run_prep(h, &loc, a->wl, a->theta, a->cos_theta, a->mat_n, a->mat_k, st));
  GridParams prm;
  std::memset(&prm, 0, sizeof(prm));
  prm.n_wl = a->n_wl;
  prm.n_theta = a->n_theta;
  prm.n_mat = a->n_mat;
  prm.n_layers = a->n_layers;
  prm.sweep_layer = -1;
  prm.flags = a->flags;
  prm.points = (long long)a->n_wl * a->n_theta;
  prm.opt = h->opt.p;
  prm.trig = h->trig.p;
  prm.layers = h->layers_flagged.p;  // `type` field = the material's lossy flag
  const int npol = __builtin_popcount(a->pol_mask);
  const int pol_a = (a->pol_mask & PST_POL_S) ? 0 : 1, pol_b = 1;
  const dim3 grid((unsigned)((prm.points + 127) / 128), (unsigned)npol);
  if (a->flags & PST_FLAG_SNELL) field_amplitudes_kernel<true><<<grid, 128, 0, st>>>(prm, pol_a, pol_b, amp);
  else field_amplitudes_kernel<false><<<grid, 128, 0, st>>>(prm, pol_a, pol_b, amp);
  PST_CUDA(cudaGetLastError());
  h->launches += 1;
  return scratch_release(h, st);
}

extern "C" int pst_find_peaks(pst_handle_t h, const double* y, int n_outer, int n_pts, int n_inner, double height,
                              int use_height, int distance, int max_peaks, int reverse, const double* axis, int* out_idx,
                              double* out_height, int* out_count, double* out_spacing, void* stream) {
  if (!h) return fail(PST_E_INVALID, "pst_find_peaks: NULL handle");
  if (!y || !out_idx || !out_count || n_outer <= 0 || n_pts <= 0 || n_inner <= 0)
    return fail(PST_E_INVALID, "pst_find_peaks: bad argument");
  if (max_peaks < 1 || max_peaks > kPeakCap) return fail(PST_E_INVALID, "pst_find_peaks: max_peaks must be in 1..128");
  if (distance < 0) return fail(PST_E_INVALID, "pst_find_peaks: `distance` must be greater or equal to 1");
  DeviceGuard g(h->device);
  const long long n_spec = (long long)n_outer * n_inner;
  find_peaks_kernel<<<(unsigned)((n_spec + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
      y, n_outer, n_pts, n_inner, height, use_height, distance, max_peaks, reverse, axis, out_idx, out_height, out_count,
      out_spacing);
  PST_CUDA(cudaGetLastError());
  h->launches += 1;
  return PST_OK;
}

extern "C" int pst_polariton_branches(pst_handle_t h, const double* y, int n_outer, int n_pts, int n_inner, int reverse,
                                      const double* axis, const int* peak_idx, const double* peak_height,
                                      const int* peak_count, int max_peaks, double* out, void* stream) {
  if (!h) return fail(PST_E_INVALID, "pst_polariton_branches: NULL handle");
  if (!y || !axis || !peak_idx || !peak_height || !peak_count || !out || n_outer <= 0 || n_pts < 3 || n_inner <= 0 ||
      max_peaks < 1)
    return fail(PST_E_INVALID, "pst_polariton_branches: bad argument");
  DeviceGuard g(h->device);
  const long long n_spec = (long long)n_outer * n_inner;
  polariton_branches_kernel<<<(unsigned)((n_spec + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
      y, n_outer, n_pts, n_inner, reverse, axis, peak_idx, peak_height, peak_count, max_peaks, out);
  PST_CUDA(cudaGetLastError());
  h->launches += 1;
  return PST_OK;
}

extern "C" int pst_fit_two_lorentzians(pst_handle_t h, const double* y, int n_outer, int n_pts, int n_inner, int reverse,
                                       const double* axis, const double* seed, int i_lo, int i_hi, double sigma0, double tol,
                                       int max_iter, double* out, void* stream) {
  if (!h) return fail(PST_E_INVALID, "pst_fit_two_lorentzians: NULL handle");
  if (!y || !axis || !seed || !out || n_outer <= 0 || n_pts < 6 || n_inner <= 0)
    return fail(PST_E_INVALID, "pst_fit_two_lorentzians: bad argument");
  if (i_lo < 0 || i_hi > n_pts || i_hi - i_lo < 6)
    return fail(PST_E_INVALID, "pst_fit_two_lorentzians: the window [i_lo, i_hi) needs at least 6 samples inside the spectrum");
  if (max_iter < 2 || !(sigma0 > 0.0) || !(tol >= 0.0))
    return fail(PST_E_INVALID, "pst_fit_two_lorentzians: max_iter >= 2, sigma0 > 0 and tol >= 0");
  DeviceGuard g(h->device);
  const long long n_spec = (long long)n_outer * n_inner;
  fit_two_lorentzians_kernel<<<(unsigned)((n_spec + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
      y, n_outer, n_pts, n_inner, reverse, axis, seed, i_lo, i_hi, sigma0, tol, max_iter, out);
  PST_CUDA(cudaGetLastError());
  h->launches += 1;
  return PST_OK;
}

// ------------------------------------------------------------------------------------------------ measurement
extern "C" int pst_fp64_peak_probe(pst_handle_t h, int iters, double* tflops_out, void* stream) {
  if (!h || !tflops_out || iters <= 0) return fail(PST_E_INVALID, "pst_fp64_peak_probe: bad argument");
  DeviceGuard g(h->device);
  cudaStream_t st = (cudaStream_t)stream;
  PST_TRY(h->sink.ensure(1));
  const int blocks = h->sm_count * 8, threads = 256;
  cudaEvent_t e0, e1;
  PST_CUDA(cudaEventCreate(&e0));
  PST_CUDA(cudaEventCreate(&e1));
  dfma_probe_kernel<<<blocks, threads, 0, st>>>(iters / 8 + 1, 1.0, h->sink.p);  // warm-up
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    PST_CUDA(cudaEventRecord(e0, st));
    dfma_probe_kernel<<<blocks, threads, 0, st>>>(iters, 1.0, h->sink.p);
    PST_CUDA(cudaEventRecord(e1, st));
    PST_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    PST_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    const double flop = 2.0 * 64.0 * (double)iters * (double)blocks * threads;  // 8 chains x 8 unrolled FMAs x 2 flop
    best = std::max(best, flop / (ms * 1e-3) / 1e12);
  }
  PST_CUDA(cudaGetLastError());
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  h->launches += 6;
  *tflops_out = best;
  return PST_OK;
}

extern "C" int pst_flush_l2(pst_handle_t h, void* stream) {
  if (!h) return fail(PST_E_INVALID, "pst_flush_l2: NULL handle");
  DeviceGuard g(h->device);
  const size_t bytes = (size_t)256 << 20;  // 2x the 126 MB L2
  PST_TRY(h->l2buf.ensure(bytes));
  PST_CUDA(cudaMemsetAsync(h->l2buf.p, 0, bytes, (cudaStream_t)stream));
  return PST_OK;
}
