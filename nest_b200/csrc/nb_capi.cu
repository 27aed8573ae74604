// nb_capi.cu -- the C-ABI of include/nest_b200.h over the sm_100a kernels.
// No CPU fallback: every compute entry point needs a CUDA device and fails with
// NB200_ENODEV / NB200_ECUDA otherwise.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <chrono>
#include <cfloat>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <string>
#include <vector>

#include "nb_host.h"
#include "nb_fit.h"
#include "nb_kernels.cuh"

using namespace nb;

struct nb200_ctx {
  int device = 0;
  int sm_count = 148;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  std::string err;
  uint64_t launches = 0;
  void* d_fit_scratch = nullptr;  // results of k_band_fit on their way to the host (grow-only)
  size_t fit_scratch_bytes = 0;
  int* d_err = nullptr;  // [0] error flags of the kernels, [1] k_hits group counter, [2] deferred-queue fill
  SlotEntry* d_hitq = nullptr;  // k_hits -> k_hits_finish queue (nb_kernels.cuh)
  unsigned int hitq_cap = 0;
  // device-resident band accumulator (async API)
  unsigned long long* d_band = nullptr;
  bool band_external = false;
  size_t band_words = 0;
  int band_numBins = 0, band_logBins = 0;
  uint64_t band_events = 0;
  // scratch accumulator for the synchronous nb200_run
  unsigned long long* d_band_tmp = nullptr;
  size_t band_tmp_words = 0;
  // cached device copies of map LUTs / WIMP tables / drift table
  struct Blob {
    uint64_t hash;
    size_t bytes;
    void* dev;
  };
  std::vector<Blob> blobs;
  // tables already resolved during the current API call (host pointer -> device copy): a sweep's points share
  // their maps, which are then hashed once per call instead of once per point
  struct Memo {
    const void* host;
    size_t bytes;
    void* dev;
  };
  std::vector<Memo> call_memo;
  // double-buffered device staging + copy stream for nb200_run's host outputs (kept across calls)
  void* d_stage[2] = {nullptr, nullptr};
  size_t stage_bytes = 0;
  void* d_list = nullptr;
  size_t list_bytes = 0;
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_done[2] = {nullptr, nullptr}, ev_copied[2] = {nullptr, nullptr};
  // optional per-stage timing
  bool timing = false;
  std::vector<cudaEvent_t> tev;   // 4 events per chunk launched since the last read
  // k_pre -> k_hits -> k_post workspace (one chunk)
  void* d_work = nullptr;
  uint64_t work_cap = 0;
  int chunk_log2 = 24;  // chunk_events(); NB200_CHUNK_LOG2 overrides
  int occ_pre = 0, occ_hits = 0, occ_post = 0;
  std::vector<double> vtable_host;
  nb200_detector vtable_det;
  double vtable_zstep = 0.;
  double* d_vtable = nullptr;
  // alias tables of Binomial(2^b, P_dphe), one per distinct P_dphe seen (a sweep may mix detectors)
  struct Dpe {
    double p;
    void* dev;
  };
  std::vector<Dpe> dpe;
  // nb200_run_sweep: per-point RunParams and band slabs
  void* d_sweep_params = nullptr;
  size_t sweep_params_bytes = 0;
  unsigned long long* d_sweep_bands = nullptr;
  size_t sweep_band_words = 0;
  unsigned long long* h_sweep_bands = nullptr;  // pinned staging of the sweep's bands
  size_t h_sweep_words = 0;
  int occ_pre_sweep = 0, occ_hits_sweep = 0;
  size_t sweep_last_points = 0, sweep_last_words = 0;  // shape of the bands the last nb200_run_sweep left on the device
};

namespace {

thread_local std::string g_err_noctx;

// NB200_TRACE=1: wall-clock phases of the synchronous entry points on stderr
struct Trace {
  bool on;
  std::chrono::steady_clock::time_point t0;
  const char* what;
  explicit Trace(const char* w) : on(std::getenv("NB200_TRACE") != nullptr), what(w) {
    if (on) t0 = std::chrono::steady_clock::now();
  }
  void mark(const char* phase) {
    if (!on) return;
    auto t1 = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[nb200 trace] %s: %s %.3f ms\n", what, phase,
                 std::chrono::duration<double, std::milli>(t1 - t0).count());
    t0 = t1;
  }
};

int fail(nb200_ctx* c, int code, const std::string& msg) {
  if (c)
    c->err = msg;
  else
    g_err_noctx = msg;
  return code;
}
#define NB_CUDA(c, call)                                                                   \
  do {                                                                                     \
    cudaError_t e_ = (call);                                                               \
    if (e_ != cudaSuccess)                                                                 \
      return fail(c, NB200_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_));     \
  } while (0)

// fixed-point scale of the band's sum of X (S1, or S2 for useS2 == 2): 2^-20 units while X stays below 1024,
// coarser for larger X so that X * scale < 2^30 and a bin holds 2^33 events before its 64-bit sum could wrap
double band_scale_x(const nb200_analysis* a) {
  const double mx = a->useS2 == 2 ? fabs(a->maxS2) : fabs(a->maxS1);
  if (!(mx > 1024.)) return NB200_FX_S1;
  const int e = int(ceil(log2(mx)));
  return ldexp(1., std::max(0, 30 - e));
}

size_t band_word_count(const nb200_analysis* a) {
  return size_t(a->numBins) * BAND_WORDS + size_t(a->numBins) * size_t(a->logBins > 0 ? a->logBins : 0);
}

// FNV-1a over the table's bytes
uint64_t blob_hash(const void* p, size_t n) {
  const unsigned char* b = static_cast<const unsigned char*>(p);
  uint64_t h = 1469598103934665603ull;
  for (size_t i = 0; i < n; ++i) h = (h ^ b[i]) * 1099511628211ull;
  return h;
}
// upload (or reuse) a host table; keyed on size + content hash: a table is copied once per context, never
// rewritten while kernels of an earlier call may still read it, whatever host address it comes from
int upload_blob(nb200_ctx* c, const void* host, size_t bytes, void** dev) {
  for (auto& m : c->call_memo)
    if (m.host == host && m.bytes == bytes) {
      *dev = m.dev;
      return 0;
    }
  const uint64_t h = blob_hash(host, bytes);
  for (auto& b : c->blobs)
    if (b.bytes == bytes && b.hash == h) {
      *dev = b.dev;
      c->call_memo.push_back({host, bytes, b.dev});
      return 0;
    }
  if (c->blobs.size() >= 256) {  // bounded: drop everything not in use by queued work
    NB_CUDA(c, cudaStreamSynchronize(c->stream));
    for (auto& b : c->blobs) cudaFree(b.dev);
    c->blobs.clear();
  }
  void* d = nullptr;
  NB_CUDA(c, cudaMalloc(&d, bytes));
  NB_CUDA(c, cudaMemcpy(d, host, bytes, cudaMemcpyHostToDevice));  // synchronous: the host table may be a temporary
  c->blobs.push_back({h, bytes, d});
  c->call_memo.push_back({host, bytes, d});
  *dev = d;
  return 0;
}

int upload_map(nb200_ctx* c, nb200_map* m) {
  if (m->kind != NB200_MAP_LUT_RZ) {
    m->lut = nullptr;
    return 0;
  }
  if (!m->lut || m->n0 < 1 || m->n1 < 1) return fail(c, NB200_EINVAL, "LUT map without table");
  void* d = nullptr;
  int rc = upload_blob(c, m->lut, sizeof(double) * size_t(m->n0) * size_t(m->n1), &d);
  if (rc) return rc;
  m->lut = static_cast<const double*>(d);
  return 0;
}

int validate(nb200_ctx* c, const nb200_detector* det, const nb200_model* mo, const nb200_analysis* ana,
             const nb200_source* src) {
  if (!det || !mo || !ana || !src) return fail(c, NB200_EINVAL, "null argument");
  // execNEST.cpp:419-426
  if (det->TopDrift <= 0. || det->anode <= 0. || det->gate <= 0.)
    return fail(c, NB200_EINVAL, "ERROR, unphysical value(s) of position within the detector geometry.");
  if (mo->n_widths < 13)  // NEST.cpp:255-259 (entries beyond the 17 that are read are ignored, as in the reference)
    return fail(c, NB200_EPHYSICS, "ERROR: You need a minimum of 13 free parameters for the resolution model.");
  if (ana->numBins < 1 || ana->numBins > 1000)  // NUMBINS_MAX TestSpectra.hh:34, execNEST.cpp:1510-1516
    return fail(c, NB200_EINVAL, "ERROR: Too many bins. Decrease numBins (analysis.hh) or increase NUMBINS_MAX");
  if (ana->s1mode < NB200_S1_FULL || ana->s1mode > NB200_S1_HYBRID || ana->s2mode != NB200_S2_FULL)
    return fail(c, NB200_EINVAL, "Waveform S1/S2 calculation modes are not part of this path");
  if (src->species < 0 || src->species > NB200_atmNu)
    return fail(c, NB200_EINVAL, "species not supported by the GPU chain");
  if (src->inField < 0. && src->inField != -1. && !src->vec_mode)  // execNEST.cpp:863-868
    return fail(c, NB200_EINVAL, "ERROR: Neg field is not permitted. We don't simulate field dir (yet). Put in magnitude.");
  if (det->E_gas < 0.) return fail(c, NB200_EINVAL, "ERROR: Neg field is not permitted.");
  if (src->kind == NB200_SRC_LIST && !src->E) return fail(c, NB200_EINVAL, "LIST source without energies");
  if (src->kind == NB200_SRC_WIMP && (!src->wimp_base || !src->wimp_exponent))
    return fail(c, NB200_EINVAL, "WIMP source without nb200_wimp_prep table");
  // RecombOmegaER's throw (NEST.cpp:183-184) is reached by every ER-like event
  const bool erlike = src->species >= NB200_gammaRay || mo->er_model >= 0;
  if (erlike && mo->NRERWidthsParam[9] > mo->NRERWidthsParam[8])
    return fail(c, NB200_EPHYSICS,
                "Low-E r-fluct Gauss peak's width greater than centroid. This is physically wrong so "
                "please check your inputs, in NRERWidthsParam: they must be in the old wrong order.");
  return 0;
}

int check_device_error(nb200_ctx* c);

// WIMP source: the spectrum table to device memory, WIMP_dRate(0) per isotope (TestSpectra.cpp:459, 466)
// device copy of the alias tables of Binomial(2^b, p) (binomial_fixed_p), built once per distinct p and context
static int alias_table_for(nb200_ctx* c, double p, const uint2** out) {
  void* dtab = nullptr;
  for (auto& e : c->dpe)
    if (e.p == p) dtab = e.dev;
  if (!dtab) {
    std::vector<AliasEntry> tab(NB_ALIAS_ENTRIES);
    build_binomial_alias(p, tab.data());
    NB_CUDA(c, cudaMalloc(&dtab, sizeof(AliasEntry) * tab.size()));
    NB_CUDA(c, cudaMemcpy(dtab, tab.data(), sizeof(AliasEntry) * tab.size(), cudaMemcpyHostToDevice));
    c->dpe.push_back({p, dtab});
  }
  *out = static_cast<const uint2*>(dtab);
  return 0;
}

int upload_wimp_source(nb200_ctx* c, const nb200_source* src, RunParams* P) {
  if (src->kind != NB200_SRC_WIMP) return 0;
  if (!src->wimp_base || !src->wimp_exponent) return fail(c, NB200_EINVAL, "WIMP source without nb200_wimp_prep table");
  void* d = nullptr;
  int r;
  if ((r = upload_blob(c, src->wimp_base, sizeof(double) * 100, &d))) return r;
  P->src.wimp_base = static_cast<const double*>(d);
  if ((r = upload_blob(c, src->wimp_exponent, sizeof(double) * 100, &d))) return r;
  P->src.wimp_exponent = static_cast<const double*>(d);
  for (int k = 0; k < 9; ++k) {
    bool ok;
    P->wimp_dr0[k] = host::wimp_drate(0., src->wimp_mass, src->wimp_day, double(host::kXeIsotopes[k]), &ok);
    if (!ok) return fail(c, NB200_EPHYSICS, "\tThe velocity integral in the WIMP generator broke!!!");
  }
  return 0;
}

// host <-> device staging of the small stage calls: inputs copied in, outputs copied back by finish()
struct Staged {
  nb200_ctx* c;
  bool host;
  std::vector<void*> owned;
  struct Out {
    void* host;
    void* dev;
    size_t bytes;
  };
  std::vector<Out> outs;
  cudaError_t err = cudaSuccess;
  Staged(nb200_ctx* c_, int mem_kind) : c(c_), host(mem_kind == NB200_MEM_HOST) {}
  template <class T>
  const T* in(const T* p, size_t count) {
    if (!host || !p) return p;
    void* d = nullptr;
    if (err == cudaSuccess) err = cudaMalloc(&d, count * sizeof(T));
    if (err == cudaSuccess) {
      owned.push_back(d);
      err = cudaMemcpyAsync(d, p, count * sizeof(T), cudaMemcpyHostToDevice, c->stream);
    }
    return static_cast<const T*>(d);
  }
  template <class T>
  T* out(T* p, size_t count) {
    if (!host || !p) return p;
    void* d = nullptr;
    if (err == cudaSuccess) err = cudaMalloc(&d, count * sizeof(T));
    if (err == cudaSuccess) {
      owned.push_back(d);
      outs.push_back({p, d, count * sizeof(T)});
    }
    return static_cast<T*>(d);
  }
  int finish(int rc, const char* what) {
    cudaError_t le = cudaGetLastError();
    if (!rc && err != cudaSuccess) rc = fail(c, NB200_ENOMEM, cudaGetErrorString(err));
    if (!rc && le != cudaSuccess) rc = fail(c, NB200_ECUDA, std::string(what) + ": " + cudaGetErrorString(le));
    if (!rc)
      for (auto& o : outs) {
        cudaError_t e = cudaMemcpyAsync(o.host, o.dev, o.bytes, cudaMemcpyDeviceToHost, c->stream);
        if (e != cudaSuccess) rc = fail(c, NB200_ECUDA, cudaGetErrorString(e));
      }
    if (!rc) rc = check_device_error(c);  // synchronises
    else cudaStreamSynchronize(c->stream);
    for (void* q : owned) cudaFree(q);
    return rc;
  }
};

// Everything the kernel needs beyond the caller's PODs
int build_params(nb200_ctx* c, const nb200_detector* det, const nb200_model* mo, const nb200_analysis* ana,
                 const nb200_source* src, const nb200_run_consts* rc, uint64_t seed, uint64_t first,
                 uint64_t n, RunParams* P, bool keep_memo = false) {
  if (!keep_memo) c->call_memo.clear();
  int r = validate(c, det, mo, ana, src);
  if (r) return r;
  if (!rc) return fail(c, NB200_EINVAL, "null run constants (call nb200_prepare)");
  std::memset(P, 0, sizeof *P);
  P->det = *det;
  P->model = *mo;
  if (P->model.n_widths > 17) P->model.n_widths = 17;
  P->ana = *ana;
  P->src = *src;
  P->rc = *rc;
  if ((r = upload_map(c, &P->det.FitS1))) return r;
  if ((r = upload_map(c, &P->det.FitS2))) return r;
  if ((r = upload_map(c, &P->det.FitEF))) return r;
  if ((r = upload_map(c, &P->det.FitTBA))) return r;
  // ion: execNEST.cpp:1056-1063 swaps the width vector
  if (!src->vec_mode && src->species == NB200_ion) {
    const double w[13] = {1.00, 1.00, 0., 0.50, 0.19, 0., 1., 0.046452, 0.45, 0.205, -0.2, 0., 0.};
    std::memset(P->model.NRERWidthsParam, 0, sizeof P->model.NRERWidthsParam);
    std::memcpy(P->model.NRERWidthsParam, w, sizeof w);
    P->model.n_widths = 13;
  }
  // yield-model factors of the uniform drift field (an event with another field evaluates its own)
  P->hoist.field = -1.;
  if (src->inField >= 0.) {
    P->hoist.field = src->inField;
    P->hoist.density = rc->rho;
    P->hoist.nr_ti = nr_thomas_imel(src->inField, rc->rho, mo->NRYieldsParam);
    beta_field_terms(src->inField, rc->rho, mo->ERYieldsParam, P->hoist.b);
    gamma_field_terms(src->inField, rc->rho, P->hoist.g);
    const double* wv = P->model.NRERWidthsParam;  // after the ion swap above
    if (!(wv[9] > wv[8])) omega_er_consts(src->inField, wv, P->model.n_widths, P->hoist.w);
    else P->hoist.w[0] = P->hoist.w[1] = P->hoist.w[2] = 0.;
  }
  P->z_direct = 0;
  if (!src->vec_mode && src->inField >= 1. /* FIELD_MIN */ && rc->vD > 0. && (src->fixed_pos == 0 || src->fixed_pos == 2) &&
      !(src->kind == NB200_SRC_LIST && src->x)) {
    const double lo = std::max(0., det->TopDrift - rc->vD * det->dt_max);
    const double hi = std::min(det->TopDrift, det->TopDrift - rc->vD * det->dt_min);
    if (hi > lo) {
      P->z_lo = lo;
      P->z_hi = hi;
      P->z_direct = 1;
    }
  }
  // per-run hoists (host maps evaluate the caller's LUT, not the device copy)
  const double zc = det->TopDrift - rc->vD_middle * det->dtCntr;
  P->s1_center = fit_s1(det->FitS1, 0., 0., zc);
  P->s2_center = fit_s2(det->FitS2, 0., 0.);
  P->coin_u_min = exp(-det->coinWind / 20.);
  for (int k = 0; k <= 10; ++k) P->coin_pow[k] = pow(double(det->numPMTs), 1. - double(k));
  {
    double t = ceil(det->P_dphe * 4294967296.0 - 0.5);  // (w + 1/2)/2^32 < P_dphe  <=>  w < t
    P->dphe_thr = t <= 0. ? 0ull : (t >= 4294967296.0 ? 4294967296ull : (unsigned long long)t);
  }
  {
    // S = T + N decomposition of the single-phe area (nb_params.h)
    const double mu = 1. / rc->NewMean, sg = det->sPEres, nb0 = det->noiseBaseline[0], nu = det->noiseBaseline[1];
    const double var = sg * sg + nu * nu;
    P->pe_mu = mu + nb0;
    P->pe_sig = sqrt(var);
    P->pe_rho = var > 0. ? sg * sg / var : 0.;
    P->pe_m0 = mu - P->pe_rho * (mu + nb0);
    P->pe_tau = var > 0. ? sg * fabs(nu) / sqrt(var) : 0.;
    // P(T <= 0 | S = s) = Phi(-(m0 + rho s)/tau) < 2^-64  <=>  m0 + rho s > 9.1 tau
    P->pe_cut = P->pe_rho > 0. ? (9.1 * P->pe_tau - P->pe_m0) / P->pe_rho : -DBL_MAX;
    P->pe_fast = 1.5341205443525465 * P->pe_tau;  // Phi(1.5341205...) = 15/16
    // m = pe_m0 + pe_rho S < pe_fast as a threshold on S itself (pe_rho = 0: m is constant)
    P->pe_sfast = P->pe_rho > 0. ? (P->pe_fast - P->pe_m0) / P->pe_rho : (P->pe_m0 < P->pe_fast ? DBL_MAX : -DBL_MAX);
    P->pe_inv_tau16 = P->pe_tau > 0. ? 16. / P->pe_tau : 0.;
  }
  {
    const double sq2 = sqrt(2.);
    double a_spe = 0.5 * (1. + erf(-1. / det->sPEres / sq2));
    double x_spe = 0.5 * (1. + erf((det->sPEthr - 1.) / det->sPEres / sq2));
    double spe = (x_spe - a_spe) / (1. - a_spe);
    double a_dpe = 0.5 * (1. + erf(-2. / (sq2 * det->sPEres) / sq2));
    double x_dpe = 0.5 * (1. + erf((det->sPEthr - 2.) / (sq2 * det->sPEres) / sq2));
    double dpe = (x_dpe - a_dpe) / (1. - a_dpe);
    P->below_thr_pct = spe * (1. - det->P_dphe) + dpe * det->P_dphe;
    P->recon_mult = 0.5 * (1. + erf((det->sPEthr - 1.) / (det->sPEres * sqrt(2.)))) /
                    (det->sPEres * sqrt(2. * M_PI));
  }
  P->band_fx_x = band_scale_x(ana);
  if (ana->useS2 == 2) {
    P->band_width = (ana->maxS2 - ana->minS2) / double(ana->numBins);
    P->band_border = ana->minS2;
  } else {
    P->band_width = (ana->maxS1 - ana->minS1) / double(ana->numBins);
    P->band_border = ana->minS1;
  }
  {
    double K = det->T_Kelvin;
    int i = host::drift_bracket(K);
    if (i < 0) return fail(c, NB200_EPHYSICS, "ERROR: TEMPERATURE OUT OF RANGE (100-230 K)");
    std::memcpy(P->dv_row[0], host::kPolyExp[i], sizeof(double) * 7);
    std::memcpy(P->dv_row[1], host::kPolyExp[i + 1], sizeof(double) * 7);
    P->dv_Ti = host::kDriftT[i];
    P->dv_Tf = host::kDriftT[i + 1];
    P->dv_T = K;
    P->dv_mode = nearly_equal(K, P->dv_Ti) ? 1 : (nearly_equal(K, P->dv_Tf) ? 2 : 0);
  }
  if (src->inField == -1. && !src->vec_mode) {
    const double zs = ana->z_step > 0. ? ana->z_step : 0.1;
    if (c->vtable_host.empty() || std::memcmp(&c->vtable_det, det, sizeof *det) != 0 || c->vtable_zstep != zs) {
      c->vtable_host = host::drift_table_nonuniform(*det, zs, 0., 0.);
      c->vtable_det = *det;
      c->vtable_zstep = zs;
      if (c->d_vtable) cudaFree(c->d_vtable);
      NB_CUDA(c, cudaMalloc(&c->d_vtable, sizeof(double) * c->vtable_host.size()));
      NB_CUDA(c, cudaMemcpyAsync(c->d_vtable, c->vtable_host.data(), sizeof(double) * c->vtable_host.size(),
                                 cudaMemcpyHostToDevice, c->stream));
    }
    P->vtable = c->d_vtable;
    P->vtable_n = int(c->vtable_host.size());
  }
  P->dpe_alias = nullptr;
  if (det->P_dphe > 0. && det->P_dphe < 1.)  // alias tables of Binomial(2^b, P_dphe) for k_pre's double-phe count
    if ((r = alias_table_for(c, det->P_dphe, &P->dpe_alias))) return r;
  if ((r = upload_wimp_source(c, src, P))) return r;
  P->seed = seed;
  philox_round_keys(seed, P->rk);
  P->first_event = first;
  P->n_events = n;
  P->species = src->species;
  P->vec_mode = src->vec_mode;
  return 0;
}

// events per internal chunk: the workspace holds one chunk (140 B/event, + 48 B/event of deferred queue).
// ~2^chunk_log2, rounded to a whole number of waves of the one-block-per-SM lockstep kernel (k_post: sm_count
// blocks of 1024 events per sweep; 113 sweeps of 148 x 1024 on B200 at 2^24), which also makes it a whole number of
// k_hits groups per resident block.  Every chunk costs four kernel boundaries (drain, launch, ramp, the blocks'
// table loads and band flushes) of ~20 us each; measured on B200, 10^8 events: 2^21 83.1 ms, 2^22 80.7, 2^23 79.2,
// 2^24 78.8, 2^25 78.6, 2^26 78.6.  2^24 (3.2 GB of device memory once a call is that large) is the default;
// a device that cannot give that much falls back to smaller chunks (ensure_workspace).
uint64_t chunk_events(const nb200_ctx* c) {
  const uint64_t wave = uint64_t(c->sm_count > 0 ? c->sm_count : 148) * 1024u;
  return std::max<uint64_t>(1, ((1ull << c->chunk_log2) + wave / 2) / wave) * wave;
}

int ensure_workspace(nb200_ctx* c, uint64_t n, Work* w) {
  for (;;) {
    const uint64_t chunk = chunk_events(c);
    const uint64_t need = std::min<uint64_t>(chunk, std::max<uint64_t>(n, 1024));
    if (need <= c->work_cap) break;
    if (c->d_work) cudaFree(c->d_work);
    c->d_work = nullptr;
    c->work_cap = 0;
    const cudaError_t e = cudaMalloc(&c->d_work, need * (WORK_F64 * 8 + WORK_I32 * 4));
    if (e == cudaSuccess) {
      c->work_cap = need;
      break;
    }
    cudaGetLastError();  // clear the allocation failure
    if (need < chunk || c->chunk_log2 <= 20) return fail(c, NB200_ENOMEM, std::string("workspace: ") + cudaGetErrorString(e));
    --c->chunk_log2;  // a full chunk did not fit: halve it, for this context's lifetime
  }
  const uint64_t cap = c->work_cap;
  double* f = static_cast<double*>(c->d_work);
  double** fp[WORK_F64] = {&w->keV, &w->px, &w->py, &w->pz, &w->sx, &w->sy, &w->field, &w->dt,
                           &w->yNph, &w->yNe, &w->yL, &w->posDepSm, &w->fitSm, &w->area};
  for (int k = 0; k < WORK_F64; ++k) *fp[k] = f + cap * k;
  int32_t* i32 = reinterpret_cast<int32_t*>(f + cap * WORK_F64);
  int32_t** ip[WORK_I32] = {&w->photons, &w->electrons, &w->ions, &w->excitons, &w->nHits, &w->spike, &w->nphe};
  for (int k = 0; k < WORK_I32; ++k) *ip[k] = i32 + cap * k;
  return 0;
}

// once per context: opt-in shared-memory sizes and resident blocks per SM of the chain kernels
int ensure_occupancy(nb200_ctx* c) {
  if (!c->occ_pre) {
    NB_CUDA(c, cudaFuncSetAttribute(k_post, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    NB_CUDA(c, cudaFuncSetAttribute(k_hits, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kHitDynSmem)));
    NB_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c->occ_pre, k_pre, kPreBlock, 0));
    NB_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c->occ_hits, k_hits, kHitBlock, kHitDynSmem));
    if (c->occ_pre < 1) c->occ_pre = 1;
    if (c->occ_hits < 1) c->occ_hits = 1;
  }
  return 0;
}

// the deferred-finisher queue: 6 entries per workspace event (LUX NR 0-100 keV queues 2 per event; what does not
// fit is finished inside k_hits), allocated with the workspace
int ensure_hit_queue(nb200_ctx* c, HitQueue* q) {
  uint64_t want = std::min<uint64_t>(std::max<uint64_t>(c->work_cap, 1024) * 6, 1u << 28);
  while (want > c->hitq_cap) {
    if (c->d_hitq) cudaFree(c->d_hitq);
    c->d_hitq = nullptr;
    c->hitq_cap = 0;
    const cudaError_t e = cudaMalloc(&c->d_hitq, want * sizeof(SlotEntry));
    if (e == cudaSuccess) {
      c->hitq_cap = unsigned(want);
      break;
    }
    cudaGetLastError();
    if (want <= 1024) return fail(c, NB200_ENOMEM, std::string("deferred queue: ") + cudaGetErrorString(e));
    want /= 2;  // a shorter queue only moves more of the finishing into k_hits
  }
  q->entries = c->d_hitq;
  q->count = reinterpret_cast<unsigned int*>(c->d_err + 2);
  q->cap = c->hitq_cap;
  return 0;
}

// k_hits + k_hits_finish for the m events of Q's workspace chunk
void launch_hits(nb200_ctx* c, const RunParams& Q, uint64_t m, const HitQueue& hq) {
  const uint64_t sm = uint64_t(c->sm_count);
  const unsigned g_hits = unsigned(std::min<uint64_t>((m + kHitGroup - 1) / kHitGroup, sm * c->occ_hits));  // persistent
  cudaMemsetAsync(c->d_err + 1, 0, 2 * sizeof(int), c->stream);
  k_hits<<<g_hits, kHitBlock, kHitDynSmem, c->stream>>>(Q, reinterpret_cast<unsigned int*>(c->d_err + 1), hq);
  k_hits_finish<<<unsigned(sm * NB_FIN_MINB * 2), kFinBlock, 0, c->stream>>>(Q, hq);
}

// k_pre -> k_hits -> k_post over the call's events in chunks of the workspace size
int launch_event(nb200_ctx* c, RunParams& P) {
  if (P.n_events == 0) return 0;
  size_t smem = 0;
  P.band_in_smem = 0;
  if (P.do_band) {
    smem = size_t(P.ana.numBins) * BAND_WORDS * 8 + size_t(P.ana.numBins) * size_t(P.ana.logBins) * 4;
    if (smem <= 160 * 1024)
      P.band_in_smem = 1;
    else
      smem = 0;
  }
  int ro = ensure_occupancy(c);
  if (ro) return ro;
  NB_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c->occ_post, k_post, kPostBlock, smem));
  if (c->occ_post < 1) c->occ_post = 1;
  int r = ensure_workspace(c, P.n_events, &P.work);
  if (r) return r;
  HitQueue hq;
  if ((r = ensure_hit_queue(c, &hq))) return r;
  const uint64_t total = P.n_events, cap = c->work_cap;
  for (uint64_t off = 0; off < total; off += cap) {
    const uint64_t m = std::min<uint64_t>(cap, total - off);
    RunParams Q = P;
    Q.n_events = m;
    Q.chunk_off = P.chunk_off + off;
    const uint64_t sm = uint64_t(c->sm_count);
    unsigned g_pre = unsigned(std::min<uint64_t>((m + kPreBlock - 1) / kPreBlock, sm * c->occ_pre * 8));
    unsigned g_post = unsigned(std::min<uint64_t>((m + kPostBlock - 1) / kPostBlock, sm * c->occ_post));
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    if (c->timing)
      for (auto& e : ev) cudaEventCreate(&e);
    if (c->timing) cudaEventRecord(ev[0], c->stream);
    k_pre<<<g_pre, kPreBlock, 0, c->stream>>>(Q, c->d_err);
    if (c->timing) cudaEventRecord(ev[1], c->stream);
    launch_hits(c, Q, m, hq);
    if (c->timing) cudaEventRecord(ev[2], c->stream);
    k_post<<<g_post, kPostBlock, smem, c->stream>>>(Q, c->d_err);
    if (c->timing) {
      cudaEventRecord(ev[3], c->stream);
      for (auto& e : ev) c->tev.push_back(e);
    }
    NB_CUDA(c, cudaGetLastError());
    c->launches += 4;
  }
  return 0;
}

int check_device_error(nb200_ctx* c) {
  int h = 0;
  NB_CUDA(c, cudaMemcpyAsync(&h, c->d_err, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  NB_CUDA(c, cudaStreamSynchronize(c->stream));
  if (h) {
    NB_CUDA(c, cudaMemsetAsync(c->d_err, 0, sizeof(int), c->stream));
    if (h & 2) return fail(c, NB200_EINVAL, "species not supported by the GPU chain");
    if (h & 4) return fail(c, NB200_EPHYSICS, "ERROR: min greater than max");  // NEST.cpp:957-959
    return fail(c, NB200_EPHYSICS, "ERROR: Quanta not conserved. Tell Matthew Immediately!");
  }
  return 0;
}

// SplitMix-free URBG over the Philox host stream, for libstdc++ distributions
struct HostUrbg {
  Stream s;
  using result_type = uint64_t;
  static constexpr result_type min() { return 0; }
  static constexpr result_type max() { return ~result_type(0); }
  result_type operator()() {
    uint32_t a, b;
    s.next2(a, b);
    return (uint64_t(a) << 32) | b;
  }
};

}  // namespace

extern "C" {

int nb200_abi_version(void) { return NB200_ABI_VERSION; }

int nb200_sizeof(int which) {
  switch (which) {
    case 0: return int(sizeof(nb200_map));
    case 1: return int(sizeof(nb200_detector));
    case 2: return int(sizeof(nb200_model));
    case 3: return int(sizeof(nb200_analysis));
    case 4: return int(sizeof(nb200_source));
    case 5: return int(sizeof(nb200_run_consts));
    case 6: return int(sizeof(nb200_soa_out));
    case 7: return int(sizeof(nb200_band_hist));
    default: return -1;
  }
}

int nb200_create(int device, nb200_ctx** out) {
  if (!out) return NB200_EINVAL;
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= 0)
    return fail(nullptr, NB200_ENODEV, std::string("no CUDA device: ") + cudaGetErrorString(e));
  if (device < 0 || device >= count) return fail(nullptr, NB200_EINVAL, "bad device ordinal");
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return fail(nullptr, NB200_ECUDA, "cudaGetDeviceProperties");
  if (prop.major != 10)
    return fail(nullptr, NB200_ENODEV, "device is not sm_100 (this library carries sm_100a code only)");
  nb200_ctx* c = new nb200_ctx();
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaMalloc(&c->d_err, 4 * sizeof(int)) != cudaSuccess || cudaMemset(c->d_err, 0, 4 * sizeof(int)) != cudaSuccess) {
    g_err_noctx = std::string("context setup: ") + cudaGetErrorString(cudaGetLastError());
    delete c;
    return NB200_ECUDA;
  }
  c->stream = c->own_stream;
  if (const char* e = std::getenv("NB200_CHUNK_LOG2")) c->chunk_log2 = std::min(25, std::max(16, std::atoi(e)));  // <= 2^25: queue entries index a chunk with 26 bits
  {  // Phi table of the photo-electron acceptance bounds (nb_chain.cuh)
    double tab[NB_PHI_N];
    for (int k = 0; k < NB_PHI_N; ++k) tab[k] = 0.5 * erfc(-(-8. + double(k) / 16.) * 0.70710678118654752440);
    if (cudaMemcpyToSymbol(g_phitab, tab, sizeof tab) != cudaSuccess) {
      g_err_noctx = std::string("context setup: ") + cudaGetErrorString(cudaGetLastError());
      nb200_destroy(c);
      return NB200_ECUDA;
    }
  }
  {  // ziggurat layers of the hit loop's normal sampler (nb_rng.cuh)
    std::vector<double> zx(NB_ZIG_N + 1), zf(NB_ZIG_N + 1), zw(NB_ZIG_N);
    std::vector<uint32_t> zk(NB_ZIG_N);
    build_zig_tables(zx.data(), zf.data(), zk.data(), zw.data());
    const double R = zx[1];
    if (cudaMemcpyToSymbol(g_zig_k, zk.data(), sizeof(uint32_t) * NB_ZIG_N) != cudaSuccess ||
        cudaMemcpyToSymbol(g_zig_w, zw.data(), sizeof(double) * NB_ZIG_N) != cudaSuccess ||
        cudaMemcpyToSymbol(g_zig_f, zf.data(), sizeof(double) * (NB_ZIG_N + 1)) != cudaSuccess ||
        cudaMemcpyToSymbol(g_zig_R, &R, sizeof R) != cudaSuccess) {
      g_err_noctx = std::string("context setup: ") + cudaGetErrorString(cudaGetLastError());
      nb200_destroy(c);
      return NB200_ECUDA;
    }
  }
  {
    double rcp[NB_RCP_N];
    rcp[0] = 0.;
    for (int k = 1; k < NB_RCP_N; ++k) rcp[k] = 1. / double(k);
    if (cudaMemcpyToSymbol(g_rcptab, rcp, sizeof rcp) != cudaSuccess) {
      g_err_noctx = std::string("context setup: ") + cudaGetErrorString(cudaGetLastError());
      nb200_destroy(c);
      return NB200_ECUDA;
    }
  }
  {  // ln k! for the binomial sampler's exact acceptance test (nb_rng.cuh)
    static std::vector<double> lf;  // once per process
    if (lf.empty()) {
      lf.resize(NB_LFACT_N);
      for (int k = 0; k < NB_LFACT_N; ++k) lf[k] = double(lgammal((long double)k + 1.0L));
    }
    if (cudaMemcpyToSymbol(g_lfact, lf.data(), sizeof(double) * NB_LFACT_N) != cudaSuccess) {
      g_err_noctx = std::string("context setup: ") + cudaGetErrorString(cudaGetLastError());
      nb200_destroy(c);
      return NB200_ECUDA;
    }
  }
  *out = c;
  return 0;
}

void nb200_destroy(nb200_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  for (auto e : c->tev) cudaEventDestroy(e);
  for (auto& b : c->blobs) cudaFree(b.dev);
  if (c->d_band && !c->band_external) cudaFree(c->d_band);
  if (c->d_band_tmp) cudaFree(c->d_band_tmp);
  if (c->d_vtable) cudaFree(c->d_vtable);
  if (c->d_fit_scratch) cudaFree(c->d_fit_scratch);
  for (auto& e : c->dpe) cudaFree(e.dev);
  if (c->d_sweep_params) cudaFree(c->d_sweep_params);
  if (c->d_sweep_bands) cudaFree(c->d_sweep_bands);
  if (c->h_sweep_bands) cudaFreeHost(c->h_sweep_bands);
  if (c->d_work) cudaFree(c->d_work);
  for (int b = 0; b < 2; ++b) {
    if (c->d_stage[b]) cudaFree(c->d_stage[b]);
    if (c->ev_done[b]) cudaEventDestroy(c->ev_done[b]);
    if (c->ev_copied[b]) cudaEventDestroy(c->ev_copied[b]);
  }
  if (c->d_list) cudaFree(c->d_list);
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  if (c->d_err) cudaFree(c->d_err);
  if (c->d_hitq) cudaFree(c->d_hitq);
  if (c->own_stream) cudaStreamDestroy(c->own_stream);
  delete c;
}

const char* nb200_last_error(const nb200_ctx* c) { return c ? c->err.c_str() : g_err_noctx.c_str(); }

int nb200_set_stream(nb200_ctx* c, void* s) {
  if (!c) return NB200_EINVAL;
  cudaStream_t ns = s ? static_cast<cudaStream_t>(s) : c->own_stream;
  if (ns != c->stream) {
    // kernels queued on the old stream share this context's workspace, queue and counters with whatever is
    // launched next: let them finish before the switch
    NB_CUDA(c, cudaSetDevice(c->device));
    NB_CUDA(c, cudaStreamSynchronize(c->stream));
    c->stream = ns;
  }
  return 0;
}

// also reports what the kernels of asynchronous calls flagged (the reference throws / returns 1 for these:
// "Quanta not conserved", "min greater than max", unsupported species)
int nb200_synchronize(nb200_ctx* c) {
  if (!c) return NB200_EINVAL;
  NB_CUDA(c, cudaSetDevice(c->device));
  return check_device_error(c);  // synchronises the stream
}

uint64_t nb200_launch_count(const nb200_ctx* c) { return c ? c->launches : 0; }

// ---- host-only helpers ------------------------------------------------------------------

int nb200_detector_builtin(const char* name, nb200_detector* d) {
  if (!name || !d) return NB200_EINVAL;
  std::memset(d, 0, sizeof *d);
  std::string n(name);
  auto cmap = [](double v) {
    nb200_map m;
    std::memset(&m, 0, sizeof m);
    m.kind = NB200_MAP_CONST;
    m.c[0] = v;
    return m;
  };
  if (n == "VDetector" || n == "LArDetector") {  // VDetector.hh:157-224
    d->g1 = 0.0760; d->sPEres = 0.58; d->sPEthr = 0.35; d->sPEeff = 1.00;
    d->P_dphe = 0.2; d->coinWind = 100; d->coinLevel = 2; d->numPMTs = 89; d->OldW13eV = 1;
    d->noiseLinear[0] = 3e-2; d->noiseLinear[1] = 3e-2;
    d->g1_gas = 0.06; d->s2Fano = 3.61; d->s2_thr = 300.; d->E_gas = 12.; d->eLife_us = 2200.;
    d->T_Kelvin = 177.; d->p_bar = 2.14; d->dtCntr = 40.; d->dt_min = 20.; d->dt_max = 60.;
    d->radius = 50.; d->radmax = 50.; d->TopDrift = 150.; d->anode = 152.5; d->gate = 147.5; d->cathode = 1.00;
    d->PosResFlat = 3.00; d->PosResBase = 70.8364; d->molarMass = 131.293;
    d->FitS1 = cmap(1.); d->FitS2 = cmap(1.); d->FitEF = cmap(730.); d->FitTBA = cmap(0.5);
    return 0;
  }
  if (n == "LUX_Run03" || n == "LUX Run03") {  // LUX_Run03.hh:30-88
    d->g1 = 0.1170; d->sPEres = 0.37; d->sPEthr = (0.3 * 1.173) / 0.915; d->sPEeff = 1.00;
    d->noiseBaseline[0] = 0.00; d->noiseBaseline[1] = 0.08;
    d->P_dphe = 0.173; d->coinWind = 100; d->coinLevel = 2; d->numPMTs = 119; d->OldW13eV = 1;
    d->noiseLinear[0] = 0.0e-2; d->noiseLinear[1] = 9.0e-2;
    d->g1_gas = 0.1; d->s2Fano = 3.6; d->s2_thr = (150. * 1.173) / 0.915; d->E_gas = 6.25; d->eLife_us = 800.;
    d->T_Kelvin = 173.; d->p_bar = 1.57; d->dtCntr = 160.; d->dt_min = 38.; d->dt_max = 305.;
    d->radius = 200.; d->radmax = 235.; d->TopDrift = 544.95; d->anode = 549.2; d->gate = 539.2; d->cathode = 55.90;
    d->PosResFlat = 3.00; d->PosResBase = 70.8364; d->molarMass = 131.293;
    nb200_map m;
    std::memset(&m, 0, sizeof m);
    m.kind = NB200_MAP_LUX_RUN03;
    const double s1c[7] = {307.9, 0.3071, 0.0002257, 1.1525e-7, 318.84, 307.9, 235.};
    std::memcpy(m.c, s1c, sizeof s1c);
    d->FitS1 = m;
    std::memset(&m, 0, sizeof m);
    m.kind = NB200_MAP_LUX_RUN03;
    const double s2c[10] = {9156.3, 6.22750, 0.38126, 0.017144, 0.0002474, 1.6953e-6, 5.6513e-9, 7.3989e-12, 9156.3, 235.};
    std::memcpy(m.c, s2c, sizeof s2c);
    d->FitS2 = m;
    std::memset(&m, 0, sizeof m);
    m.kind = NB200_MAP_LUX_RUN03;
    const double efc[6] = {158.92, 0.2209000, 0.0024485, 8.7098e-6, 1.5049e-8, 1.0110e-11};
    std::memcpy(m.c, efc, sizeof efc);
    d->FitEF = m;
    d->FitTBA = cmap(0.449);
    return 0;
  }
  return fail(nullptr, NB200_EINVAL, "unknown built-in detector " + n);
}

int nb200_model_default(int which, nb200_model* m) {
  if (!m) return NB200_EINVAL;
  std::memset(m, 0, sizeof *m);
  const double nr[12] = {11., 1.1, 0.0480, -0.0533, 12.6, 0.3, 2., 0.3, 2., 0.5, 1., 1.};  // NEST.hh:120-121
  std::memcpy(m->NRYieldsParam, nr, sizeof nr);
  for (int i = 0; i < 10; ++i) m->ERYieldsParam[i] = -1.;  // NEST.hh:124
  if (which == 0) {  // NEST.hh:122-123
    const double w[13] = {0.4, 0.4, 0.04, 0.5, 0.19, 2.25, -0.001, 0.046452, 0.45, 0.205, -0.2, 0., 0.};
    std::memcpy(m->NRERWidthsParam, w, sizeof w);
    m->SkewnessER = -999.;
  } else {  // execNEST.cpp:188-204, :1065
    const double w[13] = {0.4, 0.4, 0.04, 0.50, 0.19, 2.25, -0.00050, 0.03290, 2.629, 0.2964, -.3343, 0.0000, 0.0000};
    std::memcpy(m->NRERWidthsParam, w, sizeof w);
    m->SkewnessER = 0.;
  }
  m->n_widths = 13;
  m->er_model = -1;
  m->multFact = 1.;
  m->massNum = 0.;
  m->atomNum = 0.;
  return 0;
}

int nb200_analysis_default(int which, nb200_analysis* a) {
  if (!a) return NB200_EINVAL;
  std::memset(a, 0, sizeof *a);
  a->s1mode = NB200_S1_HYBRID;
  a->s2mode = NB200_S2_FULL;
  a->logBins = 30;
  if (which == 0) {  // analysis.hh
    a->usePD = 2; a->useS2 = 0; a->numBins = 98; a->minS1 = 1.5; a->maxS1 = 99.5;
    a->minS2 = 42.; a->maxS2 = 1e4; a->logMin = 0.6; a->logMax = 3.6;
    a->MCtruthE = 0; a->MCtruthPos = 0; a->z_step = 0.1;
  } else {  // analysis_G2.hh
    a->usePD = 1; a->useS2 = 1; a->numBins = 118; a->minS1 = 2.5; a->maxS1 = 120.5;
    a->minS2 = 200.; a->maxS2 = 9e4; a->logMin = 2.; a->logMax = 5.;
    a->MCtruthE = 1; a->MCtruthPos = 0; a->z_step = 1.;
  }
  return 0;
}

int nb200_prepare(nb200_detector* det, double inField, double z_step, nb200_run_consts* out) {
  if (!det || !out) return NB200_EINVAL;
  std::memset(out, 0, sizeof *out);
  bool inGas = false;
  std::string warn;
  double rho = host::density(det->T_Kelvin, det->p_bar, inGas, det->molarMass, &warn);
  det->inGas = inGas ? 1 : 0;
  if (rho <= 0. || det->T_Kelvin <= 0. || det->p_bar <= 0.)  // execNEST.cpp:613-617
    return fail(nullptr, NB200_EINVAL, "ERR: Unphysical thermodynamic property!");
  if (rho < 1.75) det->inGas = 1;  // execNEST.cpp:618-619
  if (det->inGas)
    return fail(nullptr, NB200_EINVAL, "gas-phase detectors (MagBoltz drift) are not part of this path");
  Wvalue w = work_function(rho, det->molarMass, det->OldW13eV != 0);
  out->rho = rho;
  out->Wq_eV = det->OldW13eV ? w.Wq_eV : 11.5;  // execNEST.cpp:626-627
  out->alpha = w.alpha;
  if (!host::calculate_g2(*det, out->g2_params))
    return fail(nullptr, NB200_EPHYSICS, "\tERR: The gas gap in the S2 calculation broke!!!!");
  out->NewMean = host::find_new_mean(det->sPEres);
  out->field = inField;
  bool ok = true;
  if (inField == -1.) {
    const double zs = z_step > 0. ? z_step : 0.1;
    std::vector<double> t = host::drift_table_nonuniform(*det, zs, 0., 0.);
    double centralZ = (det->gate * 0.8 + det->cathode * 1.03) / 2.;
    size_t i = size_t(floor(centralZ / zs + 0.5));
    if (t.empty() || i >= t.size()) return fail(nullptr, NB200_EINVAL, "drift table does not reach centralZ");
    out->vD_middle = t[i];
    out->vD = out->vD_middle;
  } else {
    out->vD_middle = host::drift_velocity_liquid(det->T_Kelvin, inField, &ok);
    out->vD = out->vD_middle;
  }
  if (!ok) return fail(nullptr, NB200_EPHYSICS, "ERROR: TEMPERATURE OUT OF RANGE (100-230 K)");
  // execNEST.cpp:955-961
  if (inField != -1. && det->dt_min > (det->TopDrift - 0.) / out->vD && inField >= NB_FIELD_MIN)
    return fail(nullptr, NB200_EINVAL, "ERROR: dt_min is too restrictive (too large)");
  return 0;
}

}  // extern "C"

namespace {
// TestSpectra::WIMP_prep_spectrum (TestSpectra.cpp:392-454); next_iso() supplies the Xe isotope of each
// WIMP_dRate call in call order (the reference draws it inside, :278); returns false when it runs out
template <class NextIso>
int wimp_prep_core(double mass, double eStep, double dayNum, NextIso&& next_iso, double* base, double* exponent,
                   double* integral, double* xMaxOut, double* divisorOut) {
  for (int i = 0; i < 100; ++i) {  // TestSpectra.hh:44-45 initialisers
    base[i] = (i == 0) ? 1. : 0.;
    exponent[i] = 0.;
  }
  double divisor;
  int numberPoints;
  if (mass < 10.) {
    divisor = 10. / eStep;
    numberPoints = int(1000. / eStep);
  } else {
    divisor = 1.0 / eStep;
    numberPoints = int(100. / eStep);
  }
  std::vector<double> spec;
  int nZeros = 0;
  bool ok = true;
  int A = 0;
  for (int i = 0; i < numberPoints + 1; ++i) {
    if (!next_iso(&A)) return fail(nullptr, NB200_EINVAL, "nb200_wimp_prep_iso: isotope list too short");
    spec.push_back(host::wimp_drate(double(i) / divisor, mass, dayNum, double(A), &ok));
    if (!ok) return fail(nullptr, NB200_EPHYSICS, "\tThe velocity integral in the WIMP generator broke!!!");
    if (nearly_equal(spec[i], 0.))
      ++nZeros;
    else
      nZeros = 0;
    if (nZeros == 100) break;
  }
  double integ = 0.;
  for (uint64_t i = 0; i < 1000000; ++i) {
    if (!next_iso(&A)) return fail(nullptr, NB200_EINVAL, "nb200_wimp_prep_iso: isotope list too short");
    integ += host::wimp_drate(double(i) / 1e4, mass, dayNum, double(A), &ok) / 1e4;
  }
  double xMax = (double(spec.size()) - 1.) / divisor;
  for (int i = 0; i < int(spec.size()) - 1; ++i) {
    if (i >= 100) return fail(nullptr, NB200_EPHYSICS, "ERROR: WIMP E_step is too small: more than 100 table intervals");
    double x1 = double(i) / divisor, x2 = double(i + 1) / divisor;
    base[i] = spec[i + 1] * pow(spec[i + 1] / spec[i], x2 / (x1 - x2));
    exponent[i] = log(spec[i + 1] / spec[i]) / (x1 - x2);
    if (base[i] > 0. && base[i] < DBL_MAX && exponent[i] > 0. && exponent[i] < DBL_MAX)
      ;
    else {
      if (spec[i + 1] > 10.)
        return fail(nullptr, NB200_EPHYSICS,
                    "ERROR: WIMP E_step is too small (or large)! Increase(decrease) it slightly to avoid "
                    "noise in the calculation.");
      xMax = double(i - 1) / divisor;
      if (xMax <= 0.0)
        return fail(nullptr, NB200_EPHYSICS,
                    "ERROR: The maximum possible WIMP recoil is not +-ive, which usually means your E_step "
                    "is too small (OR it is too large).");
      break;
    }
  }
  *integral = integ;
  *xMaxOut = xMax;
  *divisorOut = divisor;
  return 0;
}
}  // namespace

extern "C" {

int nb200_wimp_prep(double mass, double eStep, double dayNum, uint64_t seed, double* base, double* exponent,
                    double* integral, double* xMaxOut, double* divisorOut) {
  if (!base || !exponent || !integral || !xMaxOut || !divisorOut) return NB200_EINVAL;
  uint32_t rk[20];
  philox_round_keys(seed, rk);
  Stream g;
  g.init(rk, 0, STREAM_HOST);
  return wimp_prep_core(mass, eStep, dayNum, [&](int* A) { *A = host::select_xe_atom(g); return true; }, base, exponent,
                        integral, xMaxOut, divisorOut);
}

int nb200_wimp_prep_iso(double mass, double eStep, double dayNum, const int32_t* isotopes, size_t n_isotopes, double* base,
                        double* exponent, double* integral, double* xMaxOut, double* divisorOut) {
  if (!isotopes || !base || !exponent || !integral || !xMaxOut || !divisorOut) return NB200_EINVAL;
  size_t at = 0;
  return wimp_prep_core(mass, eStep, dayNum,
                        [&](int* A) {
                          if (at >= n_isotopes) return false;
                          *A = int(isotopes[at++]);
                          return true;
                        },
                        base, exponent, integral, xMaxOut, divisorOut);
}

double nb200_wimp_drate(double ER_keV, double mass_GeV, double dayNum, int isotope_A, int* ok_out) {
  bool ok = true;
  const double v = host::wimp_drate(ER_keV, mass_GeV, dayNum, double(isotope_A), &ok);
  if (ok_out) *ok_out = ok ? 1 : 0;
  return v;
}

uint64_t nb200_poisson(double mean, uint64_t seed) {
  uint32_t rk[20];
  philox_round_keys(seed, rk);
  HostUrbg u;
  u.s.init(rk, 1, STREAM_HOST);
  return std::poisson_distribution<uint64_t>(mean)(u);
}

// ---- deterministic yields ----------------------------------------------------------------

int nb200_yields(nb200_ctx* c, const nb200_detector* det, const nb200_model* mo, int species, size_t n,
                 const double* E, const double* field, double density, int mem_kind, double* out) {
  if (!c) return NB200_EINVAL;
  if (!det || !mo || !E || !field || !out) return fail(c, NB200_EINVAL, "null argument");
  if (n == 0) return 0;
  if (species < 0 || species > NB200_NoneType) return fail(c, NB200_EINVAL, "species not supported");
  NB_CUDA(c, cudaSetDevice(c->device));
  YieldsParams P;
  std::memset(&P, 0, sizeof P);
  P.det = *det;
  P.model = *mo;
  P.species = species;
  P.which = mo->er_model;
  P.density = density;
  P.n = n;
  philox_round_keys(0, P.rk);
  double *dE = nullptr, *dF = nullptr, *dO = nullptr;
  if (mem_kind == NB200_MEM_HOST) {
    cudaError_t e = cudaMalloc(&dE, n * 8);
    if (e == cudaSuccess) e = cudaMalloc(&dF, n * 8);
    if (e == cudaSuccess) e = cudaMalloc(&dO, n * 48);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dE, E, n * 8, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dF, field, n * 8, cudaMemcpyHostToDevice, c->stream);
    if (e != cudaSuccess) {  // nothing allocated so far may leak
      if (dE) cudaFree(dE);
      if (dF) cudaFree(dF);
      if (dO) cudaFree(dO);
      return fail(c, NB200_ENOMEM, cudaGetErrorString(e));
    }
    P.E = dE;
    P.field = dF;
    P.out = dO;
  } else {
    P.E = E;
    P.field = field;
    P.out = out;
  }
  unsigned grid = unsigned(std::min<uint64_t>((n + 255) / 256, uint64_t(c->sm_count) * 8));
  k_yields<<<grid, 256, 0, c->stream>>>(P, c->d_err);
  cudaError_t le = cudaGetLastError();
  ++c->launches;
  int rc = 0;
  if (le != cudaSuccess) rc = fail(c, NB200_ECUDA, std::string("k_yields: ") + cudaGetErrorString(le));
  if (!rc && mem_kind == NB200_MEM_HOST) {
    cudaError_t e = cudaMemcpyAsync(out, dO, n * 48, cudaMemcpyDeviceToHost, c->stream);
    if (e != cudaSuccess) rc = fail(c, NB200_ECUDA, cudaGetErrorString(e));
  }
  if (!rc) rc = check_device_error(c);
  if (dE) cudaFree(dE);
  if (dF) cudaFree(dF);
  if (dO) cudaFree(dO);
  return rc;
}

int nb200_quanta(nb200_ctx* c, const nb200_detector* det, const nb200_model* mo, size_t n, const double* yields,
                 double density, uint64_t seed, uint64_t first_event, int mem_kind, int32_t* quanta, double* prob_var) {
  if (!c) return NB200_EINVAL;
  if (!det || !mo || !yields || !quanta || !prob_var) return fail(c, NB200_EINVAL, "null argument");
  if (mo->n_widths < 13)  // NEST.cpp:258-262
    return fail(c, NB200_EPHYSICS, "ERROR: You need a minimum of 13 free parameters for the resolution model.");
  if (n == 0) return 0;
  if (mo->NRERWidthsParam[9] > mo->NRERWidthsParam[8] && mem_kind == NB200_MEM_HOST) {
    // RecombOmegaER throws on its first call (NEST.cpp:183-184): any record with Lindhard == 1 reaches it
    for (size_t i = 0; i < n; ++i)
      if (nearly_equal(yields[3 * n + i], 1.) && yields[i] + yields[n + i] > 0.)
        return fail(c, NB200_EPHYSICS,
                    "Low-E r-fluct Gauss peak's width greater than centroid. This is physically wrong so "
                    "please check your inputs, in NRERWidthsParam: they must be in the old wrong order.");
  }
  NB_CUDA(c, cudaSetDevice(c->device));
  QuantaParams P;
  std::memset(&P, 0, sizeof P);
  P.det = *det;
  P.model = *mo;
  if (P.model.n_widths > 17) P.model.n_widths = 17;
  P.density = density;
  P.n = n;
  philox_round_keys(seed, P.rk);
  P.first_event = first_event;
  std::vector<void*> fr;
  auto cleanup = [&]() {
    for (void* q : fr) cudaFree(q);
  };
  double *dY = nullptr, *dD = nullptr;
  int32_t* dI = nullptr;
  if (mem_kind == NB200_MEM_HOST) {
    cudaError_t e = cudaMalloc(&dY, n * 48);
    if (e == cudaSuccess) { fr.push_back(dY); e = cudaMalloc(&dI, n * 16); }
    if (e == cudaSuccess) { fr.push_back(dI); e = cudaMalloc(&dD, n * 16); }
    if (e == cudaSuccess) { fr.push_back(dD); e = cudaMemcpyAsync(dY, yields, n * 48, cudaMemcpyHostToDevice, c->stream); }
    if (e != cudaSuccess) {
      cleanup();
      return fail(c, NB200_ENOMEM, cudaGetErrorString(e));
    }
    P.y = dY;
    P.qi = dI;
    P.qd = dD;
  } else {
    P.y = yields;
    P.qi = quanta;
    P.qd = prob_var;
  }
  unsigned grid = unsigned(std::min<uint64_t>((n + 255) / 256, uint64_t(c->sm_count) * 8));
  k_quanta<<<grid, 256, 0, c->stream>>>(P, c->d_err);
  cudaError_t le = cudaGetLastError();
  ++c->launches;
  int rc = 0;
  if (le != cudaSuccess) rc = fail(c, NB200_ECUDA, std::string("k_quanta: ") + cudaGetErrorString(le));
  if (!rc && mem_kind == NB200_MEM_HOST) {
    cudaError_t e = cudaMemcpyAsync(quanta, dI, n * 16, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(prob_var, dD, n * 16, cudaMemcpyDeviceToHost, c->stream);
    if (e != cudaSuccess) rc = fail(c, NB200_ECUDA, cudaGetErrorString(e));
  }
  if (!rc) rc = check_device_error(c);  // synchronises; "Quanta not conserved" -> NB200_EPHYSICS
  cleanup();
  return rc;
}

int nb200_s2(nb200_ctx* c, const nb200_detector* det, const nb200_run_consts* rc, size_t n, const int32_t* Ne,
             const double* pos, const double* dt, const double* field, uint64_t seed, uint64_t first_event, int mem_kind,
             double* out) {
  if (!c) return NB200_EINVAL;
  if (!det || !rc || !Ne || !pos || !dt || !field || !out) return fail(c, NB200_EINVAL, "null argument");
  if (n == 0) return 0;
  NB_CUDA(c, cudaSetDevice(c->device));
  // the per-run constants of the chain (map centre, alias tables of the second photo-electrons, round keys): the source
  // and the model are not used by GetS2
  nb200_model mo;
  nb200_model_default(1, &mo);
  nb200_analysis ana;
  nb200_analysis_default(0, &ana);
  nb200_source src;
  std::memset(&src, 0, sizeof src);
  src.kind = NB200_SRC_FLAT;
  src.species = NB200_NR;
  src.eMin = 1.;
  src.eMax = 2.;
  src.inField = 180.;
  RunParams P;
  int rc_ = build_params(c, det, &mo, &ana, &src, rc, seed, first_event, n, &P);
  if (rc_) return rc_;
  S2Args a;
  a.n = n;
  std::vector<void*> fr;
  auto cleanup = [&]() {
    for (void* q : fr) cudaFree(q);
  };
  if (mem_kind == NB200_MEM_HOST) {
    int32_t* dN = nullptr;
    double *dP = nullptr, *dT = nullptr, *dF = nullptr, *dO = nullptr;
    cudaError_t e = cudaMalloc(&dN, n * 4);
    if (e == cudaSuccess) { fr.push_back(dN); e = cudaMalloc(&dP, n * 40); }
    if (e == cudaSuccess) { fr.push_back(dP); e = cudaMalloc(&dT, n * 8); }
    if (e == cudaSuccess) { fr.push_back(dT); e = cudaMalloc(&dF, n * 8); }
    if (e == cudaSuccess) { fr.push_back(dF); e = cudaMalloc(&dO, n * 72); }
    if (e == cudaSuccess) { fr.push_back(dO); e = cudaMemcpyAsync(dN, Ne, n * 4, cudaMemcpyHostToDevice, c->stream); }
    if (e == cudaSuccess) e = cudaMemcpyAsync(dP, pos, n * 40, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dT, dt, n * 8, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dF, field, n * 8, cudaMemcpyHostToDevice, c->stream);
    if (e != cudaSuccess) {
      cleanup();
      return fail(c, NB200_ENOMEM, cudaGetErrorString(e));
    }
    a.Ne = dN;
    a.pos = dP;
    a.dt = dT;
    a.field = dF;
    a.out = dO;
  } else {
    a.Ne = Ne;
    a.pos = pos;
    a.dt = dt;
    a.field = field;
    a.out = out;
  }
  unsigned grid = unsigned(std::min<uint64_t>((n + 255) / 256, uint64_t(c->sm_count) * 8));
  k_s2<<<grid, 256, 0, c->stream>>>(P, a);
  cudaError_t le = cudaGetLastError();
  ++c->launches;
  int r = 0;
  if (le != cudaSuccess) r = fail(c, NB200_ECUDA, std::string("k_s2: ") + cudaGetErrorString(le));
  if (!r && mem_kind == NB200_MEM_HOST) {
    cudaError_t e = cudaMemcpyAsync(out, a.out, n * 72, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) r = fail(c, NB200_ECUDA, cudaGetErrorString(e));
  }
  cleanup();
  return r;
}

int nb200_s1(nb200_ctx* c, const nb200_detector* det, const nb200_run_consts* rc, int s1mode, size_t n, const int32_t* Nph,
             const double* pos, const double* energy, uint64_t seed, uint64_t first_event, int mem_kind, double* out) {
  if (!c) return NB200_EINVAL;
  if (!det || !rc || !Nph || !pos || !energy || !out) return fail(c, NB200_EINVAL, "null argument");
  if (n == 0) return 0;
  NB_CUDA(c, cudaSetDevice(c->device));
  nb200_model mo;
  nb200_model_default(1, &mo);
  nb200_analysis ana;
  nb200_analysis_default(0, &ana);
  ana.s1mode = s1mode;  // validated below: Full, Parametric or Hybrid (the Waveform mode is not part of this path)
  nb200_source src;
  std::memset(&src, 0, sizeof src);
  src.kind = NB200_SRC_FLAT;
  src.species = NB200_NR;
  src.eMin = 1.;
  src.eMax = 2.;
  src.inField = 180.;
  RunParams P;
  int r = build_params(c, det, &mo, &ana, &src, rc, seed, first_event, n, &P);
  if (r) return r;
  if ((r = ensure_occupancy(c))) return r;
  if ((r = ensure_workspace(c, n, &P.work))) return r;
  HitQueue hq;
  if ((r = ensure_hit_queue(c, &hq))) return r;
  S1Args a;
  a.stride = n;
  a.off = 0;
  std::vector<void*> fr;
  auto cleanup = [&]() {
    for (void* q : fr) cudaFree(q);
  };
  if (mem_kind == NB200_MEM_HOST) {
    int32_t* dN = nullptr;
    double *dP = nullptr, *dE = nullptr, *dO = nullptr;
    cudaError_t e = cudaMalloc(&dN, n * 4);
    if (e == cudaSuccess) { fr.push_back(dN); e = cudaMalloc(&dP, n * 48); }
    if (e == cudaSuccess) { fr.push_back(dP); e = cudaMalloc(&dE, n * 8); }
    if (e == cudaSuccess) { fr.push_back(dE); e = cudaMalloc(&dO, n * 72); }
    if (e == cudaSuccess) { fr.push_back(dO); e = cudaMemcpyAsync(dN, Nph, n * 4, cudaMemcpyHostToDevice, c->stream); }
    if (e == cudaSuccess) e = cudaMemcpyAsync(dP, pos, n * 48, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dE, energy, n * 8, cudaMemcpyHostToDevice, c->stream);
    if (e != cudaSuccess) {
      cleanup();
      return fail(c, NB200_ENOMEM, cudaGetErrorString(e));
    }
    a.Nph = dN;
    a.pos = dP;
    a.energy = dE;
    a.out = dO;
  } else {
    a.Nph = Nph;
    a.pos = pos;
    a.energy = energy;
    a.out = out;
  }
  const uint64_t total = n, cap = c->work_cap, sm = uint64_t(c->sm_count);
  for (uint64_t off = 0; off < total && !r; off += cap) {
    const uint64_t m = std::min<uint64_t>(cap, total - off);
    RunParams Q = P;
    Q.n_events = m;
    Q.chunk_off = off;
    a.off = off;
    const unsigned g1 = unsigned(std::min<uint64_t>((m + 255) / 256, sm * 8));
    k_s1pre<<<g1, 256, 0, c->stream>>>(Q, a);
    launch_hits(c, Q, m, hq);
    k_s1post<<<g1, 256, 0, c->stream>>>(Q, a);
    cudaError_t le = cudaGetLastError();
    c->launches += 4;
    if (le != cudaSuccess) r = fail(c, NB200_ECUDA, std::string("nb200_s1: ") + cudaGetErrorString(le));
  }
  if (!r && mem_kind == NB200_MEM_HOST) {
    cudaError_t e = cudaMemcpyAsync(out, a.out, n * 72, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) r = fail(c, NB200_ECUDA, cudaGetErrorString(e));
  }
  cleanup();
  return r;
}

// ---- the remaining NESTcalc model methods as stage calls -------------------------------------------------------

int nb200_widths(nb200_ctx* c, const nb200_detector* det, const nb200_model* mo, size_t n, const double* efield,
                 const double* elecFrac, const double* numQuanta, double density, int mem_kind, double* out) {
  if (!c) return NB200_EINVAL;
  if (!det || !mo || !efield || !elecFrac || !numQuanta || !out) return fail(c, NB200_EINVAL, "null argument");
  if (mo->n_widths < 13)
    return fail(c, NB200_EPHYSICS, "ERROR: You need a minimum of 13 free parameters for the resolution model.");
  if (mo->NRERWidthsParam[9] > mo->NRERWidthsParam[8])  // RecombOmegaER's throw, NEST.cpp:183-184
    return fail(c, NB200_EPHYSICS,
                "Low-E r-fluct Gauss peak's width greater than centroid. This is physically wrong so "
                "please check your inputs, in NRERWidthsParam: they must be in the old wrong order.");
  if (n == 0) return 0;
  NB_CUDA(c, cudaSetDevice(c->device));
  Staged st(c, mem_kind);
  WidthsParams P;
  std::memset(&P, 0, sizeof P);
  P.model = *mo;
  P.inGas = det->inGas;
  P.density = density;
  P.n = n;
  P.efield = st.in(efield, n);
  P.elecFrac = st.in(elecFrac, n);
  P.nq = st.in(numQuanta, n);
  P.out = st.out(out, 3 * n);
  if (st.err == cudaSuccess) {
    k_widths<<<unsigned(std::min<uint64_t>((n + 255) / 256, uint64_t(c->sm_count) * 8)), 256, 0, c->stream>>>(P);
    ++c->launches;
  }
  return st.finish(0, "k_widths");
}

namespace {
// RunParams for the stage calls that need the detector and the per-run constants but no source / analysis
int stage_params(nb200_ctx* c, const nb200_detector* det, const nb200_run_consts* rc, uint64_t seed, uint64_t first_event,
                 uint64_t n, RunParams* P) {
  nb200_model mo;
  nb200_model_default(1, &mo);
  nb200_analysis ana;
  nb200_analysis_default(0, &ana);
  nb200_source src;
  std::memset(&src, 0, sizeof src);
  src.kind = NB200_SRC_MONO;
  src.species = NB200_beta;
  src.eMin = src.eMax = 1.;
  src.fixed_pos = 1;
  src.inField = rc->field >= 0. ? rc->field : 0.;
  return build_params(c, det, &mo, &ana, &src, rc, seed, first_event, n, P);
}
}  // namespace

int nb200_spike(nb200_ctx* c, const nb200_detector* det, const nb200_run_consts* rc, size_t n, const double* pos,
                const double* scint, uint64_t seed, uint64_t first_event, int mem_kind, double* out) {
  if (!c) return NB200_EINVAL;
  if (!det || !rc || !pos || !scint || !out) return fail(c, NB200_EINVAL, "null argument");
  if (n == 0) return 0;
  NB_CUDA(c, cudaSetDevice(c->device));
  RunParams P;
  int r = stage_params(c, det, rc, seed, first_event, n, &P);
  if (r) return r;
  Staged st(c, mem_kind);
  SpikeArgs a;
  a.n = n;
  a.pos = st.in(pos, 3 * n);
  a.scint = st.in(scint, 9 * n);
  a.out = st.out(out, 2 * n);
  if (st.err == cudaSuccess) {
    k_spike<<<unsigned(std::min<uint64_t>((n + 255) / 256, uint64_t(c->sm_count) * 8)), 256, 0, c->stream>>>(P, a);
    ++c->launches;
  }
  return st.finish(0, "k_spike");
}

int nb200_xy_resolution(nb200_ctx* c, const nb200_detector* det, size_t n, const double* x, const double* y,
                        const double* A_top, uint64_t seed, uint64_t first_event, int mem_kind, double* out) {
  if (!c) return NB200_EINVAL;
  if (!det || !x || !y || !A_top || !out) return fail(c, NB200_EINVAL, "null argument");
  if (n == 0) return 0;
  NB_CUDA(c, cudaSetDevice(c->device));
  RunParams P;
  std::memset(&P, 0, sizeof P);
  P.det = *det;
  int r;
  if ((r = upload_map(c, &P.det.FitTBA))) return r;
  P.det.FitS1.lut = P.det.FitS2.lut = P.det.FitEF.lut = nullptr;  // not read by xyResolution
  P.seed = seed;
  philox_round_keys(seed, P.rk);
  P.first_event = first_event;
  Staged st(c, mem_kind);
  XyArgs a;
  a.n = n;
  a.x = st.in(x, n);
  a.y = st.in(y, n);
  a.A_top = st.in(A_top, n);
  a.out = st.out(out, 2 * n);
  if (st.err == cudaSuccess) {
    k_xyres<<<unsigned(std::min<uint64_t>((n + 255) / 256, uint64_t(c->sm_count) * 8)), 256, 0, c->stream>>>(P, a);
    ++c->launches;
  }
  return st.finish(0, "k_xyres");
}

int nb200_spectrum(nb200_ctx* c, const nb200_source* src, size_t n, uint64_t seed, uint64_t first_event, int mem_kind,
                   double* out) {
  if (!c) return NB200_EINVAL;
  if (!src || !out) return fail(c, NB200_EINVAL, "null argument");
  if (src->kind < NB200_SRC_MONO || src->kind > NB200_SRC_ATMNU || src->kind == NB200_SRC_LIST)
    return fail(c, NB200_EINVAL, "nb200_spectrum: not a sampled source kind");
  if (n == 0) return 0;
  NB_CUDA(c, cudaSetDevice(c->device));
  c->call_memo.clear();
  RunParams P;
  std::memset(&P, 0, sizeof P);
  P.src = *src;
  int r = upload_wimp_source(c, src, &P);
  if (r) return r;
  P.seed = seed;
  philox_round_keys(seed, P.rk);
  P.first_event = first_event;
  P.n_events = n;
  P.species = src->species;
  Staged st(c, mem_kind);
  double* dout = st.out(out, n);
  if (st.err == cudaSuccess) {
    k_spectrum<<<unsigned(std::min<uint64_t>((n + 255) / 256, uint64_t(c->sm_count) * 8)), 256, 0, c->stream>>>(P, dout);
    ++c->launches;
  }
  return st.finish(0, "k_spectrum");
}

// ---- the chain ---------------------------------------------------------------------------

size_t nb200_band_hist_bytes(const nb200_analysis* a) { return a ? band_word_count(a) * 8 : 0; }
double nb200_band_scale_x(const nb200_analysis* a) { return a ? band_scale_x(a) : 0.; }

int nb200_band_reset(nb200_ctx* c, const nb200_analysis* ana) {
  if (!c || !ana) return NB200_EINVAL;
  if (ana->numBins < 1 || ana->numBins > 1000 || ana->logBins < 0) return fail(c, NB200_EINVAL, "bad band geometry");
  NB_CUDA(c, cudaSetDevice(c->device));
  size_t w = band_word_count(ana);
  if (w != c->band_words || c->band_external) {
    if (c->d_band && !c->band_external) cudaFree(c->d_band);
    c->d_band = nullptr;
    c->band_external = false;
    NB_CUDA(c, cudaMalloc(&c->d_band, w * 8));
    c->band_words = w;
  }
  c->band_numBins = ana->numBins;
  c->band_logBins = ana->logBins;
  c->band_events = 0;
  NB_CUDA(c, cudaMemsetAsync(c->d_band, 0, w * 8, c->stream));
  return 0;
}

int nb200_band_attach(nb200_ctx* c, const nb200_analysis* ana, void* buf) {
  if (!c || !ana || !buf) return NB200_EINVAL;
  if (ana->numBins < 1 || ana->numBins > 1000 || ana->logBins < 0) return fail(c, NB200_EINVAL, "bad band geometry");
  if (c->d_band && !c->band_external) cudaFree(c->d_band);
  c->d_band = static_cast<unsigned long long*>(buf);
  c->band_external = true;
  c->band_words = band_word_count(ana);
  c->band_numBins = ana->numBins;
  c->band_logBins = ana->logBins;
  c->band_events = 0;
  return 0;
}

int nb200_run_async(nb200_ctx* c, const nb200_detector* det, const nb200_model* mo, const nb200_analysis* ana,
                    const nb200_source* src, const nb200_run_consts* rc, uint64_t seed, uint64_t first_event,
                    uint64_t n_events, nb200_soa_out* soa) {
  if (!c) return NB200_EINVAL;
  NB_CUDA(c, cudaSetDevice(c->device));
  RunParams P;
  int r = build_params(c, det, mo, ana, src, rc, seed, first_event, n_events, &P);
  if (r) return r;
  if (src->kind == NB200_SRC_LIST && src->list_mem_kind != NB200_MEM_DEVICE)
    return fail(c, NB200_EINVAL, "nb200_run_async needs a device-resident energy list");
  if (soa) {
    if (soa->mem_kind != NB200_MEM_DEVICE) return fail(c, NB200_EINVAL, "nb200_run_async needs device SoA pointers");
    P.write_soa = (soa->flags & NB200_SOA_F32) ? 2 : 1;
    static_assert(sizeof(SoaDev) == sizeof(nb200_soa_out) - 8, "SoA layouts must match");
    std::memcpy(&P.soa, reinterpret_cast<const char*>(soa) + 8, sizeof(SoaDev));
  }
  if (c->d_band && !src->vec_mode) {
    if (c->band_numBins != ana->numBins || c->band_logBins != ana->logBins)
      return fail(c, NB200_EINVAL, "band accumulator geometry differs from the analysis (call nb200_band_reset)");
    P.do_band = 1;
    P.band = c->d_band;
  }
  r = launch_event(c, P);
  if (!r && P.do_band) c->band_events += n_events;
  return r;
}

int nb200_band_fetch(nb200_ctx* c, nb200_band_hist* h) {
  if (!c || !h) return NB200_EINVAL;
  if (!c->d_band) return fail(c, NB200_EINVAL, "no device band accumulator (call nb200_band_reset)");
  if (h->numBins != c->band_numBins || h->logBins != c->band_logBins) return fail(c, NB200_EINVAL, "band geometry mismatch");
  std::vector<unsigned long long> w(c->band_words);
  NB_CUDA(c, cudaMemcpyAsync(w.data(), c->d_band, c->band_words * 8, cudaMemcpyDeviceToHost, c->stream));
  int rc = check_device_error(c);  // synchronises
  if (rc) return rc;
  for (int j = 0; j < h->numBins; ++j) {
    const unsigned long long* b = &w[size_t(j) * BAND_WORDS];
    if (h->accepted) h->accepted[j] += b[0];
    if (h->rejected) h->rejected[j] += b[1];
    if (h->sumS1) h->sumS1[j] += (int64_t)b[2];
    if (h->sumY) h->sumY[j] += (int64_t)b[3];
    if (h->sumY2) h->sumY2[j] += (int64_t)b[4];
    if (h->sumY3) h->sumY3[j] += (int64_t)b[5];
  }
  if (h->hist2d)
    for (size_t i = 0; i < size_t(h->numBins) * size_t(h->logBins); ++i)
      h->hist2d[i] += w[size_t(h->numBins) * BAND_WORDS + i];
  h->n_events += c->band_events;
  return 0;
}

int nb200_band_device_buffer(nb200_ctx* c, void** dptr, size_t* n_u64) {
  if (!c || !dptr || !n_u64) return NB200_EINVAL;
  *dptr = c->d_band;
  *n_u64 = c->band_words;
  return c->d_band ? 0 : fail(c, NB200_EINVAL, "no device band accumulator");
}

int nb200_allreduce_hist(nb200_ctx* c, void* comm) {
  if (!c || !comm) return NB200_EINVAL;
  if (!c->d_band) return fail(c, NB200_EINVAL, "no device band accumulator");
  // ncclAllReduce(sendbuff, recvbuff, count, ncclUint64 = 5, ncclSum = 0, comm, stream)
  typedef int (*allreduce_fn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
  static allreduce_fn fn = nullptr;
  if (!fn) {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW);
    if (!h) return fail(c, NB200_EINVAL, std::string("libnccl.so.2 not loadable: ") + dlerror());
    fn = reinterpret_cast<allreduce_fn>(dlsym(h, "ncclAllReduce"));
    if (!fn) return fail(c, NB200_EINVAL, "ncclAllReduce not found");
  }
  int r = fn(c->d_band, c->d_band, c->band_words, 5, 0, comm, c->stream);
  if (r != 0) return fail(c, NB200_ECUDA, "ncclAllReduce failed with code " + std::to_string(r));
  return 0;
}

int nb200_run(nb200_ctx* c, const nb200_detector* det, const nb200_model* mo, const nb200_analysis* ana,
              const nb200_source* src, const nb200_run_consts* rc, uint64_t seed, uint64_t first_event,
              uint64_t n_events, nb200_soa_out* soa, nb200_band_hist* hist) {
  if (!c) return NB200_EINVAL;
  NB_CUDA(c, cudaSetDevice(c->device));
  int r = validate(c, det, mo, ana, src);
  if (r) return r;
  if (hist && (hist->numBins != ana->numBins || hist->logBins != ana->logBins))
    return fail(c, NB200_EINVAL, "band geometry mismatch");
  if (n_events == 0) return 0;
  Trace tr("nb200_run");
  const bool host_soa = soa && soa->mem_kind == NB200_MEM_HOST;
  const bool f32 = soa && (soa->flags & NB200_SOA_F32);
  const bool host_list = src->kind == NB200_SRC_LIST && src->list_mem_kind == NB200_MEM_HOST;
  // Host outputs: the kernels of chunk i write into device staging buffer i%2 while the copy
  // stream drains buffer (i-1)%2 to the caller's arrays, so PCIe and compute overlap.
  // host records: smaller chunks shorten the pipeline fill (first chunk's kernels are not overlapped with
  // copies); NB200_HOST_CHUNK_LOG2 overrides for tuning
  int host_log2 = 20;  // measured on B200: 2^20 gives 5.06e8 events/s end to end, 2^21 4.93e8, 2^18 4.82e8
  if (const char* e = std::getenv("NB200_HOST_CHUNK_LOG2")) host_log2 = std::min(22, std::max(16, atoi(e)));
  const uint64_t chunk_max = host_soa ? (1ull << host_log2) : chunk_events(c);
  const uint64_t cap = std::min<uint64_t>(n_events, chunk_max);
  const size_t ncol = (sizeof(nb200_soa_out) - 8) / sizeof(void*);
  struct Col {
    size_t k, elt, off;
    void* host;
  };
  std::vector<Col> cols;
  size_t per_chunk = 0;
  if (host_soa) {
    void** hp = reinterpret_cast<void**>(reinterpret_cast<char*>(soa) + 8);
    for (size_t k = 0; k < ncol; ++k) {
      if (!hp[k]) continue;
      const size_t elt = ((k >= 8 && k < 12) || f32) ? 4 : 8;
      cols.push_back({k, elt, per_chunk, hp[k]});
      per_chunk += ((elt * cap + 255) / 256) * 256;
    }
    if (per_chunk > c->stage_bytes) {
      for (int b = 0; b < 2; ++b) {
        if (c->d_stage[b]) cudaFree(c->d_stage[b]);
        c->d_stage[b] = nullptr;
      }
      c->stage_bytes = 0;
      for (int b = 0; b < 2; ++b) {
        cudaError_t e = cudaMalloc(&c->d_stage[b], per_chunk);
        if (e != cudaSuccess) return fail(c, NB200_ENOMEM, cudaGetErrorString(e));
      }
      c->stage_bytes = per_chunk;
    }
    if (!c->copy_stream) {
      NB_CUDA(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
      for (int b = 0; b < 2; ++b) {
        NB_CUDA(c, cudaEventCreateWithFlags(&c->ev_done[b], cudaEventDisableTiming));
        NB_CUDA(c, cudaEventCreateWithFlags(&c->ev_copied[b], cudaEventDisableTiming));
      }
    }
  }
  double *dE = nullptr, *dx = nullptr, *dy = nullptr, *dz = nullptr;
  if (host_list) {
    const size_t need = 8 * cap * (src->x ? 4 : 1);
    if (need > c->list_bytes) {
      if (c->d_list) cudaFree(c->d_list);
      c->d_list = nullptr;
      c->list_bytes = 0;
      cudaError_t e = cudaMalloc(&c->d_list, need);
      if (e != cudaSuccess) return fail(c, NB200_ENOMEM, "device staging for the energy list");
      c->list_bytes = need;
    }
    dE = static_cast<double*>(c->d_list);
    if (src->x) {
      dx = dE + cap;
      dy = dx + cap;
      dz = dy + cap;
    }
  }
  if (hist) {
    size_t w = band_word_count(ana);
    if (w != c->band_tmp_words) {
      if (c->d_band_tmp) cudaFree(c->d_band_tmp);
      c->d_band_tmp = nullptr;
      cudaError_t e = cudaMalloc(&c->d_band_tmp, w * 8);
      if (e != cudaSuccess) return fail(c, NB200_ENOMEM, cudaGetErrorString(e));
      c->band_tmp_words = w;
    }
    cudaMemsetAsync(c->d_band_tmp, 0, w * 8, c->stream);
  }

  tr.mark("staging");
  RunParams P0;
  int rcode = build_params(c, det, mo, ana, src, rc, seed, first_event, n_events, &P0);
  if (rcode) return rcode;
  uint64_t ichunk = 0;
  for (uint64_t done = 0; done < n_events && !rcode; done += chunk_max, ++ichunk) {
    const uint64_t m = std::min<uint64_t>(chunk_max, n_events - done);
    const int buf = int(ichunk & 1);
    nb200_source s = *src;
    if (src->kind == NB200_SRC_LIST) {
      if (host_list) {
        cudaMemcpyAsync(dE, src->E + done, 8 * m, cudaMemcpyHostToDevice, c->stream);
        s.E = dE;
        if (src->x) {
          cudaMemcpyAsync(dx, src->x + done, 8 * m, cudaMemcpyHostToDevice, c->stream);
          cudaMemcpyAsync(dy, src->y + done, 8 * m, cudaMemcpyHostToDevice, c->stream);
          cudaMemcpyAsync(dz, src->z + done, 8 * m, cudaMemcpyHostToDevice, c->stream);
          s.x = dx;
          s.y = dy;
          s.z = dz;
        }
      } else {
        s.E = src->E + done;
        if (src->x) {
          s.x = src->x + done;
          s.y = src->y + done;
          s.z = src->z + done;
        }
      }
    }
    RunParams P = P0;  // built once for the call; per chunk only the event range and the list pointers move
    P.first_event = first_event + done;
    P.n_events = m;
    P.src.E = s.E;
    P.src.x = s.x;
    P.src.y = s.y;
    P.src.z = s.z;
    if (soa) {
      P.write_soa = f32 ? 2 : 1;
      nb200_soa_out tmp = *soa;
      void** dp = reinterpret_cast<void**>(reinterpret_cast<char*>(&tmp) + 8);
      if (host_soa) {
        for (size_t k = 0; k < ncol; ++k) dp[k] = nullptr;
        for (auto& col : cols) dp[col.k] = static_cast<char*>(c->d_stage[buf]) + col.off;
        if (ichunk >= 2) cudaStreamWaitEvent(c->stream, c->ev_copied[buf], 0);  // buffer drained?
      } else {  // caller's device columns, advanced to this chunk
        for (size_t k = 0; k < ncol; ++k)
          if (dp[k]) dp[k] = static_cast<char*>(dp[k]) + (((k >= 8 && k < 12) || f32) ? 4 : 8) * done;
      }
      std::memcpy(&P.soa, reinterpret_cast<const char*>(&tmp) + 8, sizeof(SoaDev));
    }
    if (hist && !src->vec_mode) {
      P.do_band = 1;
      P.band = c->d_band_tmp;
    }
    rcode = launch_event(c, P);
    if (rcode) break;
    if (host_soa) {
      cudaEventRecord(c->ev_done[buf], c->stream);
      cudaStreamWaitEvent(c->copy_stream, c->ev_done[buf], 0);
      for (auto& col : cols)
        cudaMemcpyAsync(static_cast<char*>(col.host) + col.elt * done, static_cast<char*>(c->d_stage[buf]) + col.off,
                        col.elt * m, cudaMemcpyDeviceToHost, c->copy_stream);
      cudaEventRecord(c->ev_copied[buf], c->copy_stream);
    }
    tr.mark("chunk enqueued");
  }
  if (host_soa && c->copy_stream) {
    cudaError_t e = cudaStreamSynchronize(c->copy_stream);
    if (e != cudaSuccess && !rcode) rcode = fail(c, NB200_ECUDA, cudaGetErrorString(e));
  }
  if (!rcode) rcode = check_device_error(c);
  tr.mark("kernels + copies done");
  if (!rcode && hist) {
    std::vector<unsigned long long> w(c->band_tmp_words);
    cudaError_t e = cudaMemcpyAsync(w.data(), c->d_band_tmp, c->band_tmp_words * 8, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess)
      rcode = fail(c, NB200_ECUDA, cudaGetErrorString(e));
    else {
      for (int j = 0; j < hist->numBins; ++j) {
        const unsigned long long* b = &w[size_t(j) * BAND_WORDS];
        if (hist->accepted) hist->accepted[j] += b[0];
        if (hist->rejected) hist->rejected[j] += b[1];
        if (hist->sumS1) hist->sumS1[j] += (int64_t)b[2];
        if (hist->sumY) hist->sumY[j] += (int64_t)b[3];
        if (hist->sumY2) hist->sumY2[j] += (int64_t)b[4];
        if (hist->sumY3) hist->sumY3[j] += (int64_t)b[5];
      }
      if (hist->hist2d)
        for (size_t i = 0; i < size_t(hist->numBins) * size_t(hist->logBins); ++i)
          hist->hist2d[i] += w[size_t(hist->numBins) * BAND_WORDS + i];
      hist->n_events += n_events;
    }
  }
  cudaStreamSynchronize(c->stream);
  tr.mark("band fetch");
  return rcode;
}

// Batched parameter sweep: see SweepArgs in nb_kernels.cuh
int nb200_run_sweep(nb200_ctx* c, size_t n_points, const nb200_detector* dets, const nb200_model* models,
                    const nb200_analysis* ana, const nb200_source* srcs, const nb200_run_consts* rcs, uint64_t seed,
                    uint64_t first_event, uint64_t n_events, nb200_band_hist* bands) {
  if (!c) return NB200_EINVAL;
  if (!dets || !models || !ana || !srcs || !rcs || !bands) return fail(c, NB200_EINVAL, "null argument");
  if (n_points == 0 || n_events == 0) return 0;
  if (n_points > (1u << 24)) return fail(c, NB200_EINVAL, "too many grid points");
  NB_CUDA(c, cudaSetDevice(c->device));
  Trace tr("nb200_run_sweep");
  std::vector<RunParams> Ps(n_points);
  for (size_t p = 0; p < n_points; ++p) {
    if (srcs[p].kind == NB200_SRC_LIST || srcs[p].vec_mode)
      return fail(c, NB200_EINVAL, "nb200_run_sweep: list sources / runNESTvec semantics are per-point calls (nb200_run)");
    if (bands[p].numBins != ana->numBins || bands[p].logBins != ana->logBins)
      return fail(c, NB200_EINVAL, "band geometry mismatch");
    int r = build_params(c, &dets[p], &models[p], ana, &srcs[p], &rcs[p], seed, first_event, n_events, &Ps[p], p != 0);
    if (r) return r;
    Ps[p].do_band = 1;
  }
  tr.mark("per-point parameters");
  const size_t words = band_word_count(ana);
  size_t smem = size_t(ana->numBins) * BAND_WORDS * 8 + size_t(ana->numBins) * size_t(ana->logBins) * 4;
  int band_in_smem = 1;
  if (smem > 150 * 1024) {
    band_in_smem = 0;
    smem = 0;
  }
  if (!c->occ_pre_sweep) {
    NB_CUDA(c, cudaFuncSetAttribute(k_post_sweep, cudaFuncAttributeMaxDynamicSharedMemorySize, 150 * 1024));
    NB_CUDA(c, cudaFuncSetAttribute(k_hits_sweep, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kHitDynSmem)));
    NB_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c->occ_pre_sweep, k_pre_sweep, kPreBlock, 0));
    NB_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c->occ_hits_sweep, k_hits_sweep, kHitBlock, kHitDynSmem));
    if (c->occ_pre_sweep < 1) c->occ_pre_sweep = 1;
    if (c->occ_hits_sweep < 1) c->occ_hits_sweep = 1;
  }
  int occ_post = 1;
  NB_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_post, k_post_sweep, kPostBlock, smem));
  if (occ_post < 1) occ_post = 1;
  const size_t pbytes = sizeof(RunParams) * n_points;
  if (pbytes > c->sweep_params_bytes) {
    if (c->d_sweep_params) cudaFree(c->d_sweep_params);
    c->d_sweep_params = nullptr;
    c->sweep_params_bytes = 0;
    NB_CUDA(c, cudaMalloc(&c->d_sweep_params, pbytes));
    c->sweep_params_bytes = pbytes;
  }
  if (words * n_points > c->sweep_band_words) {
    if (c->d_sweep_bands) cudaFree(c->d_sweep_bands);
    c->d_sweep_bands = nullptr;
    c->sweep_band_words = 0;
    NB_CUDA(c, cudaMalloc(&c->d_sweep_bands, words * n_points * 8));
    c->sweep_band_words = words * n_points;
  }
  NB_CUDA(c, cudaMemcpyAsync(c->d_sweep_params, Ps.data(), pbytes, cudaMemcpyHostToDevice, c->stream));
  NB_CUDA(c, cudaMemsetAsync(c->d_sweep_bands, 0, words * n_points * 8, c->stream));
  SweepArgs A;
  std::memset(&A, 0, sizeof A);
  A.params = static_cast<const RunParams*>(c->d_sweep_params);
  A.bands = c->d_sweep_bands;
  A.band_words = words;
  A.n_per_point = n_events;
  A.n_pad = (n_events + 1023) / 1024 * 1024;
  A.first_event = first_event;
  philox_round_keys(seed, A.rk);
  const uint64_t total = uint64_t(n_points) * A.n_pad;
  int r = ensure_workspace(c, total, &A.work);
  if (r) return r;
  HitQueue hq;
  if ((r = ensure_hit_queue(c, &hq))) return r;
  const uint64_t cap = c->work_cap / 1024 * 1024;
  const uint64_t sm = uint64_t(c->sm_count);
  for (uint64_t off = 0; off < total; off += cap) {
    const uint64_t m = std::min<uint64_t>(cap, total - off);
    A.slot0 = off;
    A.n_slots = m;
    unsigned g_pre = unsigned(std::min<uint64_t>((m + kPreBlock - 1) / kPreBlock, sm * c->occ_pre_sweep * 8));
    unsigned g_hits = unsigned(std::min<uint64_t>((m + kHitGroup - 1) / kHitGroup, sm * c->occ_hits_sweep));
    unsigned g_post = unsigned(std::min<uint64_t>((m + kPostBlock - 1) / kPostBlock, sm * occ_post));
    k_pre_sweep<<<g_pre, kPreBlock, 0, c->stream>>>(A, c->d_err);
    cudaMemsetAsync(c->d_err + 1, 0, 2 * sizeof(int), c->stream);
    k_hits_sweep<<<g_hits, kHitBlock, kHitDynSmem, c->stream>>>(A, reinterpret_cast<unsigned int*>(c->d_err + 1), hq);
    k_hits_finish_sweep<<<unsigned(sm * NB_FIN_MINB * 2), kFinBlock, 0, c->stream>>>(A, hq);
    k_post_sweep<<<g_post, kPostBlock, smem, c->stream>>>(A, band_in_smem, c->d_err);
    NB_CUDA(c, cudaGetLastError());
    c->launches += 4;
  }
  tr.mark("launched");
  c->sweep_last_points = n_points;
  c->sweep_last_words = words;
  // the bands come back through a pinned buffer kept with the context (7 MB for 250 LUX bands: 0.1 ms instead of
  // the 0.7 ms of a pageable copy, 3 % of such a sweep)
  if (words * n_points > c->h_sweep_words) {
    if (c->h_sweep_bands) cudaFreeHost(c->h_sweep_bands);
    c->h_sweep_bands = nullptr;
    c->h_sweep_words = 0;
    NB_CUDA(c, cudaMallocHost(&c->h_sweep_bands, words * n_points * 8));
    c->h_sweep_words = words * n_points;
  }
  const unsigned long long* w = c->h_sweep_bands;
  NB_CUDA(c, cudaMemcpyAsync(c->h_sweep_bands, c->d_sweep_bands, words * n_points * 8, cudaMemcpyDeviceToHost, c->stream));
  r = check_device_error(c);  // synchronises (Ps must also outlive the copy above)
  if (r) return r;
  tr.mark("kernels done");
  for (size_t p = 0; p < n_points; ++p) {
    nb200_band_hist* h = &bands[p];
    const unsigned long long* wp = &w[p * words];
    for (int j = 0; j < h->numBins; ++j) {
      const unsigned long long* b = wp + size_t(j) * BAND_WORDS;
      if (h->accepted) h->accepted[j] += b[0];
      if (h->rejected) h->rejected[j] += b[1];
      if (h->sumS1) h->sumS1[j] += (int64_t)b[2];
      if (h->sumY) h->sumY[j] += (int64_t)b[3];
      if (h->sumY2) h->sumY2[j] += (int64_t)b[4];
      if (h->sumY3) h->sumY3[j] += (int64_t)b[5];
    }
    if (h->hist2d)
      for (size_t i = 0; i < size_t(h->numBins) * size_t(h->logBins); ++i)
        h->hist2d[i] += wp[size_t(h->numBins) * BAND_WORDS + i];
    h->n_events += n_events;
  }
  return 0;
}

// Band post-processing of the last sweep's bands on the device (k_band_post)
int nb200_sweep_band_post(nb200_ctx* c, size_t n_points, const nb200_analysis* ana, const double* target_band,
                          int free_param, const double* centroid, double* stats, double* gof, double* leak) {
  if (!c || !ana) return NB200_EINVAL;
  if (n_points == 0) return 0;
  const size_t words = band_word_count(ana);
  if (!c->d_sweep_bands || words * n_points > c->sweep_band_words || c->sweep_last_points != n_points ||
      c->sweep_last_words != words)
    return fail(c, NB200_EINVAL, "nb200_sweep_band_post: no sweep of this shape on this context (call nb200_run_sweep first)");
  if ((gof && !target_band) || (leak && !centroid)) return fail(c, NB200_EINVAL, "gof needs a target band, leak a centroid line");
  if (ana->numBins > 256) return fail(c, NB200_EINVAL, "nb200_sweep_band_post: at most 256 S1 bins");
  NB_CUDA(c, cudaSetDevice(c->device));
  Staged st(c, NB200_MEM_HOST);
  BandPostArgs A;
  std::memset(&A, 0, sizeof A);
  A.bands = c->d_sweep_bands;
  A.band_words = words;
  A.numBins = ana->numBins;
  A.logBins = ana->logBins;
  A.useS2 = ana->useS2;
  A.free_param = free_param;
  if (ana->useS2 == 2) {
    A.binWidth = (ana->maxS2 - ana->minS2) / double(ana->numBins);
    A.border = ana->minS2;
  } else {
    A.binWidth = (ana->maxS1 - ana->minS1) / double(ana->numBins);
    A.border = ana->minS1;
  }
  A.fx_x = band_scale_x(ana);
  A.logMin = ana->logMin;
  A.logMax = ana->logMax;
  A.target = st.in(target_band, size_t(ana->numBins) * 7);
  A.centroid = st.in(centroid, size_t(ana->numBins));
  A.stats = st.out(stats, n_points * size_t(ana->numBins) * 7);
  A.gof = st.out(gof, n_points * 6);
  A.leak = st.out(leak, n_points * 3);
  if (st.err == cudaSuccess) {
    k_band_post<<<unsigned(n_points), 256, 0, c->stream>>>(A);
    ++c->launches;
  }
  return st.finish(0, "k_band_post");
}

// Per-bin fits of device-resident bands (k_band_fit): the last sweep's, or the context's own accumulator
static int band_fit_device(nb200_ctx* c, const unsigned long long* bands, size_t n_points, const nb200_analysis* ana,
                           int model, const double* line, double* fit, double* leak) {
  if (model != 0 && model != 1) return fail(c, NB200_EINVAL, "band fit model must be 0 (Gaussian) or 1 (skew-Gaussian)");
  if (!fit) return fail(c, NB200_EINVAL, "null argument");
  if (leak && !line) return fail(c, NB200_EINVAL, "the leakage needs a line");
  if (ana->logBins < 4) return fail(c, NB200_EINVAL, "band fits need logBins >= 4");
  NB_CUDA(c, cudaSetDevice(c->device));
  // results leave through a scratch buffer the context keeps (a cudaMalloc / cudaFree pair per call cost more than the
  // kernel: 7 of 9.4 ms for a 250-point sweep)
  const size_t n_rows = n_points * size_t(ana->numBins), cols = model == 0 ? 8 : 12;
  const size_t need = (n_rows * (cols + 1) + size_t(ana->numBins)) * sizeof(double);
  if (need > c->fit_scratch_bytes) {
    if (c->d_fit_scratch) cudaFree(c->d_fit_scratch);
    c->d_fit_scratch = nullptr;
    c->fit_scratch_bytes = 0;
    NB_CUDA(c, cudaMalloc(&c->d_fit_scratch, need));
    c->fit_scratch_bytes = need;
  }
  double* d_fit = static_cast<double*>(c->d_fit_scratch);
  double* d_leak = d_fit + n_rows * cols;
  double* d_line = d_leak + n_rows;
  if (line) NB_CUDA(c, cudaMemcpyAsync(d_line, line, size_t(ana->numBins) * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  BandFitArgs A;
  std::memset(&A, 0, sizeof A);
  A.bands = bands;
  A.band_words = band_word_count(ana);
  A.numBins = ana->numBins;
  A.logBins = ana->logBins;
  A.model = model;
  A.cols = model == 0 ? 8 : 12;
  A.n_rows = uint64_t(n_points) * uint64_t(ana->numBins);
  A.logMin = ana->logMin;
  A.logMax = ana->logMax;
  A.line = line ? d_line : nullptr;
  A.fit = d_fit;
  A.leak = leak ? d_leak : nullptr;
  k_band_fit<<<unsigned((A.n_rows + kBandFitRows - 1) / kBandFitRows), kBandFitRows * 32, 0, c->stream>>>(A);
  cudaError_t le = cudaGetLastError();
  ++c->launches;
  if (le != cudaSuccess) return fail(c, NB200_ECUDA, std::string("k_band_fit: ") + cudaGetErrorString(le));
  NB_CUDA(c, cudaMemcpyAsync(fit, d_fit, n_rows * cols * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  if (leak) NB_CUDA(c, cudaMemcpyAsync(leak, d_leak, n_rows * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  return check_device_error(c);  // synchronises
}
int nb200_sweep_band_fit(nb200_ctx* c, size_t n_points, const nb200_analysis* ana, int model, const double* line,
                         double* fit, double* leak) {
  if (!c || !ana) return NB200_EINVAL;
  if (n_points == 0) return 0;
  const size_t words = band_word_count(ana);
  if (!c->d_sweep_bands || words * n_points > c->sweep_band_words || c->sweep_last_points != n_points ||
      c->sweep_last_words != words)
    return fail(c, NB200_EINVAL, "nb200_sweep_band_fit: no sweep of this shape on this context (call nb200_run_sweep first)");
  return band_fit_device(c, c->d_sweep_bands, n_points, ana, model, line, fit, leak);
}
int nb200_band_fit_device(nb200_ctx* c, const nb200_analysis* ana, int model, const double* line, double* fit,
                          double* leak) {
  if (!c || !ana) return NB200_EINVAL;
  if (!c->d_band) return fail(c, NB200_EINVAL, "no device band accumulator (call nb200_band_reset)");
  if (ana->numBins != c->band_numBins || ana->logBins != c->band_logBins || band_word_count(ana) != c->band_words)
    return fail(c, NB200_EINVAL, "band geometry mismatch");
  return band_fit_device(c, c->d_band, 1, ana, model, line, fit, leak);
}

// GetBand's closing statistics (execNEST.cpp:1582-1618) from the integer accumulator
int nb200_band_finalize(const nb200_analysis* ana, const nb200_band_hist* h, double* band) {
  if (!ana || !h || !band) return NB200_EINVAL;
  if (!h->accepted || !h->rejected || !h->sumS1 || !h->sumY || !h->sumY2 || !h->sumY3) return NB200_EINVAL;
  double binWidth, border;
  if (ana->useS2 == 2) {
    binWidth = (ana->maxS2 - ana->minS2) / double(ana->numBins);
    border = ana->minS2;
  } else {
    binWidth = (ana->maxS1 - ana->minS1) / double(ana->numBins);
    border = ana->minS1;
  }
  // The sums are 64-bit integers: refuse a bin that holds so many events that one of them could have wrapped
  // (conservative bound from the analysis ranges; ~1.4e9 accepted events per bin for the shipped analyses)
  const double fx_x = band_scale_x(ana);
  {
    const double mx = ana->useS2 == 2 ? fabs(ana->maxS2) : fabs(ana->maxS1);
    const double ly = std::max(fabs(ana->logMin), fabs(ana->logMax));
    const double per_event = std::max(std::max(mx * fx_x, ly * NB200_FX_Y), std::max(ly * ly * NB200_FX_Y2, ly * ly * ly * NB200_FX_Y3));
    for (int j = 0; j < ana->numBins; ++j)
      if (double(h->accepted[j]) * per_event >= 9.2e18)
        return fail(nullptr, NB200_EINVAL,
                    "band accumulator: a bin holds more events than its 64-bit fixed-point sums can carry; "
                    "accumulate fewer events per nb200_band_hist (or more bins) and merge the statistics");
  }
  for (int j = 0; j < ana->numBins; ++j) {
    double* b = band + 7 * size_t(j);
    const double N = double(h->accepted[j]);
    const double Sy = double(h->sumY[j]) / NB200_FX_Y, Sy2 = double(h->sumY2[j]) / NB200_FX_Y2,
                 Sy3 = double(h->sumY3[j]) / NB200_FX_Y3, Sx = double(h->sumS1[j]) / fx_x;
    b[0] = border + binWidth / 2. + double(j) * binWidth;
    b[1] = Sx / N;
    const double mu = Sy / N;
    b[2] = mu;
    double var = (Sy2 - 2. * mu * Sy + N * mu * mu) / (N - 1.);
    b[3] = sqrt(var);
    b[4] = b[3] / sqrt(N);
    b[5] = N / (N + double(h->rejected[j]));
    double m3 = Sy3 - 3. * mu * Sy2 + 3. * mu * mu * Sy - N * mu * mu * mu;
    b[6] = m3 / (b[3] * b[3] * b[3]) / (N - 2.);
  }
  return 0;
}

// ---- band post-processing on the host (no device needed) ----------------------------------
// examples/rootNEST.cpp:600-643: goodness of fit between two bands
int nb200_band_chi2(int numBins, int useS2, const double* a, const double* b, int free_param, double* out) {
  if (!a || !b || !out || numBins < 1) return NB200_EINVAL;
  double chi2[4] = {0., 0., 0., 0.};
  const int DoF = numBins - std::abs(free_param);
  for (int i = 0; i < numBins; ++i) {
    const double* x = a + 7 * size_t(i);  // NEST ("band")
    const double* y = b + 7 * size_t(i);  // data ("band2")
    if (std::fabs(x[0] - y[0]) > 0.05) return NB200_EINVAL;  // "Binning doesn't match for GoF calculation"
    if (!(std::isfinite(x[2]) && std::isfinite(y[2]) && std::isfinite(x[3]) && std::isfinite(y[3]) && x[4] > 0. && y[4] > 0.))
      continue;  // empty bin on either side
    if ((useS2 == 0 && std::fabs(x[2]) < 5.) || useS2 != 0) {
      double error = std::sqrt(x[4] * x[4] + y[4] * y[4]);
      chi2[0] += (y[2] - x[2]) / error * ((y[2] - x[2]) / error);
      error = std::sqrt(0.5 * (x[4] * x[4] + y[4] * y[4]));  // sigma/sqrt(2N) on each side
      chi2[1] += (y[3] - x[3]) / error * ((y[3] - x[3]) / error);
      chi2[2] += 100. * (y[2] - x[2]) / y[2];
      chi2[3] += 100. * (y[3] - x[3]) / y[3];
    }
  }
  chi2[0] /= double(DoF - 1);
  chi2[1] /= double(DoF - 1);
  chi2[2] /= double(numBins);
  chi2[3] /= double(numBins);
  out[0] = chi2[0];
  out[1] = chi2[1];
  out[2] = 0.5 * (chi2[0] + chi2[1]);
  out[3] = std::sqrt(chi2[0] * chi2[1]);
  out[4] = -chi2[2];
  out[5] = -chi2[3];
  return 0;
}

// examples/rootNEST.cpp:880-915 (counting, per-bin centroid): leakage of the accumulator's events below a centroid
int nb200_band_leakage(const nb200_analysis* ana, const nb200_band_hist* h, const double* centroid, double* leak,
                       double* err_hi, double* err_lo, double* total) {
  if (!ana || !h || !centroid || !leak || !h->accepted || !h->hist2d || h->logBins < 1) return NB200_EINVAL;
  const int nb = h->numBins, lb = h->logBins;
  const double w = (ana->logMax - ana->logMin) / double(lb);
  double sum_below = 0., sum_all = 0.;
  for (int i = 0; i < nb; ++i) {
    const uint64_t* row = h->hist2d + size_t(i) * lb;
    double in_hist = 0., below = 0.;
    const double pos = (centroid[i] - ana->logMin) / w;  // in units of log bins
    for (int j = 0; j < lb; ++j) {
      const double c = double(row[j]);
      in_hist += c;
      if (double(j + 1) <= pos)
        below += c;
      else if (double(j) < pos)
        below += c * (pos - double(j));  // the bin holding the centroid, split linearly
    }
    const double N = double(h->accepted[i]);
    // events whose y fell outside (logMin, logMax) entered the reference's vectors as 0 (execNEST.cpp:1551-1555)
    if (centroid[i] > 0.) below += N - in_hist;
    leak[i] = below / N;
    if (err_hi) err_hi[i] = (below + std::sqrt(below)) / (N - std::sqrt(N)) - leak[i];
    if (err_lo) err_lo[i] = leak[i] - (below - std::sqrt(below)) / (N + std::sqrt(N));
    if (N > 0.) {
      sum_below += below;
      sum_all += N;
    }
  }
  if (total) {
    total[0] = sum_below;
    total[1] = sum_all;
    total[2] = sum_below / sum_all;
  }
  return 0;
}

}  // extern "C"
namespace {
// The fits themselves live in nb_fit.h (one implementation for these host entry points and for k_band_fit on the
// device): weighted least squares by Levenberg-Marquardt with analytic Jacobians, errors from (J^T W J)^-1.
inline int model_npar(int model) { return fit_npar(model); }
bool fit_points(int model, int npts, const double* xs, const double* ys, const double* ws, double* p, double* err, double* chi2,
                double* cov = nullptr) {
  return fit_lm(model, FitPoints{xs, ys, ws, npts}, p, err, chi2, cov) != 0;
}
// a histogram row as weighted points: bin centres, contents, weights 1/content, empty bins left out (TH1::Fit's default
// chi^2); returns the number of bins used, 0 when there are too few or no minimum was found
int fit_hist_row(int model, const uint64_t* row, int lb, double x0, double w, double* p, double* err, double* chi2,
                 double* cov = nullptr) {
  return fit_lm(model, FitRow<uint64_t>{row, lb, x0, w}, p, err, chi2, cov);
}
}  // namespace
extern "C" {

// examples/rootNEST.cpp:1309-1359 (GetBand_Gaussian, skewness == 0): Gaussian fit of every S1 bin's histogram
int nb200_band_fit_gauss(const nb200_analysis* ana, const nb200_band_hist* h, double* fit) {
  if (!ana || !h || !fit || !h->accepted || !h->sumY || !h->sumY2 || !h->hist2d || h->logBins < 4) return NB200_EINVAL;
  const int nb = h->numBins, lb = h->logBins;
  const double w = (ana->logMax - ana->logMin) / double(lb);
  for (int j = 0; j < nb; ++j)
    fit_gauss_row<FitSerial>(h->hist2d + size_t(j) * lb, lb, ana->logMin, w, double(h->accepted[j]), fit + 8 * size_t(j));
  return 0;
}

// examples/rootNEST.cpp:700-738, 848 (skewness == 0): leakage from the Gaussian picture of the two bands
int nb200_band_leakage_gauss(int numBins, const double* er, const double* nr, double num_sig_above, double* out) {
  if (!er || !nr || !out || numBins < 1) return NB200_EINVAL;
  for (int i = 0; i < numBins; ++i) {
    const double* e = er + 8 * size_t(i);
    const double* r = nr + 8 * size_t(i);
    double* o = out + 4 * size_t(i);
    const double line = r[1] + num_sig_above * r[2];  // NUMSIGABV (:39, :718)
    const double numSigma = (e[1] - line) / e[2];
    const double up = (e[1] - e[3] - line - r[3]) / (e[2] + e[4]);  // more leakage
    const double dn = (e[1] + e[3] - line + r[3]) / (e[2] - e[4]);  // less leakage
    const double leak = (1. - std::erf(numSigma / std::sqrt(2.))) / 2.;
    o[0] = numSigma;
    o[1] = leak;
    o[2] = std::fabs((1. - std::erf(up / std::sqrt(2.))) / 2. - leak);
    o[3] = std::fabs(leak - (1. - std::erf(dn / std::sqrt(2.))) / 2.);
  }
  return 0;
}

// One histogram, one fit: the building block of GetBand_Gaussian for callers that keep their own histograms
// (rootNEST's skew-Gaussian mode re-bins every S1 slice on mean +- 3 sigma, rootNEST.cpp:1319-1330)
int nb200_hist_fit(int model, const uint64_t* counts, int bins, double lo, double hi, const double* start, double* par,
                   double* err, double* chi2_ndf) {
  if (!counts || !start || !par || !err || bins < 1 || !(hi > lo) || (model != 0 && model != 1)) return NB200_EINVAL;
  const int n = model_npar(model);
  double p[kFitMaxPar], e[kFitMaxPar] = {0., 0., 0., 0.}, chi = 0.;
  for (int i = 0; i < n; ++i) p[i] = start[i];
  const int used = fit_hist_row(model, counts, bins, lo, (hi - lo) / double(bins), p, e, &chi);
  if (!used) return NB200_EFIT;
  for (int i = 0; i < n; ++i) {
    par[i] = p[i];
    err[i] = e[i];
  }
  if (chi2_ndf) *chi2_ndf = chi / double(used - n);
  return 0;
}

// the same fit with the full covariance matrix (J^T W J)^-1 of the parameters, row-major [npar][npar]: what rootNEST's
// skewness = 2 reads from TFitResult::GetCovarianceMatrix (rootNEST.cpp:1399-1405)
int nb200_hist_fit_cov(int model, const uint64_t* counts, int bins, double lo, double hi, const double* start, double* par,
                       double* cov, double* chi2_ndf) {
  if (!counts || !start || !par || !cov || bins < 1 || !(hi > lo) || (model != 0 && model != 1)) return NB200_EINVAL;
  const int n = model_npar(model);
  double p[kFitMaxPar], e[kFitMaxPar] = {0., 0., 0., 0.}, chi = 0.;
  for (int i = 0; i < n; ++i) p[i] = start[i];
  const int used = fit_hist_row(model, counts, bins, lo, (hi - lo) / double(bins), p, e, &chi, cov);
  if (!used) return NB200_EFIT;
  for (int i = 0; i < n; ++i) par[i] = p[i];
  if (chi2_ndf) *chi2_ndf = chi / double(used - n);
  return 0;
}

// examples/rootNEST.cpp:774-845 (skewness == 2): the skew-normal leakage below line[i] with its error propagated from the
// fit's parameter errors and covariances and the line's own error.  skew [numBins][9] = {xi, xi error, omega, omega
// error, alpha, alpha error, cov(xi, omega), cov(xi, alpha), cov(omega, alpha)}
int nb200_band_leakage_skew_cov(int numBins, const double* skew, const double* line, const double* line_err, double* leak,
                                double* err) {
  if (!skew || !line || !line_err || !leak || !err || numBins < 1) return NB200_EINVAL;
  for (int i = 0; i < numBins; ++i) {
    const double* k = skew + 9 * size_t(i);
    const double xi = k[0], om = k[2], al = k[4], c = line[i];
    leak[i] = 0.5 + 0.5 * std::erf((c - xi) / om / std::sqrt(2.)) - 2. * owens_t_ref((c - xi) / om, al);
    const double pdf = (1. / (om * std::sqrt(2. * M_PI))) * std::exp(-0.5 * (c - xi) * (c - xi) / (om * om)) *
                       (1. + std::erf(al * (c - xi) / (om * std::sqrt(2.))));
    const double d_x = pdf, d_xi = -1. * pdf, d_om = -1. * (c - xi) / om * pdf;
    const double d_al = -1. / M_PI / (1. + al * al) * std::exp(-1. * (1. + al * al) * (c - xi) * (c - xi) / 2. / (om * om));
    err[i] = std::sqrt(0. + d_x * d_x * line_err[i] * line_err[i] + d_xi * d_xi * k[1] * k[1] + d_om * d_om * k[3] * k[3] +
                       d_al * d_al * k[5] * k[5] + 2 * d_xi * d_om * k[6] + 2 * d_om * d_al * k[8] + 2 * d_xi * d_al * k[7]);
  }
  return 0;
}

// TGraph::Fit without ROOT for the two curves rootNEST puts through the NR band's means (rootNEST.cpp:855-878): unweighted
// least squares, so "chi^2 / NDF" is the mean squared residual the reference tests against 1.5 and 2
int nb200_graph_fit(int model, int n, const double* x, const double* y, const double* start, double* par, double* chi2_ndf) {
  if (!x || !y || !start || !par || n < 1 || (model != 2 && model != 3)) return NB200_EINVAL;
  std::vector<double> ws(size_t(n), 1.);
  double p[kFitMaxPar], e[kFitMaxPar], chi = 0.;
  for (int i = 0; i < 4; ++i) p[i] = start[i];
  if (!fit_points(model, n, x, y, ws.data(), p, e, &chi)) return NB200_EFIT;
  for (int i = 0; i < 4; ++i) par[i] = p[i];
  if (chi2_ndf) *chi2_ndf = chi / double(n - 4);
  return 0;
}

// examples/rootNEST.cpp:700-773 (skewness == 1): leakage below a line from a skew-Gaussian ER band
int nb200_band_leakage_skew(int numBins, const double* xi, const double* omega, const double* omega_err, const double* alpha,
                            const double* line, double* leak, double* err_hi, double* err_lo) {
  if (!xi || !omega || !alpha || !line || !leak || numBins < 1) return NB200_EINVAL;
  auto cdf = [](double c, double xi_, double om, double al) {
    return 0.5 + 0.5 * std::erf((c - xi_) / om / std::sqrt(2.)) - 2. * owens_t_ref((c - xi_) / om, al);
  };
  for (int i = 0; i < numBins; ++i) {
    leak[i] = cdf(line[i], xi[i], omega[i], alpha[i]);
    if (omega_err && err_hi) err_hi[i] = cdf(line[i], xi[i], omega[i] + omega_err[i], alpha[i]);  // :741-771, "focused on omega"
    if (omega_err && err_lo) err_lo[i] = cdf(line[i], xi[i], omega[i] - omega_err[i], alpha[i]);
  }
  return 0;
}

// ---- FP64 pipe peak ---------------------------------------------------------------------
}  // extern "C"
namespace {
__global__ void __launch_bounds__(256) k_dfma_peak(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x * 1e-3, x1 = x0 + 1., x2 = x0 + 2., x3 = x0 + 3., x4 = x0 + 4., x5 = x0 + 5.,
         x6 = x0 + 6., x7 = x0 + 7.;
  for (int i = 0; i < iters; ++i) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
  if (s == 123.456) out[0] = s;
}
}  // namespace
extern "C" {
int nb200_se_stats(nb200_ctx* c, const nb200_detector* det, const nb200_run_consts* rc, int n, uint64_t seed,
                   double* se_mean, double* se_width) {
  if (!c || !det || !rc || !se_mean || !se_width || n < 2) return c ? fail(c, NB200_EINVAL, "bad argument") : NB200_EINVAL;
  NB_CUDA(c, cudaSetDevice(c->device));
  RunParams P;
  std::memset(&P, 0, sizeof P);
  P.det = *det;
  P.rc = *rc;
  int r;
  if ((r = upload_map(c, &P.det.FitS2))) return r;
  if ((r = upload_map(c, &P.det.FitTBA))) return r;
  P.det.FitS1.lut = nullptr;
  P.det.FitEF.lut = nullptr;
  P.s2_center = fit_s2(det->FitS2, 0., 0.);
  P.seed = seed;
  philox_round_keys(seed, P.rk);
  double* d = nullptr;
  NB_CUDA(c, cudaMalloc(&d, sizeof(double) * n));
  k_se<<<(n + 127) / 128, 128, 0, c->stream>>>(P, n, d);
  ++c->launches;
  std::vector<double> h(n);
  cudaMemcpyAsync(h.data(), d, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream);
  cudaError_t e = cudaStreamSynchronize(c->stream);
  cudaFree(d);
  if (e != cudaSuccess) return fail(c, NB200_ECUDA, cudaGetErrorString(e));
  const double SE = rc->g2_params[0] * det->g1_gas *
                    (det->s2_thr < 0 ? fit_tba(det->FitTBA, 0., 0., det->TopDrift / 2.) : 1.);
  double mu = 0., sd = 0.;
  for (double v : h) {
    sd += (SE - v) * (SE - v);
    mu += v;
  }
  *se_mean = mu / double(n);
  *se_width = sqrt(sd) / sqrt(double(n) - 1.);
  return 0;
}

int nb200_timing_enable(nb200_ctx* c, int on) {
  if (!c) return NB200_EINVAL;
  c->timing = on != 0;
  return 0;
}

int nb200_timing_read(nb200_ctx* c, double* ms, uint64_t* launches) {
  if (!c || !ms || !launches) return NB200_EINVAL;
  NB_CUDA(c, cudaStreamSynchronize(c->stream));
  for (int k = 0; k < 3; ++k) {
    ms[k] = 0.;
    launches[k] = 0;
  }
  for (size_t i = 0; i + 3 < c->tev.size(); i += 4) {
    for (int k = 0; k < 3; ++k) {
      float t = 0.f;
      cudaEventElapsedTime(&t, c->tev[i + k], c->tev[i + k + 1]);
      ms[k] += t;
      ++launches[k];
    }
  }
  for (auto e : c->tev) cudaEventDestroy(e);
  c->tev.clear();
  return 0;
}

int nb200_fp64_peak(nb200_ctx* c, double* tflops) {
  if (!c || !tflops) return NB200_EINVAL;
  NB_CUDA(c, cudaSetDevice(c->device));
  double* d = nullptr;
  NB_CUDA(c, cudaMalloc(&d, 8));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 1 << 15, blocks = c->sm_count * 8;
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0, c->stream);
    k_dfma_peak<<<blocks, 256, 0, c->stream>>>(d, iters, 0.999999, 1e-9);
    cudaEventRecord(e1, c->stream);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
    ++c->launches;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  NB_CUDA(c, cudaGetLastError());
  *tflops = 2.0 * 8.0 * double(iters) * 256.0 * double(blocks) / (double(best) * 1e-3) / 1e12;
  return 0;
}

// test hook (tests/test_gpu_hitmath.py): {-ln u, sin, cos} of the hit loop's bit-built math
int nb200_debug_hitmath(nb200_ctx* c, size_t n, const uint32_t* bits3, double* out3) {
  if (!c || !bits3 || !out3) return NB200_EINVAL;
  if (n == 0) return 0;
  NB_CUDA(c, cudaSetDevice(c->device));
  uint32_t* db = nullptr;
  double* dout = nullptr;
  NB_CUDA(c, cudaMalloc(&db, n * 12));
  NB_CUDA(c, cudaMalloc(&dout, n * 24));
  cudaMemcpyAsync(db, bits3, n * 12, cudaMemcpyHostToDevice, c->stream);
  k_hitmath<<<unsigned(std::min<size_t>((n + 255) / 256, 4096)), 256, 0, c->stream>>>(n, db, dout);
  ++c->launches;
  cudaMemcpyAsync(out3, dout, n * 24, cudaMemcpyDeviceToHost, c->stream);
  cudaError_t e = cudaStreamSynchronize(c->stream);
  cudaFree(db);
  cudaFree(dout);
  if (e != cudaSuccess) return fail(c, NB200_ECUDA, cudaGetErrorString(e));
  return 0;
}

int nb200_debug_normal(nb200_ctx* c, uint64_t seed, uint64_t first_event, size_t slots, double* out,
                       unsigned char* kind) {
  if (!c || !out) return NB200_EINVAL;
  if (slots == 0) return 0;
  NB_CUDA(c, cudaSetDevice(c->device));
  double* dout = nullptr;
  unsigned char* dk = nullptr;
  NB_CUDA(c, cudaMalloc(&dout, slots * 24));
  if (kind && cudaMalloc(&dk, slots * 3) != cudaSuccess) {
    cudaFree(dout);
    return fail(c, NB200_ENOMEM, "cudaMalloc");
  }
  k_zignormal<<<unsigned(std::min<size_t>((slots + 255) / 256, 148 * 8)), 256, 0, c->stream>>>(seed, first_event, slots,
                                                                                             dout, dk);
  ++c->launches;
  cudaMemcpyAsync(out, dout, slots * 24, cudaMemcpyDeviceToHost, c->stream);
  if (kind) cudaMemcpyAsync(kind, dk, slots * 3, cudaMemcpyDeviceToHost, c->stream);
  cudaError_t e = cudaStreamSynchronize(c->stream);
  cudaFree(dout);
  cudaFree(dk);
  if (e != cudaSuccess) return fail(c, NB200_ECUDA, cudaGetErrorString(e));
  return 0;
}

int nb200_debug_binomial(nb200_ctx* c, uint64_t seed, uint64_t first_event, size_t count, int64_t n, double p,
                         int64_t* out) {
  if (!c || !out) return NB200_EINVAL;
  if (count == 0) return 0;
  NB_CUDA(c, cudaSetDevice(c->device));
  long long* dout = nullptr;
  NB_CUDA(c, cudaMalloc(&dout, count * 8));
  RoundKeys K;
  philox_round_keys(seed, K.k);
  k_binomial<<<unsigned(std::min<size_t>((count + 255) / 256, 148 * 8)), 256, 0, c->stream>>>(
      K, first_event, count, (long long)n, p, dout);
  ++c->launches;
  cudaMemcpyAsync(out, dout, count * 8, cudaMemcpyDeviceToHost, c->stream);
  cudaError_t e = cudaStreamSynchronize(c->stream);
  cudaFree(dout);
  if (e != cudaSuccess) return fail(c, NB200_ECUDA, cudaGetErrorString(e));
  return 0;
}

int nb200_debug_binomial_list(nb200_ctx* c, uint64_t seed, uint64_t first_event, size_t count, const int64_t* n, double p,
                              int which_stream, int64_t* out) {
  if (!c || !out || !n || which_stream < 0 || which_stream > 1) return NB200_EINVAL;
  if (count == 0) return 0;
  NB_CUDA(c, cudaSetDevice(c->device));
  long long *dn = nullptr, *dout = nullptr;
  NB_CUDA(c, cudaMalloc(&dn, count * 8));
  if (cudaMalloc(&dout, count * 8) != cudaSuccess) {
    cudaFree(dn);
    return fail(c, NB200_ENOMEM, "cudaMalloc");
  }
  cudaMemcpyAsync(dn, n, count * 8, cudaMemcpyHostToDevice, c->stream);
  RoundKeys K;
  philox_round_keys(seed, K.k);
  k_binomial_list<<<unsigned(std::min<size_t>((count + 255) / 256, 148 * 8)), 256, 0, c->stream>>>(
      K, first_event, count, dn, p, which_stream == 1 ? uint32_t(STREAM_BIN_LAR) : uint32_t(STREAM_BIN_S1), dout);
  ++c->launches;
  cudaMemcpyAsync(out, dout, count * 8, cudaMemcpyDeviceToHost, c->stream);
  cudaError_t e = cudaStreamSynchronize(c->stream);
  cudaFree(dn);
  cudaFree(dout);
  if (e != cudaSuccess) return fail(c, NB200_ECUDA, cudaGetErrorString(e));
  return 0;
}

// host-only test hooks (no device needed): the tables the samplers above are built on
double nb200_drift_velocity_liquid(double Kelvin, double field) {
  bool ok = true;
  const double v = host::drift_velocity_liquid(Kelvin, field, &ok);
  return ok ? v : std::nan("");
}

double nb200_density(double Kelvin, double bara, double molarMass, int* in_gas) {
  bool g = in_gas && *in_gas != 0;
  const double rho = host::density(Kelvin, bara, g, molarMass, nullptr);
  if (in_gas) *in_gas = g ? 1 : 0;
  return rho;
}

int nb200_drift_table(const nb200_detector* det, double z_step, double x, double y, double* table, size_t* n) {
  if (!det || !n || !(z_step > 0.) || !(det->TopDrift > 0.)) return NB200_EINVAL;
  if (!table) {
    size_t k = 0;
    for (double z = 0.0; z < det->TopDrift; z += z_step) ++k;  // the reference's own loop (:2669)
    *n = k;
    return 0;
  }
  const std::vector<double> t = host::drift_table_nonuniform(*det, z_step, x, y);
  if (*n < t.size()) return NB200_EINVAL;
  for (size_t i = 0; i < t.size(); ++i) table[i] = t[i];
  *n = t.size();
  return 0;
}

int nb200_debug_maps(const nb200_detector* det, size_t n, const double* x, const double* y, const double* z, double* s1,
                     double* s2, double* ef, double* tba) {
  if (!det || !x || !y || !z) return NB200_EINVAL;
  for (size_t i = 0; i < n; ++i) {
    if (s1) s1[i] = fit_s1(det->FitS1, x[i], y[i], z[i]);
    if (s2) s2[i] = fit_s2(det->FitS2, x[i], y[i]);
    if (ef) ef[i] = fit_ef(det->FitEF, x[i], y[i], z[i]);
    if (tba) tba[i] = fit_tba(det->FitTBA, x[i], y[i], z[i]);
  }
  return 0;
}

int nb200_debug_zig_tables(int n_layers, double* x, double* f, uint32_t* k, double* w) {
  if (n_layers != NB_ZIG_N || !x || !f || !k || !w) return NB200_EINVAL;
  build_zig_tables(x, f, k, w);
  return 0;
}
int nb200_debug_alias_tables(double p, int n_entries, uint32_t* thr, uint32_t* alias) {
  if (n_entries != NB_ALIAS_ENTRIES || !(p > 0. && p < 1.) || !thr || !alias) return NB200_EINVAL;
  std::vector<AliasEntry> tab(NB_ALIAS_ENTRIES);
  build_binomial_alias(p, tab.data());
  for (int i = 0; i < NB_ALIAS_ENTRIES; ++i) {
    thr[i] = tab[i].thr;
    alias[i] = tab[i].alias;
  }
  return 0;
}

int nb200_debug_binomial_fixed(nb200_ctx* c, uint64_t seed, uint64_t first_event, size_t count, int n, double p,
                               int64_t* out) {
  if (!c || !out || !(p > 0. && p < 1.) || n < 0 || n >= NB_ALIAS_MAXN) return NB200_EINVAL;
  if (count == 0) return 0;
  NB_CUDA(c, cudaSetDevice(c->device));
  std::vector<AliasEntry> tab(NB_ALIAS_ENTRIES);
  build_binomial_alias(p, tab.data());
  long long* dout = nullptr;
  void* dtab = nullptr;
  NB_CUDA(c, cudaMalloc(&dout, count * 8));
  if (cudaMalloc(&dtab, sizeof(AliasEntry) * tab.size()) != cudaSuccess) {
    cudaFree(dout);
    return fail(c, NB200_ENOMEM, "cudaMalloc");
  }
  cudaMemcpyAsync(dtab, tab.data(), sizeof(AliasEntry) * tab.size(), cudaMemcpyHostToDevice, c->stream);
  RoundKeys K;
  philox_round_keys(seed, K.k);
  k_binomial_fixed<<<unsigned(std::min<size_t>((count + 255) / 256, 148 * 8)), 256, 0, c->stream>>>(
      K, first_event, count, n, static_cast<const uint2*>(dtab), dout);
  ++c->launches;
  cudaMemcpyAsync(out, dout, count * 8, cudaMemcpyDeviceToHost, c->stream);
  cudaError_t e = cudaStreamSynchronize(c->stream);
  cudaFree(dout);
  cudaFree(dtab);
  if (e != cudaSuccess) return fail(c, NB200_ECUDA, cudaGetErrorString(e));
  return 0;
}

// ---- liquid argon ---------------------------------------------------------------------------

int nb200_lar(nb200_ctx* c, int species, size_t n, const double* E, const double* field, double density,
              uint64_t seed, uint64_t first_event, int mem_kind, double* yields, double* fluct) {
  if (!c) return NB200_EINVAL;
  if (!E || !field) return fail(c, NB200_EINVAL, "null argument");
  if (species < 0 || species > 2)
    return fail(c, NB200_EINVAL, "LAr species must be NR(0), ER(1) or Alpha(2); dE/dx and LET models are not part of this path");
  if (n == 0) return 0;
  NB_CUDA(c, cudaSetDevice(c->device));
  LArParams P;
  std::memset(&P, 0, sizeof P);
  P.species = species;
  P.n = n;
  philox_round_keys(seed, P.rk);
  P.first_event = first_event;
  P.seed = seed;
  P.density = density;
  std::vector<void*> fr;
  auto cleanup = [&]() {
    for (void* p : fr) cudaFree(p);
  };
  double *dE = nullptr, *dF = nullptr, *dY = nullptr, *dFl = nullptr;
  if (mem_kind == NB200_MEM_HOST) {
    cudaError_t e = cudaMalloc(&dE, n * 8);
    if (e == cudaSuccess) { fr.push_back(dE); e = cudaMalloc(&dF, n * 8); }
    if (e == cudaSuccess) { fr.push_back(dF); if (yields) { e = cudaMalloc(&dY, n * 72); if (e == cudaSuccess) fr.push_back(dY); } }
    if (e == cudaSuccess && fluct) { e = cudaMalloc(&dFl, n * 32); if (e == cudaSuccess) fr.push_back(dFl); }
    if (e != cudaSuccess) {
      cleanup();
      return fail(c, NB200_ENOMEM, cudaGetErrorString(e));
    }
    cudaMemcpyAsync(dE, E, n * 8, cudaMemcpyHostToDevice, c->stream);
    cudaMemcpyAsync(dF, field, n * 8, cudaMemcpyHostToDevice, c->stream);
    P.E = dE;
    P.field = dF;
    P.yields = dY;
    P.fluct = dFl;
  } else {
    P.E = E;
    P.field = field;
    P.yields = yields;
    P.fluct = fluct;
  }
  unsigned grid = unsigned(std::min<uint64_t>((n + 255) / 256, uint64_t(c->sm_count) * 8));
  if (species == 0)
    k_lar<0><<<grid, 256, 0, c->stream>>>(P);
  else if (species == 1)
    k_lar<1><<<grid, 256, 0, c->stream>>>(P);
  else
    k_lar<2><<<grid, 256, 0, c->stream>>>(P);
  cudaError_t le = cudaGetLastError();
  ++c->launches;
  int rc = 0;
  if (le != cudaSuccess) rc = fail(c, NB200_ECUDA, std::string("k_lar: ") + cudaGetErrorString(le));
  if (!rc && mem_kind == NB200_MEM_HOST) {
    if (yields) cudaMemcpyAsync(yields, dY, n * 72, cudaMemcpyDeviceToHost, c->stream);
    if (fluct) cudaMemcpyAsync(fluct, dFl, n * 32, cudaMemcpyDeviceToHost, c->stream);
    cudaError_t e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) rc = fail(c, NB200_ECUDA, cudaGetErrorString(e));
  }
  cleanup();
  return rc;
}

}  // extern "C"
