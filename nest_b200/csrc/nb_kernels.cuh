// nb_kernels.cuh -- the sm_100a kernels: k_pre / k_hits / k_post (the chain, with the S1 hit
// loops flattened across events), k_yields (deterministic means) and k_lar (liquid-argon
// yields + fluctuations).
#pragma once
#include "nb_chain.cuh"
#include "nb_fit.h"
#include "nb_lar.cuh"

namespace nb {

// -------------------------------------------------------------------------------------------
// k_yields: NESTcalc::GetYields + YieldResultValidity (NEST.cpp:1211-1271) for n (E, field) pairs.
//   which: -1 species dispatch, 0 GetYieldBetaGR(multFact), 1 GetYieldGamma(multFact),
//          2 GetYieldERWeighted.  out: 6 columns of n (PhotonYield, ElectronYield, ExcitonRatio,
//          Lindhard, ElectricField, DeltaT_Scint).
// -------------------------------------------------------------------------------------------
struct YieldsParams {
  nb200_detector det;
  nb200_model model;
  int32_t species, which;
  double density;
  uint64_t n;
  uint32_t rk[20];  // Philox round keys of the seed
  const double *E, *field;
  double* out;
};

__global__ void __launch_bounds__(256) k_yields(const __grid_constant__ YieldsParams P, int* err_flag) {
  for (uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; i < P.n;
       i += uint64_t(gridDim.x) * blockDim.x) {
    const double E = P.E[i], F = P.field[i];
    int err = 0;
    Yields y;
    if (P.which == 0)
      y = yield_beta(E, P.density, F, P.det, P.model.ERYieldsParam, P.model.multFact);
    else if (P.which == 1)
      y = yield_gamma(E, P.density, F, P.det, P.model.multFact);
    else if (P.which == 2)
      y = yield_er_weighted(E, P.density, F, P.det, P.model.ERYieldsParam);
    else {
      int A2 = 131;
      double u01 = 0.5;
      const bool nrlike = P.species <= NB200_Cf || P.species == NB200_atmNu;
      if ((nrlike && nearly_equal(P.model.massNum, 0.)) || P.species == NB200_ion) {
        Stream g;
        g.init(P.rk, i, STREAM_MAIN);
        A2 = select_xe_atom(g);
      } else if (P.species == NB200_Kr83m) {
        Stream g;
        g.init(P.rk, i, STREAM_MAIN);
        u01 = g.uniform();
      }
      y = get_yields(P.species, E, P.density, F, P.model.massNum, P.model.atomNum, A2, P.det, P.model,
                     &err, u01);
    }
    if (err) atomicOr(err_flag, err);
    P.out[i] = y.Nph;
    P.out[P.n + i] = y.Ne;
    P.out[2 * P.n + i] = y.NexONi;
    P.out[3 * P.n + i] = y.L;
    P.out[4 * P.n + i] = y.field;
    P.out[5 * P.n + i] = y.dT;
  }
}

// -------------------------------------------------------------------------------------------
// The chain: the loop body of execNEST() (execNEST.cpp:676-1414) / runNESTvec() (:357-406) as
// three kernels over a chunk of events whose per-event state lives in a device workspace
// (structure of arrays, ~140 B/event, written by one stage and read by the next):
//
//   k_pre   one thread = one event: energy, vertex (drift-window retry), yields, quanta, xy
//           smearing and the S1 prelude up to nHits = binomial(Nph, g1(x,y,z)).
//   k_hits  the per-hit S1 loop (NEST.cpp:1367-1429), the dominant cost.  Hit counts differ by
//           orders of magnitude between events, so a block takes a GROUP of events, prefix-sums
//           their work items (slots of two photo-electrons, see the kernel) in shared memory and
//           deals the flattened range out evenly: every thread runs the same number of slots
//           (+-1) whatever events they belong to.  Each slot draws from its own Philox block keyed
//           (seed; event, slot index) and the areas are accumulated as 2^-32 phe integers, so the
//           result for an event does not depend on how the range was cut (bit-identical for any
//           batch split / GPU count).
//   k_post  one thread = one event: S1 epilogue (noise, thresholds, coincidence, spike smearing),
//           S2, signal selection, energy reconstruction, band accumulation (shared-memory
//           integer accumulators flushed once per block) and the optional SoA record.
// -------------------------------------------------------------------------------------------
// block-wide exclusive prefix sum of a pair of counts (k_hits: slots, and events that have slots)
template <int T>
__device__ __forceinline__ int2 block_exclusive_scan2(int2 v, int2* s_warp) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int2 x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int tx = __shfl_up_sync(0xffffffffu, x.x, o), ty = __shfl_up_sync(0xffffffffu, x.y, o);
    if (lane >= o) {
      x.x += tx;
      x.y += ty;
    }
  }
  if (lane == 31) s_warp[wid] = x;
  __syncthreads();
  if (wid == 0) {
    int2 w = (lane < T / 32) ? s_warp[lane] : make_int2(0, 0);
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int tx = __shfl_up_sync(0xffffffffu, w.x, o), ty = __shfl_up_sync(0xffffffffu, w.y, o);
      if (lane >= o) {
        w.x += tx;
        w.y += ty;
      }
    }
    if (lane < T / 32) s_warp[lane] = w;
  }
  __syncthreads();
  const int2 base = wid ? s_warp[wid - 1] : make_int2(0, 0);
  return make_int2(base.x + x.x - v.x, base.y + x.y - v.y);
}

// GetDriftVelocity_Liquid (NEST.cpp:2482-2502) for the detector temperature; the two
// bracketing polyExp rows and the temperature blend are chosen on the host.
NB_D double drift_velocity_dev(const RunParams& P, double eField) {
  const double* a = P.dv_row[0];
  const double* b = P.dv_row[1];
  double vi = a[0] * exp(-eField / a[1]) + a[2] * exp(-eField / a[3]) + a[4] * exp(-eField / a[5]) + a[6];
  if (P.dv_mode == 1) return vi;
  double vf = b[0] * exp(-eField / b[1]) + b[2] * exp(-eField / b[3]) + b[4] * exp(-eField / b[5]) + b[6];
  if (P.dv_mode == 2) return vf;
  const double Ti = P.dv_Ti, Tf = P.dv_Tf, Kelvin = P.dv_T;
  double speed;
  if (vf < vi) {
    double offset = (sqrt((Tf * (vf - vi) - Ti * (vf - vi) - 4.) * (vf - vi)) + sqrt(Tf - Ti) * (vf + vi)) /
                    (2. * sqrt(Tf - Ti));
    double slope = -(sqrt(Tf - Ti) * sqrt((Tf * (vf - vi) - Ti * (vf - vi) - 4.) * (vf - vi)) -
                     (Tf + Ti) * (vf - vi)) /
                   (2. * (vf - vi));
    speed = 1. / (Kelvin - slope) + offset;
  } else {
    double slope = (vf - vi) / (Tf - Ti);
    speed = slope * (Kelvin - Ti) + vi;
  }
  if (speed <= 0.) speed = 0.1;
  return speed;
}

#ifndef NB_NO_LOCKSTEP
#define NB_LOCKSTEP() __syncthreads()
#else
#define NB_LOCKSTEP() ((void)0)
#endif
// launch geometry, measured on B200 (tools/sweep_variants.sh): k_pre and k_post are fastest at 64
// registers with 32 warps per SM in large lockstep blocks (k_post one 1024-thread block, k_pre two
// 512-thread blocks: the spills cost less than the FP64 dependency stalls the extra warps hide, and
// 128-thread blocks are 35 % slower for lack of instruction-cache sharing); k_hits in 128-thread
// blocks at 128 registers (4 per SM)
#ifndef NB_PRE_BLOCK
#define NB_PRE_BLOCK 512
#endif
#ifndef NB_PRE_MINB
#define NB_PRE_MINB 2
#endif
#ifndef NB_POST_BLOCK
#define NB_POST_BLOCK 1024
#endif
#ifndef NB_POST_MINB
#define NB_POST_MINB 1
#endif
#ifndef NB_HIT_MINB
#define NB_HIT_MINB 4  // 128-thread blocks at 128 registers (no spills, 16 warps/SM) with the 24 KB ziggurat tables: measured best of
                       // {256x4@64, 256x3@80, 256x2@128, 128x6@80, 128x5@96, 128x4@128} x {1024, 2048 layers}
#endif
constexpr int kPreBlock = NB_PRE_BLOCK;
constexpr int kPostBlock = NB_POST_BLOCK;
#ifndef NB_HIT_GROUP
#define NB_HIT_GROUP 512
#endif
#ifndef NB_HIT_BLOCK
#define NB_HIT_BLOCK 128
#endif
constexpr int kHitGroup = NB_HIT_GROUP;   // events per k_hits group
constexpr int kHitBlock = NB_HIT_BLOCK;   // threads per k_hits block
#define NB_AREA_FX 4294967296.0  // 2^32: S1 hit areas are summed in units of 2^-32 phe
#define NB_AREA_MAGIC 6755399441055744.0           // 1.5 * 2^52: ulp = 1 for |a 2^32| < 2^51
#define NB_AREA_MAGIC_BITS 0x4338000000000000ull  // its bit pattern

// ---- batched parameter sweep (BASELINE config 4) -----------------------------------------------------------
// n_points grid points (detector, model, run constants, source: one RunParams each, in global memory), every point
// simulating the same event ids [first_event, first_event + n_per_point); one band per point.  The launch walks a
// flattened range of SLOTS: slot s = point s / n_pad, event index s % n_pad (n_pad = n_per_point rounded up to a
// multiple of 1024, so that no block tile or k_hits group straddles two points; the padding slots are idle).
// A block copies its point's RunParams to shared memory and runs the very device code of the single-point kernels
// on it, so a point's events are bit for bit those of nb200_run with the same parameters.
struct SweepArgs {
  const RunParams* params;  // [n_points], device
  unsigned long long* bands;  // [n_points][band_words], device
  uint64_t band_words;
  uint64_t n_per_point, n_pad;
  uint64_t slot0, n_slots;  // this launch: slots [slot0, slot0 + n_slots); workspace index = slot - slot0
  uint64_t first_event;
  uint32_t rk[20];  // Philox round keys of the sweep's seed (one seed for all points): constant-bank operands
  Work work;
};
// RunParams of `point` into this block's shared copy (all threads; barriers inside)
NB_D void sweep_load_params(RunParams* sP, const RunParams* gP, uint32_t point) {
  __syncthreads();  // nobody still reads the previous point's copy
  const uint32_t* src = reinterpret_cast<const uint32_t*>(gP + point);
  uint32_t* dst = reinterpret_cast<uint32_t*>(sP);
  static_assert(sizeof(RunParams) % 4 == 0, "RunParams is copied word by word");
  for (uint32_t i = threadIdx.x; i < sizeof(RunParams) / 4; i += blockDim.x) dst[i] = src[i];
  __syncthreads();
}

// SoA column store: double, or float when the caller asked for the packed record (write_soa == 2)
NB_D void soa_put(double* col, uint64_t i, double v, bool f32) {
  if (f32)
    reinterpret_cast<float*>(col)[i] = float(v);
  else
    col[i] = v;
}

// one event of k_pre: W index idx (workspace), gidx (index within the call: lists, SoA columns), event id ev
NB_D void pre_body(const uint32_t* rk, const RunParams& P, const Work& W, uint64_t idx, uint64_t gidx, uint64_t ev, bool active,
                   int* __restrict__ err_flag) {
  const nb200_detector& det = P.det;
  const bool cli = !P.vec_mode;
  Stream g;
  g.init(rk, ev, STREAM_MAIN);
  double px, py, pz, sx, sy, field = 0., dt = 0., vD = P.rc.vD;
  Yields y{0., 0., 0., 0., 0., 0.};
  Quanta q{0, 0, 0, 0, 0., 0.};
  int err = 0;
  double keV = sample_energy(g, P, gidx);
  NB_LOCKSTEP();
  const nb200_source& S = P.src;
  const bool nonuni = (S.inField == -1.);
  const bool listed = (S.kind == NB200_SRC_LIST && S.x != nullptr);
  if (listed) {
    px = S.x[gidx];
    py = S.y[gidx];
    pz = S.z[gidx];
  } else {
    px = S.xyz[0];
    py = S.xyz[1];
    pz = S.xyz[2];
  }
  const bool ranXY = (S.fixed_pos == 0 || S.fixed_pos == 3) && !listed;
  const bool ranZ = (S.fixed_pos == 0 || S.fixed_pos == 2) && !listed;
  for (int guard = 0; guard < 10000; ++guard) {  // Z_NEW loop execNEST.cpp:806-967
    if (ranZ) pz = P.z_direct ? fma(P.z_hi - P.z_lo, g.uniform(), P.z_lo) : 0. + (det.TopDrift - 0.) * g.uniform();
    if (ranXY && (guard == 0 || S.fixed_pos == 0)) {
      double r = det.radius * sqrt(g.uniform());
      double sn, cs;
      uint32_t pa, pb;
      g.next2(pa, pb);  // phi = 2 pi u, u from 32 bits
      sincos_bits(pa, sn, cs);
      px = r * cs;
      py = r * sn;
    }
    if (cli)
      field = nonuni ? fit_ef(det.FitEF, px, py, pz) : S.inField;
    else
      field = (S.inField < 0.0) ? fit_ef(det.FitEF, px, py, pz) : S.inField;
    if (cli) {
      if (nonuni) vD = P.vtable[min(max(int(floor(pz / P.ana.z_step + 0.5)), 0), P.vtable_n - 1)];
    } else if (S.inField < 0.0)
      vD = drift_velocity_dev(P, field);
    dt = (det.TopDrift - pz) / vD;
    if (cli && ranZ && (dt > det.dt_max || dt < det.dt_min) && field >= NB_FIELD_MIN) continue;
    break;
  }
  if (cli) {  // execNEST.cpp:975-992
    if (pz <= 0) pz = P.ana.z_step;
    if ((pz > (det.TopDrift + P.ana.z_step) || dt < 0.0) && field >= NB_FIELD_MIN) {
      dt = 0.0;
      pz = det.TopDrift - P.ana.z_step;
    }
  }
  NB_LOCKSTEP();
  // yields + quanta: execNEST.cpp:1031-1077 / FullCalculation NEST.cpp:22-42
  const int sp = P.species;
  const bool nrlike = sp <= NB200_Cf || sp == NB200_atmNu;
  if (!cli || keV > 0.001 * P.rc.Wq_eV) {
    int A2 = 131;
    if (P.model.er_model < 0 && ((nrlike && nearly_equal(P.model.massNum, 0.)) || sp == NB200_ion))
      A2 = select_xe_atom(g);
    if (P.model.er_model == 0)
      y = yield_beta(keV, P.rc.rho, field, det, P.model.ERYieldsParam, P.model.multFact, &P.hoist);
    else if (P.model.er_model == 1)
      y = yield_gamma(keV, P.rc.rho, field, det, P.model.multFact, &P.hoist);
    else if (P.model.er_model == 2)
      y = yield_er_weighted(keV, P.rc.rho, field, det, P.model.ERYieldsParam, &P.hoist);
    else {
      double u01 = 0.5;
      if (sp == NB200_Kr83m && !nearly_equal(P.model.massNum, P.model.atomNum)) u01 = g.uniform();
      y = get_yields(sp, keV, P.rc.rho, field, P.model.massNum, P.model.atomNum, A2, det, P.model, &err, u01,
                     &P.hoist);
    }
  }
  NB_LOCKSTEP();
  if (!cli || keV > 0.001 * P.rc.Wq_eV) q = get_quanta(g, ev, y, P.rc.rho, det, P.model, &err, &P.hoist);
  NB_LOCKSTEP();
  if (cli && (det.noiseBaseline[2] != 0. || det.noiseBaseline[3] != 0.))
    q.electrons += to_int_round(gauss(g, det.noiseBaseline[2], det.noiseBaseline[3], false));
  // position smearing execNEST.cpp:1090-1099
  sx = px;
  sy = py;
  if (cli) {
    double Nphd_S2 = fabs(P.rc.g2_params[3]) * q.electrons * exp(-dt / det.eLife_us);
    if (!P.ana.MCtruthPos && Nphd_S2 > NB_PHE_MIN) xy_resolution(g, det, px, py, Nphd_S2, sx, sy);
  }
  NB_LOCKSTEP();
  S1Pre pre = s1_prelude(rk, ev, P, q.photons, px, py, pz, sx, sy, pz, keV);
  NB_LOCKSTEP();
  if (!active) return;

  if (P.write_soa && P.soa.deltaT_ns) soa_put(P.soa.deltaT_ns, gidx, y.dT, P.write_soa == 2);
  W.keV[idx] = keV;
  W.px[idx] = px;
  W.py[idx] = py;
  W.pz[idx] = pz;
  W.sx[idx] = sx;
  W.sy[idx] = sy;
  W.field[idx] = field;
  W.dt[idx] = dt;
  W.yNph[idx] = y.Nph;
  W.yNe[idx] = y.Ne;
  W.yL[idx] = y.L;
  W.posDepSm[idx] = pre.posDepSm;
  W.fitSm[idx] = pre.fitSm;
  W.photons[idx] = q.photons;
  W.electrons[idx] = q.electrons;
  W.ions[idx] = q.ions;
  W.excitons[idx] = q.excitons;
  W.nHits[idx] = pre.full ? pre.nHits : -(pre.nHits + 1);  // negative: Parametric S1 in k_post
  W.nphe[idx] = pre.nDP;  // in: double-phe hits; k_hits replaces it by Nphe = nHits + nDP
  if (err) atomicOr(err_flag, err);
}

__global__ void __launch_bounds__(kPreBlock, NB_PRE_MINB)
    k_pre(const __grid_constant__ RunParams P, int* __restrict__ err_flag) {
  // The per-event code is long and straight-line: its cost is instruction fetch, not arithmetic.
  // NB_LOCKSTEP() barriers keep the warps of a block inside the same few KB of code, so that one
  // warp's instruction-cache fill serves the others (all threads reach every barrier: the loop
  // bound is block-uniform and inactive lanes only skip the stores).
  for (uint64_t base = blockIdx.x * uint64_t(kPreBlock); base < P.n_events;
       base += uint64_t(gridDim.x) * kPreBlock) {
    const bool active = base + threadIdx.x < P.n_events;
    const uint64_t idx = active ? base + threadIdx.x : P.n_events - 1;
    const uint64_t gidx = P.chunk_off + idx;
    pre_body(P.rk, P, P.work, idx, gidx, P.first_event + gidx, active, err_flag);
  }
}

__global__ void __launch_bounds__(kPreBlock, NB_PRE_MINB)
    k_pre_sweep(const __grid_constant__ SweepArgs A, int* __restrict__ err_flag) {
  __shared__ __align__(16) RunParams sP;
  uint32_t cur = 0xffffffffu;
  // a block takes a contiguous range of tiles (not a grid stride): it then stays on one grid point for many tiles
  // and reloads its parameter copy only when the range crosses into the next point
  const uint64_t n_tiles = (A.n_slots + kPreBlock - 1) / kPreBlock;
  const uint64_t t0 = n_tiles * blockIdx.x / gridDim.x, t1 = n_tiles * (blockIdx.x + 1) / gridDim.x;
  for (uint64_t t = t0; t < t1; ++t) {
    const uint64_t base = t * kPreBlock;
    const uint64_t slot = A.slot0 + base;  // first slot of this tile: multiple of kPreBlock
    const uint32_t point = uint32_t(slot / A.n_pad);
    if (point != cur) {
      sweep_load_params(&sP, A.params, point);
      cur = point;
    }
    const uint64_t i = slot - uint64_t(point) * A.n_pad + threadIdx.x;  // event index within the point
    const bool active = i < A.n_per_point;
    const uint64_t ie = active ? i : A.n_per_point - 1;
    pre_body(A.rk, sP, A.work, base + threadIdx.x, ie, A.first_event + ie, active, err_flag);
  }
}

// -------------------------------------------------------------------------------------------
// k_hits: the per-PMT-hit loop of GetS1, NEST.cpp:1367-1429
//
// Work item = SLOT = one Philox block = THREE photo-electrons (27 bits of abscissa, 4 of truncation pre-test and
// 11 of ziggurat layer each: nb_rng.cuh pe_fields).  The hits of an event are exchangeable (iid draws, and the
// event observables -- spike count, summed area, Nphe -- are symmetric in them), so instead of deciding hit by
// hit whether it carries a second photo-electron (NEST.cpp:1391), k_pre draws their number nD ~
// Binomial(nHits, P_dphe) once, and the nD double hits and nS = nHits - nD single hits are dealt into slots:
//   slot j < min(nD, nS)        one double hit (photo-electrons 0 + 1, summed before the sPEthr test) and one
//                               single hit (photo-electron 2)
//   nD > nS: slots nS .. nD-1   one double hit (photo-electron 2 unused)
//   nD <= nS: slots from nD on  three single hits, the last slot one to three
// Nphe = nHits + nD.  No second-photo-electron queue, no per-hit Bernoulli; 0.39 Philox blocks per hit for LUX
// (P_dphe = 0.173) against 0.59 with two photo-electrons per block.
//
// A slot whose photo-electrons all land in their ziggurat layer's core above the truncation
// pre-test (93 % of slots for LUX) is finished in the main loop.  Any other slot -- wedge / tail
// of the ziggurat, exact truncation test, and every slot of an event that needs efficiency or
// coincidence-window draws -- is queued per warp as (event, slot) and redone in full, 32 at a
// time, by the out-of-line finisher: rare per lane but frequent per warp.
struct SlotEntry {
  uint32_t j, e;
};
// What the deferred finisher needs to know of a queued slot's event travels in the top bits of e (the workspace index
// takes the low 26: chunks are <= 2^25 events): valid photo-electrons (slot_valid), double hit, special event.  Without
// them the finisher had to fetch the event's hit counts before it could start on the slot -- a second dependent trip
// to device memory for a kernel that does nothing but wait on it; now only special events' counts are fetched.
constexpr uint32_t kEntIdxBits = 26, kEntIdxMask = (1u << kEntIdxBits) - 1u;
constexpr uint32_t kEntDbl = 1u << 28, kEntSpecial = 1u << 29;
NB_D uint32_t ent_flags(int nv, bool dbl, bool special) {
  return (uint32_t(nv) << kEntIdxBits) | (dbl ? kEntDbl : 0u) | (special ? kEntSpecial : 0u);
}
NB_D int ent_nv(uint32_t e) { return int((e >> kEntIdxBits) & 3u); }
constexpr int kSlotCap = 64;  // <= 31 left over + 32 pushed per iteration
constexpr size_t kHitDynSmem = NB_ZIG_N * (sizeof(double) + sizeof(uint32_t));  // ziggurat tables (opt-in: > 48 KB total)

// slots of an event with nh hits of which nd carry a second photo-electron
NB_D int hit_slots(int nh, int nd) {
  const int ns = nh - nd;
  return nd + (ns > nd ? (ns - nd + 2) / 3 : 0);
}
// valid photo-electrons of slot j: 3 = all, 2 = the first two (a double hit alone, or two single hits), 1 = one single hit
NB_D int slot_valid(int j, int nd, int ns) {
  return j < nd ? (j < ns ? 3 : 2) : min(3, (ns - nd) - 3 * (j - nd));
}

// 64-bit add to shared memory as two 32-bit atomics with the carry passed on by whoever wraps the low word
// (the native shared-memory atomic is 32 bits wide; a 64-bit atomicAdd compiles to a compare-and-swap loop,
// 6 % of the kernel's stall samples).  The adds to a word are totally ordered, so every wrap of the low
// word is seen by exactly one adder and the pair ends up holding the exact 64-bit sum.
NB_D void smem_add64(unsigned long long* p, unsigned long long v) {
  unsigned int* w = reinterpret_cast<unsigned int*>(p);
  const unsigned int lo = (unsigned int)v;
  const unsigned int old = atomicAdd(w, lo);
  const unsigned int hi = (unsigned int)(v >> 32) + ((old + lo) < old ? 1u : 0u);
  if (hi) atomicAdd(w + 1, hi);
}

NB_D void hit_commit(const RunParams& P, unsigned long long* s_area, unsigned int* s_spike, int slot, double a,
                     bool pass_coin) {
  if (a > P.det.sPEthr && pass_coin) {
    smem_add64(&s_area[slot], (unsigned long long)__double2ll_rn(a * NB_AREA_FX));
    atomicAdd(&s_spike[slot], 1u);
  }
}

// the standard normal of photo-electron `which` of slot j: ziggurat core, wedge or tail.  r: the slot's
// STREAM_HIT block, t: its STREAM_HIT2 block (word w: first wedge test of photo-electron 0).  A: the abscissa
// word (pre-test bits); fast: the draw was a layer core (what the main loop of k_hits finishes by itself)
NB_D double pe_normal(const u128& r, const u128& t, int which, uint32_t e0, uint32_t e1, uint32_t j, uint32_t k0,
                      uint32_t k1, const uint32_t* zk, const double* zw, uint32_t& A, bool& fast) {
  uint32_t L;
  pe_fields(r, which, A, L);
  double z;
  fast = zig_fast(A, L, zk, zw, z);
  if (!fast) z = zig_finish(A, L, t.w, which == 0, e0, e1, 3u * j + uint32_t(which), STREAM_ZIG, k0, k1);
  return z;
}

// status of a photo-electron of a queued slot while it is being finished
enum { PE_FINAL = 0, PE_NEED_ZIG = 1, PE_NEED_REDRAW = 2 };

// a photo-electron after its normal z: area sv, truncation pre-test (pre-test bits of the abscissa word A) and exact
// residual test (uniform: word `which` of t).  PE_FINAL, or PE_NEED_REDRAW when the pair (S, T) has to be redrawn.
NB_D int pe_after_normal(const RunParams& P, const u128& t, int which, uint32_t A, double z, double& sv) {
  sv = fma(P.pe_sig, z, P.pe_mu);
  if (sv <= P.pe_cut && (sv < P.pe_sfast || (A & NB_PE_PRETEST_MASK) == NB_PE_PRETEST_MASK)) {
    const uint32_t uw = which == 0 ? t.x : (which == 1 ? t.y : t.z);
    if (!pe_accept_exact(P, sv, u32d(uw))) return PE_NEED_REDRAW;
  }
  return PE_FINAL;
}
NB_D double pe_redraw_of(const RunParams& P, uint64_t ev, uint32_t j, int which) {
  return pe_redraw(P.seed, ev, j * 96u + uint32_t(which) * 32u, P.pe_mu, P.pe_sig, P.pe_cut, P.pe_tau, P.pe_rho, P.pe_m0);
}

// one photo-electron in full: normal, truncation pre-test, exact residual test, joint redraw
NB_D double pe_full(const RunParams& P, const u128& r, const u128& t, int which, uint64_t ev, uint32_t j,
                    const uint32_t* zk, const double* zw) {
  uint32_t A;
  bool fast;
  const double z = pe_normal(r, t, which, uint32_t(ev), uint32_t(ev >> 32), j, uint32_t(P.seed), uint32_t(P.seed >> 32), zk,
                             zw, A, fast);
  double sv;
  if (pe_after_normal(P, t, which, A, z, sv) == PE_NEED_REDRAW) sv = pe_redraw_of(P, ev, j, which);
  return sv;
}

// One queued slot in full -- everything the reference does for its hits (wedge / tail of the ziggurat, exact
// truncation test, joint redraw, efficiency and coincidence-window draws).  commit(area, pass_coin) books one hit.
// The photo-electrons are walked by a rolled loop: one call site each for the wedge / tail completion and the joint
// redraw, so that the lanes of a warp that need one -- a few per pass, for different photo-electrons -- go through it
// together.
template <class Commit>
NB_D void finish_slot(const RunParams& P, int nh, int nd, uint32_t j, uint64_t ev, const uint32_t* zk, const double* zw,
                      bool live, const Commit& commit) {
  const nb200_detector& det = P.det;
  const bool dbl = int(j) < nd;
  const int nv = live ? slot_valid(int(j), nd, nh - nd) : 0;
  const uint32_t e0 = uint32_t(ev), e1 = uint32_t(ev >> 32);
  const u128 r = philox4x32_10_rk(e0, e1, j, STREAM_HIT, P.rk);
  const u128 t = philox4x32_10_rk(e0, e1, j, STREAM_HIT2, P.rk);
  // efficiency and coincidence-window draws (sPEeff < 1 detectors; events of <= coinLevel hits)
  const bool special = det.sPEeff < 1. || !(nh > det.coinLevel);
  const double eff = s1_eff(det, nh);
  u128 c{0u, 0u, 0u, 0u}, c2{0u, 0u, 0u, 0u};
  if (live && special) {
    c = philox4x32_10_rk(e0, e1, j, STREAM_HIT3, P.rk);
    if (!(nh > det.coinLevel)) c2 = philox4x32_10_rk(e0, e1, j, STREAM_HIT4, P.rk);
  }
  const bool coin_draws = special && !(nh > det.coinLevel);  // NEST.cpp:1422-1424, one draw per hit
  double first = 0.;   // first photo-electron of a double hit
  bool lost0 = false;
#pragma unroll 1
  for (int which = 0; which < 3; ++which) {
    if (which < nv) {
      double sv = pe_full(P, r, t, which, ev, j, zk, zw);
      if (special && eff < 1.) {  // NEST.cpp:1381-1386, 1403: a second photo-electron is lost only with the first
        const bool lost = u32d(which == 0 ? c.x : (which == 1 ? c.y : c.z)) > eff;
        if (which == 0) lost0 = lost;
        if ((dbl && which == 1) ? (lost && lost0) : lost) sv = 0.;
      }
      if (dbl && which == 0) {
        first = sv;
      } else {
        // hit index of this photo-electron within the slot: the double hit is hit 0, the single after it hit 1
        const int hit = dbl ? which - 1 : which;
        const uint32_t cw = hit == 0 ? c.w : (hit == 1 ? c2.x : c2.y);
        const bool coin = !coin_draws || u32d(cw) > P.coin_u_min;
        commit((dbl && which == 1) ? first + sv : sv, coin);
      }
    }
  }
}

// up to 32 queued slots finished inside k_hits (out of line: its register needs must not shape the main loop).
// Only the overflow path of the deferred queue (below) and detectors whose every slot is special end up here.
__device__ __noinline__ void hits_finish_slow(const RunParams& P, const int* s_nh, const int* s_nd,
                                             unsigned long long* s_area, unsigned int* s_spike, uint64_t ev0,
                                             const SlotEntry* q, int count, int lane, const uint32_t* zk,
                                             const double* zw) {
  const bool live = lane < count;
  const SlotEntry en = live ? q[lane] : SlotEntry{0u, 0u};
  const int e = int(en.e & kEntIdxMask);
  finish_slot(P, s_nh[e], s_nd[e], en.j, ev0 + uint64_t(e), zk, zw, live,
              [&](double a, bool coin) { hit_commit(P, s_area, s_spike, e, a, coin); });
  __syncwarp();
}

// ---- deferred finisher --------------------------------------------------------------------------------------
// Measured (profiles/r02_k_hits_deferred.txt): finishing the queued slots inside k_hits cost 36 % of its time for
// 6.5 % of the slots -- the finisher ran at 16 of 32 lanes with its rare sub-paths at 1-2, and stalled a main loop
// that has only four warps per scheduler to hide anything behind.  k_hits therefore only WRITES its per-warp queue,
// 32 entries at a time, to a queue in device memory (one atomicAdd per 32 entries), leaves the event sums as
// integers in the workspace, and a second kernel finishes the queued slots -- one per thread, every lane busy --
// adding their hits to the event sums with global atomics (integer adds: order does not matter, results stay
// bit-identical for any batch split).  Entries that do not fit the queue are finished in place as before.
struct HitQueue {
  SlotEntry* entries;   // e = workspace index of the event, j = slot
  unsigned int* count;  // entries pushed (may exceed cap: the excess was finished in place)
  unsigned int cap;
};

// A queued slot is slow for one of its photo-electrons only, and for one of three rare reasons: the exact
// truncation test (most), a wedge / tail completion, a joint redraw.  Finishing every slot straight through
// (finish_slot) makes a warp execute the redraw and the completion in nearly every pass for the one or two lanes
// that need them (measured: 1500 warp-instructions per 32 slots).  The finisher therefore works in stages, each
// dense: stage 0 does what every queued slot needs (its two Philox blocks, the layer-core normals, the pre-tests
// and exact tests) and books the slots that this settles; the others are parked, with what is already known, in a
// per-warp queue and completed 32 at a time -- every lane then has a pending photo-electron.  Slots of events that
// need efficiency / coincidence draws go through finish_slot, also 32 at a time.  Every value is computed by the
// functions finish_slot uses, with the same streams: which path finishes a slot does not change its result.
struct SlowItem {
  uint32_t e, j;
  double v[3];
  uint32_t st;  // status of photo-electron k in bits [2k+1 : 2k]
  uint32_t pad;
};
#ifndef NB_FIN_MINB
#define NB_FIN_MINB 8
#endif

constexpr int kFinBlock = 128;
constexpr int kFinWarps = kFinBlock / 32;

struct FinCtx {
  const RunParams* P;
  const Work* W;
  uint64_t ev;
  int nh, nd;  // fin_counts only
};
// where a queued slot's event lives: parameters, workspace, event id (no memory access)
template <bool SWEEP>
NB_D FinCtx fin_resolve(const RunParams* P0, const SweepArgs* Ap, uint64_t idx) {
  FinCtx c;
  if (SWEEP) {
    const uint64_t slot = Ap->slot0 + idx;
    const uint32_t point = uint32_t(slot / Ap->n_pad);
    c.P = Ap->params + point;  // device memory: the lanes of a warp may belong to different grid points
    c.W = &Ap->work;
    c.ev = Ap->first_event + (slot - uint64_t(point) * Ap->n_pad);
  } else {
    c.P = P0;
    c.W = &P0->work;
    c.ev = P0->first_event + P0->chunk_off + idx;
  }
  c.nh = c.nd = 0;
  return c;
}
// the event's hit counts (special events only: the others' entries carry what the finisher needs)
NB_D void fin_counts(FinCtx& c, uint64_t idx) {
  int nh = c.W->nHits[idx];
  c.nh = nh > 0 ? nh : 0;
  c.nd = c.W->nphe[idx] - c.nh;  // k_hits left Nphe = nHits + nD there
}
// the hits of a finished slot (no efficiency / coincidence draws) into the event's sums
NB_D void fin_commit(const RunParams& P, const Work& W, uint64_t idx, bool dbl, int nv, double v0, double v1, double v2) {
  const double thr = P.det.sPEthr;
  const double a0 = dbl ? v0 + v1 : v0;
  const double a1 = dbl ? (nv >= 3 ? v2 : -DBL_MAX) : (nv >= 2 ? v1 : -DBL_MAX);
  const double a2 = (!dbl && nv >= 3) ? v2 : -DBL_MAX;
  unsigned long long fx = 0ull;
  int spikes = 0;
  if (a0 > thr) fx += (unsigned long long)__double2ll_rn(a0 * NB_AREA_FX), ++spikes;
  if (a1 > thr) fx += (unsigned long long)__double2ll_rn(a1 * NB_AREA_FX), ++spikes;
  if (a2 > thr) fx += (unsigned long long)__double2ll_rn(a2 * NB_AREA_FX), ++spikes;
  if (spikes) {
    atomicAdd(&reinterpret_cast<unsigned long long*>(W.area)[idx], fx);
    atomicAdd(&W.spike[idx], spikes);
  }
}

template <bool SWEEP>
NB_D void hits_finish_body(const RunParams* P0, const SweepArgs* Ap, const HitQueue Q) {
  // the kernel waits on device memory (queue entries, the events' hit counts, the atomics): it wants warps, not
  // shared-memory tables -- the ziggurat layers are read through L1 from their global copies
  const uint32_t* const s_zk = g_zig_k;
  const double* const s_zw = g_zig_w;
  const uint32_t* const rk = SWEEP ? Ap->rk : P0->rk;  // kernel-parameter constants either way
  __shared__ __align__(8) SlowItem s_slow[kFinWarps][64];
  __shared__ __align__(8) SlotEntry s_spec[kFinWarps][64];
  const unsigned int n = min(*Q.count, Q.cap);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const unsigned lt = (1u << lane) - 1u;
  const unsigned int nwarps = gridDim.x * kFinWarps, warp_id = blockIdx.x * kFinWarps + wid;

  // stage 1: up to 32 parked slots, each with at least one pending photo-electron
  auto run_slow = [&](const SlowItem* items, int count) {
    const bool live = lane < count;
    SlowItem it = live ? items[lane] : SlowItem{0u, 0u, {0., 0., 0.}, 0u, 0u};
    const FinCtx c = fin_resolve<SWEEP>(P0, Ap, it.e & kEntIdxMask);
    const RunParams& P = *c.P;
    const uint32_t e0 = uint32_t(c.ev), e1 = uint32_t(c.ev >> 32);
    const u128 r = philox4x32_10_rk(e0, e1, it.j, STREAM_HIT, rk);
    const u128 t = philox4x32_10_rk(e0, e1, it.j, STREAM_HIT2, rk);
    double v0 = it.v[0], v1 = it.v[1], v2 = it.v[2];
    uint32_t st = live ? it.st : 0u;
#pragma unroll 1
    while (__any_sync(0xffffffffu, st != 0u)) {
      if (st != 0u) {
        const int which = (st & 3u) ? 0 : ((st & 12u) ? 1 : 2);
        int s1 = int((st >> (2 * which)) & 3u);
        double sv = 0.;
        if (s1 == PE_NEED_ZIG) {
          uint32_t A, L;
          pe_fields(r, which, A, L);
          const double z = zig_finish(A, L, t.w, which == 0, e0, e1, 3u * it.j + uint32_t(which), STREAM_ZIG, uint32_t(P.seed),
                                      uint32_t(P.seed >> 32));
          s1 = pe_after_normal(P, t, which, A, z, sv);
        }
        if (s1 == PE_NEED_REDRAW) sv = pe_redraw_of(P, c.ev, it.j, which);
        if (which == 0)
          v0 = sv;
        else if (which == 1)
          v1 = sv;
        else
          v2 = sv;
        st &= ~(3u << (2 * which));
      }
    }
    if (live) fin_commit(P, *c.W, it.e & kEntIdxMask, (it.e & kEntDbl) != 0u, ent_nv(it.e), v0, v1, v2);
    __syncwarp();
  };
  // slots of events that draw efficiencies / coincidence windows: the full per-slot code
  auto run_special = [&](const SlotEntry* ents, int count) {
    const bool live = lane < count;
    const SlotEntry en = live ? ents[lane] : SlotEntry{0u, 0u};
    const uint64_t idx = en.e & kEntIdxMask;
    FinCtx c = fin_resolve<SWEEP>(P0, Ap, idx);
    fin_counts(c, idx);
    const RunParams& P = *c.P;
    const Work& W = *c.W;
    finish_slot(P, c.nh, c.nd, en.j, c.ev, s_zk, s_zw, live, [&](double a, bool coin) {
      if (a > P.det.sPEthr && coin) {
        atomicAdd(&reinterpret_cast<unsigned long long*>(W.area)[idx], (unsigned long long)__double2ll_rn(a * NB_AREA_FX));
        atomicAdd(&W.spike[idx], 1);
      }
    });
    __syncwarp();
  };

  int nslow = 0, nspec = 0;  // warp-uniform fills of the two per-warp queues
  for (unsigned int base = warp_id * 32u; base < n; base += nwarps * 32u) {
    const unsigned int i = base + lane;
    const SlotEntry en = i < n ? Q.entries[i] : SlotEntry{0xffffffffu, 0u};
    const bool live = en.j != 0xffffffffu;  // (sentinel: an entry reserved but never written, see k_hits)
    const FinCtx c = fin_resolve<SWEEP>(P0, Ap, live ? (en.e & kEntIdxMask) : 0u);
    const RunParams& P = *c.P;
    const bool special = live && (en.e & kEntSpecial) != 0u;
    bool slow = false;
    SlowItem it{en.e, en.j, {0., 0., 0.}, 0u, 0u};
    if (live && !special) {
      const uint32_t e0 = uint32_t(c.ev), e1 = uint32_t(c.ev >> 32);
      const u128 r = philox4x32_10_rk(e0, e1, en.j, STREAM_HIT, rk);
      const u128 t = philox4x32_10_rk(e0, e1, en.j, STREAM_HIT2, rk);
      const int nv = ent_nv(en.e);
#pragma unroll
      for (int which = 0; which < 3; ++which) {
        if (which < nv) {
          uint32_t A, L;
          pe_fields(r, which, A, L);
          double z, sv = 0.;
          int s1 = PE_NEED_ZIG;
          if (zig_fast(A, L, s_zk, s_zw, z)) s1 = pe_after_normal(P, t, which, A, z, sv);
          it.v[which] = sv;
          it.st |= uint32_t(s1) << (2 * which);
        }
      }
      if (it.st == 0u)
        fin_commit(P, *c.W, en.e & kEntIdxMask, (en.e & kEntDbl) != 0u, nv, it.v[0], it.v[1], it.v[2]);
      else
        slow = true;
    }
    const unsigned ms = __ballot_sync(0xffffffffu, special);
    if (ms) {
      if (special) s_spec[wid][nspec + __popc(ms & lt)] = en;
      nspec += __popc(ms);
      __syncwarp();
      if (nspec >= 32) {
        nspec -= 32;
        run_special(&s_spec[wid][nspec], 32);
      }
    }
    const unsigned mq = __ballot_sync(0xffffffffu, slow);
    if (mq) {
      if (slow) s_slow[wid][nslow + __popc(mq & lt)] = it;
      nslow += __popc(mq);
      __syncwarp();
      if (nslow >= 32) {
        nslow -= 32;
        run_slow(&s_slow[wid][nslow], 32);
      }
    }
  }
  if (nspec > 0) run_special(&s_spec[wid][0], nspec);
  if (nslow > 0) run_slow(&s_slow[wid][0], nslow);
}

__global__ void __launch_bounds__(kFinBlock, NB_FIN_MINB) k_hits_finish(const __grid_constant__ RunParams P, const HitQueue Q) {
  hits_finish_body<false>(&P, nullptr, Q);
}
__global__ void __launch_bounds__(kFinBlock, NB_FIN_MINB) k_hits_finish_sweep(const __grid_constant__ SweepArgs A, const HitQueue Q) {
  hits_finish_body<true>(nullptr, &A, Q);
}

// `count` (<= 32) entries of a warp's queue to the deferred finisher's queue (e made a workspace index); false when
// they have to be finished in place: detectors whose every slot is special, or a full queue (the part of a
// reservation that still lies below the capacity is filled with sentinels the finisher skips)
NB_D bool queue_push(const HitQueue& Q, bool in_place, const SlotEntry* q, int count, int lane, uint32_t gbase) {
  if (in_place) return false;
  unsigned int at = 0;
  if (lane == 0) at = atomicAdd(Q.count, (unsigned int)count);
  at = __shfl_sync(0xffffffffu, at, 0);
  const bool fits = at + (unsigned int)count <= Q.cap;
  if (lane < count && at + lane < Q.cap) {
    SlotEntry en = q[lane];
    en.e += gbase;
    if (!fits) en.j = 0xffffffffu;
    Q.entries[at + lane] = en;
  }
  return fits;
}

// the body of k_hits / k_hits_sweep.  SWEEP = false: P0 is the launch's RunParams (kernel parameter, constant bank).
// SWEEP = true: a group's RunParams is copied from A.params[point] into the shared sP whenever the block moves to
// another grid point; n_events and the event ids come from the sweep's slot layout.
template <bool SWEEP>
NB_D void hits_body(const RunParams* P0, const SweepArgs* Ap, RunParams* sP, unsigned int* __restrict__ next_group,
                    const HitQueue Q) {
  const RunParams& P = SWEEP ? *sP : *P0;
  const SweepArgs& A = *Ap;  // dereferenced in the sweep instantiation only
  const uint32_t* const rk = SWEEP ? Ap->rk : P0->rk;  // kernel-parameter constants either way
  // the flattened work items: the group's events that have slots, in order -- {end of the event's slots in the
  // group's slot numbering, event index in the group | its special flag << 31, double-phe hits, single hits};
  // entry [number of such events] is a sentinel no slot index reaches
  __shared__ __align__(16) uint4 s_rec[kHitGroup + 1];
  __shared__ int s_nh[kHitGroup];         // hits of the event
  __shared__ int s_nd[kHitGroup];         // of which double-phe
  __shared__ unsigned long long s_area[kHitGroup];   // the main loop's sums
  __shared__ unsigned int s_spike[kHitGroup];
  __shared__ unsigned long long s_farea[kHitGroup];  // what hits_finish_slow adds (kept apart: see the flush below)
  __shared__ unsigned int s_fspike[kHitGroup];
  __shared__ int2 s_warp[32];
  __shared__ int s_tot[2];  // slots, events with slots
  __shared__ __align__(8) SlotEntry s_q[kHitBlock / 32][kSlotCap];
  extern __shared__ __align__(16) double s_zw[];  // kHitDynSmem: zw[NB_ZIG_N] then zk[NB_ZIG_N]
  uint32_t* const s_zk = reinterpret_cast<uint32_t*>(s_zw + NB_ZIG_N);
  for (int i = threadIdx.x; i < NB_ZIG_N; i += kHitBlock) {
    s_zw[i] = g_zig_w[i];
    s_zk[i] = g_zig_k[i];
  }
  constexpr int PER = kHitGroup / kHitBlock;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const nb200_detector& det = P.det;
  const Work& W = SWEEP ? A.work : P0->work;
  const uint64_t n_launch = SWEEP ? A.n_slots : P0->n_events;  // workspace entries of this launch
  const uint64_t n_groups = (n_launch + kHitGroup - 1) / kHitGroup;
  uint32_t cur_point = 0xffffffffu;
  uint32_t q_addr = uint32_t(__cvta_generic_to_shared(&s_q[wid][0]));
  asm volatile("" : "+r"(q_addr));  // opaque: keep it in a register instead of recomputing it at every push

  // groups are handed out by a device counter (zeroed before the launch): their cost follows their hit
  // counts, so a static stride leaves the last wave ragged
  __shared__ unsigned int s_grp;
  for (;;) {
    if (tid == 0) s_grp = atomicAdd(next_group, 1u);
    __syncthreads();
    const uint64_t grp = s_grp;
    if (grp >= n_groups) break;
    const uint64_t gbase = grp * kHitGroup;
    uint64_t ev0, n_valid = n_launch;  // event id of workspace entry gbase; entries >= n_valid are idle
    if (SWEEP) {
      const uint64_t slot = A.slot0 + gbase;
      const uint32_t point = uint32_t(slot / A.n_pad);
      if (point != cur_point) {
        sweep_load_params(sP, A.params, point);
        cur_point = point;
      }
      const uint64_t i0 = slot - uint64_t(point) * A.n_pad;
      ev0 = A.first_event + i0;
      n_valid = i0 < A.n_per_point ? min(n_launch, gbase + (A.n_per_point - i0)) : gbase;
    } else
      ev0 = P0->first_event + P0->chunk_off + gbase;
    const bool all_special = det.sPEeff < 1.;  // efficiency draws: every slot goes through the finisher
    // slot counts of my PER consecutive events -> exclusive prefix over the group, and the list of events with slots
    int cnt[PER], nhs[PER], nds[PER];
    int2 mine = make_int2(0, 0);
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      const uint64_t idx = gbase + uint64_t(tid) * PER + k;
      int nh = (idx < n_valid) ? W.nHits[idx] : 0;  // negative: Parametric S1 (k_post)
      nh = nh > 0 ? nh : 0;
      int nd = (idx < n_valid) ? W.nphe[idx] : 0;  // k_pre left the double-phe count here
      nd = nd < 0 ? 0 : (nd > nh ? nh : nd);
      cnt[k] = hit_slots(nh, nd);
      nhs[k] = nh;
      nds[k] = nd;
      mine.x += cnt[k];
      mine.y += cnt[k] > 0 ? 1 : 0;
      s_nh[tid * PER + k] = nh;
      s_nd[tid * PER + k] = nd;
      s_area[tid * PER + k] = 0ull;
      s_spike[tid * PER + k] = 0u;
      s_farea[tid * PER + k] = 0ull;
      s_fspike[tid * PER + k] = 0u;
    }
    int2 off = block_exclusive_scan2<kHitBlock>(mine, s_warp);
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      if (cnt[k] > 0) {
        off.x += cnt[k];
        const bool spk = all_special || !(nhs[k] > det.coinLevel);
        s_rec[off.y++] = make_uint4(uint32_t(off.x), uint32_t(tid * PER + k) | (spk ? 0x80000000u : 0u), uint32_t(nds[k]),
                                    uint32_t(nhs[k] - nds[k]));
      }
    }
    if (tid == kHitBlock - 1) {
      s_tot[0] = off.x;
      s_tot[1] = off.y;
      s_rec[off.y] = make_uint4(0x7fffffffu, 0u, 0u, 0u);
    }
    __syncthreads();
    const int H = s_tot[0];  // slots in this group
    const int h0 = int((long long)H * tid / kHitBlock);
    const int h1 = int((long long)H * (tid + 1) / kHitBlock);
    const int maxlen = __reduce_max_sync(0xffffffffu, h1 - h0);
    int c = 0, e = 0, e_end = 0, nd = 0, ns = 0, j = 0;
    bool sp = false;
    int h = h0;  // running slot index; this thread owns [h0, h1)
    if (h1 > h0) {
      // the event owning slot h0: the first entry whose slots end beyond h0
      int lo = -1, hi = s_tot[1] - 1;  // end[lo] <= h0 < end[hi]
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (int(s_rec[mid].x) <= h0)
          lo = mid;
        else
          hi = mid;
      }
      c = hi;
      const uint4 rec = s_rec[c];
      e = int(rec.y & 0x7fffffffu);
      sp = (rec.y >> 31) != 0u;
      e_end = int(rec.x);
      nd = int(rec.z);
      ns = int(rec.w);
      j = h0 - (c > 0 ? int(s_rec[c - 1].x) : 0);
    }
    PhiloxEvent pev;  // the per-event part of the slot blocks' first Philox round
    {
      const uint64_t hev = ev0 + e;
      pev = philox_event_prep(uint32_t(hev), uint32_t(hev >> 32), STREAM_HIT, rk);
    }
    unsigned long long a_fx = 0ull;
    unsigned int a_spike = 0u;
    // An event a thread crosses INTO and later out of lies wholly inside its slot range: nobody else's main loop adds
    // to it, so its sums are stored, not added.  Only the first event of the range (it may have begun in the previous
    // thread's) and the last (flushed after the loop) go through the shared-memory atomics.  A lane crosses every
    // ~27 slots, a warp in 70 % of its iterations, at 1-2 active lanes: the crossing is one record load, two stores
    // and the per-event part of the Philox block (it was the atomics with their carry, a walk over the events
    // without slots and four dependent loads: 12 % of the kernel's warp instructions).
    bool first = true;
    int qcount = 0;  // warp-uniform queue fill
    for (int it = maxlen; it > 0; --it) {
      bool queue = false;
      uint32_t qflags = 0u;  // what the deferred finisher needs of the slot's event (ent_flags)
      if (h < h1) {
        if (h == e_end) {  // crossed into the next event with slots: flush, then advance
          const unsigned long long a_ev = a_fx - (unsigned long long)a_spike * NB_AREA_MAGIC_BITS;
          if (first) {
            smem_add64(&s_area[e], a_ev);
            atomicAdd(&s_spike[e], a_spike);
            first = false;
          } else {
            s_area[e] = a_ev;
            s_spike[e] = a_spike;
          }
          a_fx = 0ull;
          a_spike = 0u;
          ++c;
          const uint4 rec = s_rec[c];
          e = int(rec.y & 0x7fffffffu);
          sp = (rec.y >> 31) != 0u;
          e_end = int(rec.x);
          nd = int(rec.z);
          ns = int(rec.w);
          j = 0;
          const uint64_t hev = ev0 + e;
          pev = philox_event_prep(uint32_t(hev), uint32_t(hev >> 32), STREAM_HIT, rk);
        }
        // one Philox block per slot, 42 bits per photo-electron: 27 (abscissa) + 11 (layer) for the
        // ziggurat normal, 4 for the truncation pre-test
        const u128 r = philox4x32_10_rk_event(pev, uint32_t(j), rk);
        uint32_t A0, A1, A2, L0, L1, L2;
        pe_fields(r, 0, A0, L0);
        pe_fields(r, 1, A1, L1);
        pe_fields(r, 2, A2, L2);
        double z0, z1, z2;
        const bool f0 = zig_fast(A0, L0, s_zk, s_zw, z0), f1 = zig_fast(A1, L1, s_zk, s_zw, z1),
                   f2 = zig_fast(A2, L2, s_zk, s_zw, z2);
        const double v0 = fma(P.pe_sig, z0, P.pe_mu), v1 = fma(P.pe_sig, z1, P.pe_mu), v2 = fma(P.pe_sig, z2, P.pe_mu);
        const bool dbl = j < nd;
        const int nv = slot_valid(j, nd, ns);
        // truncation pre-test (nb_params.h): above pe_cut nothing to do; below it a 4-bit uniform
        // settles it unless it is 15 or the acceptance is under 15/16 (S below pe_sfast)
        const bool c0 = v0 <= P.pe_cut && (v0 < P.pe_sfast || (A0 & NB_PE_PRETEST_MASK) == NB_PE_PRETEST_MASK);
        const bool c1 = v1 <= P.pe_cut && (v1 < P.pe_sfast || (A1 & NB_PE_PRETEST_MASK) == NB_PE_PRETEST_MASK);
        const bool c2 = v2 <= P.pe_cut && (v2 < P.pe_sfast || (A2 & NB_PE_PRETEST_MASK) == NB_PE_PRETEST_MASK);
        queue = sp || !f0 || c0 || (nv >= 2 && (!f1 || c1)) || (nv >= 3 && (!f2 || c2));
        qflags = ent_flags(nv, dbl, sp);
        if (!queue) {
          // areas are summed as integers in units of 2^-32 phe: fma(a, 2^32, 1.5 * 2^52) rounds a 2^32 to the
          // nearest integer (ties to even, as __double2ll_rn) and leaves it in the low mantissa bits; the raw
          // words are added up and the offsets taken out at the flush (a_spike of them)
          const double a1 = dbl ? v0 + v1 : v0;
          const double a2 = dbl ? (nv >= 3 ? v2 : -DBL_MAX) : (nv >= 2 ? v1 : -DBL_MAX);
          const double a3 = (!dbl && nv >= 3) ? v2 : -DBL_MAX;
          if (a1 > det.sPEthr) {
            a_spike += 1u;
            a_fx += (unsigned long long)__double_as_longlong(fma(a1, NB_AREA_FX, NB_AREA_MAGIC));
          }
          if (a2 > det.sPEthr) {
            a_spike += 1u;
            a_fx += (unsigned long long)__double_as_longlong(fma(a2, NB_AREA_FX, NB_AREA_MAGIC));
          }
          if (a3 > det.sPEthr) {
            a_spike += 1u;
            a_fx += (unsigned long long)__double_as_longlong(fma(a3, NB_AREA_FX, NB_AREA_MAGIC));
          }
        }
      }
      const unsigned m = __ballot_sync(0xffffffffu, queue);
      if (m) {
        if (queue) {
          const uint32_t at = q_addr + uint32_t(qcount + __popc(m & ((1u << lane) - 1u))) * uint32_t(sizeof(SlotEntry));
          asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(at), "r"(uint32_t(j)), "r"(uint32_t(e) | qflags) : "memory");
        }
        qcount += __popc(m);
        __syncwarp();
        if (qcount >= 32) {
          qcount -= 32;
          // hand the 32 entries at [qcount, qcount + 32) to the deferred finisher
          if (!queue_push(Q, all_special, &s_q[wid][qcount], 32, lane, uint32_t(gbase)))
            hits_finish_slow(P, s_nh, s_nd, s_farea, s_fspike, ev0, &s_q[wid][qcount], 32, lane, s_zk, s_zw);
        }
      }
      ++j;
      ++h;
    }
    if (qcount > 0) {  // the group's leftovers
      if (!queue_push(Q, all_special, &s_q[wid][0], qcount, lane, uint32_t(gbase)))
        hits_finish_slow(P, s_nh, s_nd, s_farea, s_fspike, ev0, &s_q[wid][0], qcount, lane, s_zk, s_zw);
    }
    if (h1 > h0) {
      smem_add64(&s_area[e], a_fx - (unsigned long long)a_spike * NB_AREA_MAGIC_BITS);
      atomicAdd(&s_spike[e], a_spike);
    }
    __syncthreads();
#pragma unroll
    for (int jj = 0; jj < PER; ++jj) {
      const int i = jj * kHitBlock + tid;  // coalesced
      const uint64_t idx = gbase + i;
      if (idx < n_valid) {
        reinterpret_cast<unsigned long long*>(W.area)[idx] = s_area[i] + s_farea[i];  // 2^-32 phe; the finisher adds to it
        W.spike[idx] = int(s_spike[i] + s_fspike[i]);
        W.nphe[idx] = s_nh[i] + s_nd[i];
      }
    }
    __syncthreads();
  }
}


__global__ void __launch_bounds__(kHitBlock, NB_HIT_MINB)
    k_hits(const __grid_constant__ RunParams P, unsigned int* __restrict__ next_group, const HitQueue Q) {
  hits_body<false>(&P, nullptr, nullptr, next_group, Q);
}

__global__ void __launch_bounds__(kHitBlock, NB_HIT_MINB)
    k_hits_sweep(const __grid_constant__ SweepArgs A, unsigned int* __restrict__ next_group, const HitQueue Q) {
  __shared__ __align__(16) RunParams sP;
  hits_body<true>(nullptr, &A, &sP, next_group, Q);
}

// one event of k_post.  s_band / s_h2: the block's shared-memory band accumulator (band_smem) ; gband: the global
// accumulator the block flushes into (used directly when the band does not fit in shared memory)
NB_D void post_body(const uint32_t* rk, const RunParams& P, const Work& W, uint64_t idx, uint64_t gidx, uint64_t ev, bool active,
                    unsigned long long* s_band, unsigned int* s_h2, unsigned long long* gband, bool band_smem,
                    int* __restrict__ err_flag) {
  const nb200_detector& det = P.det;
  const int nb = P.ana.numBins, lb = P.ana.logBins;
  const bool cli = !P.vec_mode;
  Stream g;
  g.init(rk, ev, STREAM_POST);
  double keV = W.keV[idx];
  const double px = W.px[idx], py = W.py[idx], pz = W.pz[idx], sx = W.sx[idx], sy = W.sy[idx];
  const double field = W.field[idx], dt = W.dt[idx];
  Yields y{W.yNph[idx], W.yNe[idx], 0., W.yL[idx], field, -999.};
  S1Pre pre;
  pre.posDepSm = W.posDepSm[idx];
  pre.fitSm = W.fitSm[idx];
  const int nh = W.nHits[idx];
  pre.full = nh >= 0;
  pre.nHits = pre.full ? nh : -(nh + 1);
  pre.eff = 0.;
  S1Acc acc{0., 0, 0};
  if (pre.full) {
    if (nh > 0) {
      acc.area = double(reinterpret_cast<const unsigned long long*>(W.area)[idx]) * (1. / NB_AREA_FX);
      acc.spike = W.spike[idx];
      acc.nphe = W.nphe[idx];
    }
  } else
    s1_parametric(g, ev, P, pre.nHits, acc);
  double scint[9], scint2[9];
  const PostDraws D = post_draws(g);
  NB_LOCKSTEP();
  s1_epilogue(g, P, pre, acc, D, scint);
  int Ne = W.electrons[idx];
  if (cli && pz < det.cathode) Ne = 0;  // execNEST.cpp:1107
  NB_LOCKSTEP();
  get_s2(g, ev, P, D, Ne, px, py, pz, sx, sy, dt, field, scint2);
  NB_LOCKSTEP();
  double signal1 = 0., signal2 = 0., signalE = 0., keVee = 0.;
  const double keV_true = keV;
  if (cli) {
    select_signals(P.ana, scint, scint2, signal1, signal2);
    signalE = reconstruct_energy(P, y, scint, scint2, signal1, signal2, field, keV, keVee);
  }
  NB_LOCKSTEP();
  if (!active) return;
  if (cli) {
    if (P.do_band) {
      bool acc_ok = false;
      double xv = 0., yv = 0.;
      int bin = band_bin(P, signal1, signal2, acc_ok, xv, yv);
      if (bin >= 0) {
        unsigned long long* B = band_smem ? s_band : gband;
        unsigned long long* b = B + size_t(bin) * BAND_WORDS;
        if (acc_ok) {
          atomicAdd(&b[0], 1ull);
          atomicAdd(&b[2], (unsigned long long)(long long)llrint(xv * P.band_fx_x));
          atomicAdd(&b[3], (unsigned long long)(long long)llrint(yv * NB200_FX_Y));
          atomicAdd(&b[4], (unsigned long long)(long long)llrint(yv * yv * NB200_FX_Y2));
          atomicAdd(&b[5], (unsigned long long)(long long)llrint(yv * yv * yv * NB200_FX_Y3));
          if (yv != 0. && lb > 0) {
            int jb = int((yv - P.ana.logMin) / (P.ana.logMax - P.ana.logMin) * double(lb));
            jb = min(max(jb, 0), lb - 1);
            if (band_smem)
              atomicAdd(&s_h2[bin * lb + jb], 1u);
            else
              atomicAdd(&gband[size_t(nb) * BAND_WORDS + size_t(bin) * lb + jb], 1ull);
          }
        } else
          atomicAdd(&b[1], 1ull);
      }
    }
  }
  if (P.write_soa) {
    const SoaDev& o = P.soa;
    const bool f32 = P.write_soa == 2;
    if (o.energy_kev) soa_put(o.energy_kev, gidx, keV_true, f32);
    if (o.x_mm) soa_put(o.x_mm, gidx, px, f32);
    if (o.y_mm) soa_put(o.y_mm, gidx, py, f32);
    if (o.z_mm) soa_put(o.z_mm, gidx, pz, f32);
    if (o.field) soa_put(o.field, gidx, field, f32);
    if (o.driftTime) soa_put(o.driftTime, gidx, dt, f32);
    if (o.smear_x) soa_put(o.smear_x, gidx, sx, f32);
    if (o.smear_y) soa_put(o.smear_y, gidx, sy, f32);
    if (o.photons) o.photons[gidx] = W.photons[idx];
    if (o.electrons) o.electrons[gidx] = Ne;
    if (o.ions) o.ions[gidx] = W.ions[idx];
    if (o.excitons) o.excitons[gidx] = W.excitons[idx];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      if (o.s1[k]) soa_put(o.s1[k], gidx, scint[k], f32);
      if (o.s2[k]) soa_put(o.s2[k], gidx, scint2[k], f32);
    }
    if (o.signal1) soa_put(o.signal1, gidx, signal1, f32);
    if (o.signal2) soa_put(o.signal2, gidx, signal2, f32);
    if (o.signalE) soa_put(o.signalE, gidx, cli ? signalE : keV, f32);
    if (o.Nph_mean) soa_put(o.Nph_mean, gidx, y.Nph, f32);
    if (o.Ne_mean) soa_put(o.Ne_mean, gidx, y.Ne, f32);
    if (o.keVee) soa_put(o.keVee, gidx, keVee, f32);
  }
}

NB_D void band_smem_zero(unsigned long long* s_band, unsigned int* s_h2, int nb, int lb) {
  for (int i = threadIdx.x; i < nb * BAND_WORDS; i += kPostBlock) s_band[i] = 0ull;
  for (int i = threadIdx.x; i < nb * lb; i += kPostBlock) s_h2[i] = 0u;
  __syncthreads();
}
NB_D void band_smem_flush(const unsigned long long* s_band, const unsigned int* s_h2, int nb, int lb,
                          unsigned long long* gband) {
  __syncthreads();
  for (int i = threadIdx.x; i < nb * BAND_WORDS; i += kPostBlock) {
    unsigned long long v = s_band[i];
    if (v) atomicAdd(&gband[i], v);
  }
  for (int i = threadIdx.x; i < nb * lb; i += kPostBlock) {
    unsigned int v = s_h2[i];
    if (v) atomicAdd(&gband[size_t(nb) * BAND_WORDS + i], (unsigned long long)v);
  }
}

__global__ void __launch_bounds__(kPostBlock, NB_POST_MINB)
    k_post(const __grid_constant__ RunParams P, int* __restrict__ err_flag) {
  extern __shared__ __align__(16) unsigned long long s_band[];  // [numBins*6] then u32 hist2d
  const int tid = threadIdx.x;
  const int nb = P.ana.numBins, lb = P.ana.logBins;
  const bool band_smem = P.do_band && P.band_in_smem;
  unsigned int* s_h2 = reinterpret_cast<unsigned int*>(s_band + size_t(nb) * BAND_WORDS);
  if (band_smem) band_smem_zero(s_band, s_h2, nb, lb);
  for (uint64_t base = blockIdx.x * uint64_t(kPostBlock); base < P.n_events;
       base += uint64_t(gridDim.x) * kPostBlock) {
    const bool active = base + tid < P.n_events;
    const uint64_t idx = active ? base + tid : P.n_events - 1;
    const uint64_t gidx = P.chunk_off + idx;
    post_body(P.rk, P, P.work, idx, gidx, P.first_event + gidx, active, s_band, s_h2, P.band, band_smem, err_flag);
  }
  if (band_smem) band_smem_flush(s_band, s_h2, nb, lb, P.band);
}

// sweep: a block takes a contiguous run of 1024-slot tiles (few point changes) and flushes its shared band into
// the point's slab whenever the point changes
__global__ void __launch_bounds__(kPostBlock, NB_POST_MINB)
    k_post_sweep(const __grid_constant__ SweepArgs A, int band_in_smem, int* __restrict__ err_flag) {
  extern __shared__ __align__(16) unsigned long long s_band[];
  __shared__ __align__(16) RunParams sP;
  const int tid = threadIdx.x;
  const uint64_t n_tiles = (A.n_slots + kPostBlock - 1) / kPostBlock;
  const uint64_t t0 = n_tiles * blockIdx.x / gridDim.x, t1 = n_tiles * (blockIdx.x + 1) / gridDim.x;
  uint32_t cur = 0xffffffffu;
  int nb = 0, lb = 0;
  unsigned int* s_h2 = nullptr;
  unsigned long long* gband = nullptr;
  bool band_smem = false;
  for (uint64_t t = t0; t < t1; ++t) {
    const uint64_t base = t * kPostBlock;
    const uint64_t slot = A.slot0 + base;
    const uint32_t point = uint32_t(slot / A.n_pad);
    if (point != cur) {
      if (cur != 0xffffffffu && band_smem) band_smem_flush(s_band, s_h2, nb, lb, gband);
      sweep_load_params(&sP, A.params, point);
      cur = point;
      nb = sP.ana.numBins;
      lb = sP.ana.logBins;
      s_h2 = reinterpret_cast<unsigned int*>(s_band + size_t(nb) * BAND_WORDS);
      gband = A.bands + uint64_t(point) * A.band_words;
      band_smem = sP.do_band && band_in_smem;
      if (band_smem) band_smem_zero(s_band, s_h2, nb, lb);
    }
    const uint64_t it0 = slot - uint64_t(point) * A.n_pad;  // event index of the tile's first slot (< n_per_point)
    const uint64_t i = it0 + tid;
    const bool active = i < A.n_per_point;
    const uint64_t ie = active ? i : A.n_per_point - 1;
    // padding lanes redo the point's last event (its workspace entry is in this tile) and store nothing
    post_body(A.rk, sP, A.work, base + (ie - it0), ie, A.first_event + ie, active, s_band, s_h2, gband, band_smem, err_flag);
  }
  if (cur != 0xffffffffu && band_smem) band_smem_flush(s_band, s_h2, nb, lb, gband);
}

// CalculateG2's single-electron Monte Carlo (NEST.cpp:2169-2215): one thread = one electron
__global__ void k_se(const __grid_constant__ RunParams P, int n, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const nb200_detector& d = P.det;
  const double elYield = P.rc.g2_params[0];
  Stream g;
  g.init(P.rk, uint64_t(i), STREAM_SE);
  long long Nph = (long long)floor(gauss(g, elYield, sqrt(d.s2Fano * elYield), true) + 0.5);
  double sn, cs;
  sincospi(2. * g.uniform(), &sn, &cs);
  const double r = d.radius * sqrt(g.uniform());
  const double posDep = fit_s2(d.FitS2, r * cs, r * sn) / P.s2_center;
  long long nHits = binomial(P.rk, uint64_t(i), STREAM_BIN_SE_HIT, Nph, d.g1_gas * posDep);
  double Nphe = double(nHits + binomial(P.rk, uint64_t(i), STREAM_BIN_SE_DPE, nHits, d.P_dphe));
  double area = gauss(g, Nphe, d.sPEres * sqrt(Nphe), true);
  double sig;
  if (d.noiseQuadratic[1] != 0) {
    double a = d.noiseQuadratic[1] * (area * area), b = d.noiseLinear[1] * area;
    sig = sqrt(a * a + b * b);
  } else
    sig = d.noiseLinear[1] * area;
  area = gauss(g, area, sig, true);
  if (d.s2_thr < 0.) {
    const double t = fit_tba(d.FitTBA, 0., 0., d.TopDrift / 2.);
    area = gauss(g, t * area, sqrt(t * area * (1. - t)), true);
  }
  out[i] = (area / posDep) / (1. + d.P_dphe);
}

// test hook: the hit loop's specialised math on caller-supplied bits; out = {-ln u, sin, cos}
__global__ void k_hitmath(size_t n, const uint32_t* __restrict__ bits, double* __restrict__ out) {
  __shared__ __align__(16) double s_lntab[64];
  if (threadIdx.x < 64) s_lntab[threadIdx.x] = c_lntab[threadIdx.x];
  __syncthreads();
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
    double s, c;
    sincos_bits(bits[3 * i + 2], s, c);
    out[3 * i] = neg_log_bits(bits[3 * i], bits[3 * i + 1], s_lntab);
    out[3 * i + 1] = s;
    out[3 * i + 2] = c;
  }
}

// test hook (nb200_debug_normal): the hit loop's ziggurat normals exactly as k_hits and its finisher draw them
// (pe_normal, the function both go through): the three photo-electrons of slot i % 1024 of event
// first + i / 1024; kind[.] = 0 layer core (finished in the main loop), 1 wedge / tail completion
__global__ void __launch_bounds__(256) k_zignormal(uint64_t seed, uint64_t first, size_t slots,
                                                   double* __restrict__ out, unsigned char* __restrict__ kind) {
  __shared__ double s_zw[NB_ZIG_N];
  __shared__ uint32_t s_zk[NB_ZIG_N];
  for (int i = threadIdx.x; i < NB_ZIG_N; i += blockDim.x) {
    s_zw[i] = g_zig_w[i];
    s_zk[i] = g_zig_k[i];
  }
  __syncthreads();
  uint32_t rk[20];
  philox_round_keys(seed, rk);
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < slots; i += size_t(gridDim.x) * blockDim.x) {
    const uint64_t ev = first + (i >> 10);
    const uint32_t j = uint32_t(i & 1023u), e0 = uint32_t(ev), e1 = uint32_t(ev >> 32);
    const u128 r = philox4x32_10_rk(e0, e1, j, STREAM_HIT, rk);
    const u128 t = philox4x32_10_rk(e0, e1, j, STREAM_HIT2, rk);
#pragma unroll
    for (int which = 0; which < 3; ++which) {
      uint32_t A;
      bool fast;
      out[3 * i + which] = pe_normal(r, t, which, e0, e1, j, uint32_t(seed), uint32_t(seed >> 32), s_zk, s_zw, A, fast);
      if (kind) kind[3 * i + which] = fast ? 0 : 1;
    }
  }
}

// k_binomial: the binomial sampler alone (test hook, nb200_debug_binomial)
struct RoundKeys {
  uint32_t k[20];  // philox_round_keys(seed): a kernel parameter, so constant-bank operands
};
__global__ void __launch_bounds__(256) k_binomial(const __grid_constant__ RoundKeys K, uint64_t first, size_t count,
                                                  long long n, double p, long long* __restrict__ out) {
  const uint32_t* const rk = K.k;
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < count; i += size_t(gridDim.x) * blockDim.x)
    out[i] = binomial(rk, first + i, STREAM_BIN_S1, n, p);
}

// test hook (nb200_debug_binomial_list): the in-line sampler with a trial count per draw, on the stream a given caller
// uses -- what k_lar's queued draws (nb_lar.cuh) must reproduce bit for bit
__global__ void __launch_bounds__(256) k_binomial_list(const __grid_constant__ RoundKeys K, uint64_t first, size_t count,
                                                       const long long* __restrict__ n, double p, uint32_t stream,
                                                       long long* __restrict__ out) {
  const uint32_t* const rk = K.k;
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < count; i += size_t(gridDim.x) * blockDim.x)
    out[i] = binomial(rk, first + i, stream, n[i], p);
}

// test hook (nb200_debug_binomial_fixed): the alias-table sampler of k_pre's double-phe count
__global__ void __launch_bounds__(256) k_binomial_fixed(const __grid_constant__ RoundKeys K, uint64_t first,
                                                        size_t count, int n, const uint2* __restrict__ tab,
                                                        long long* __restrict__ out) {
  const uint32_t* const rk = K.k;
  for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < count; i += size_t(gridDim.x) * blockDim.x)
    out[i] = binomial_fixed_p(rk, first + i, STREAM_BIN_S1DPE, n, tab);
}

// -------------------------------------------------------------------------------------------
// k_quanta: NESTcalc::GetQuanta (NEST.cpp:248-432) for n independent yield records -- the stage call of the
// reference's per-event API (FullCalculation NEST.cpp:17-40, bareNEST.cpp:186), one thread per record.
// y: 6 columns of n as k_yields writes them; qi: 4 int columns (photons, electrons, ions, excitons); qd: 2
// double columns (recombProb, Variance).  Record i draws from the Philox streams of event first_event + i.
struct QuantaParams {
  nb200_detector det;
  nb200_model model;
  double density;
  uint64_t n, first_event;
  uint32_t rk[20];  // Philox round keys of the seed
  const double* y;
  int32_t* qi;
  double* qd;
};

__global__ void __launch_bounds__(256) k_quanta(const __grid_constant__ QuantaParams P, int* err_flag) {
  for (uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; i < P.n;
       i += uint64_t(gridDim.x) * blockDim.x) {
    Yields y;
    y.Nph = P.y[i];
    y.Ne = P.y[P.n + i];
    y.NexONi = P.y[2 * P.n + i];
    y.L = P.y[3 * P.n + i];
    y.field = P.y[4 * P.n + i];
    y.dT = P.y[5 * P.n + i];
    const uint64_t ev = P.first_event + i;
    Stream g;
    g.init(P.rk, ev, STREAM_MAIN);
    int err = 0;
    const Quanta q = get_quanta(g, ev, y, P.density, P.det, P.model, &err);
    if (err) atomicOr(err_flag, err);
    P.qi[i] = q.photons;
    P.qi[P.n + i] = q.electrons;
    P.qi[2 * P.n + i] = q.ions;
    P.qi[3 * P.n + i] = q.excitons;
    P.qd[i] = q.recombProb;
    P.qd[P.n + i] = q.Variance;
  }
}

// k_s2: NESTcalc::GetS2, S2CalculationMode::Full (NEST.cpp:1724-1776, 1975-2063) for n independent records -- the
// stage call of the reference's per-event API (bareNEST.cpp:210), one thread per record, the same device function
// k_post calls.  pos: 5 columns of n (truth x, y, z, smeared x, y); out: 9 columns of n (ionization[0..8]).
struct S2Args {
  uint64_t n;
  const int32_t* Ne;
  const double *pos, *dt, *field;
  double* out;
};

__global__ void __launch_bounds__(256) k_s2(const __grid_constant__ RunParams P, const S2Args a) {
  for (uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; i < a.n; i += uint64_t(gridDim.x) * blockDim.x) {
    const uint64_t ev = P.first_event + i;
    Stream g;
    g.init(P.rk, ev, STREAM_POST);
    double o[9];
    const PostDraws D = post_draws(g);
    get_s2(g, ev, P, D, a.Ne[i], a.pos[i], a.pos[a.n + i], a.pos[2 * a.n + i], a.pos[3 * a.n + i], a.pos[4 * a.n + i], a.dt[i],
           a.field[i], o);
#pragma unroll
    for (int k = 0; k < 9; ++k) a.out[k * a.n + i] = o[k];
  }
}

// k_s1pre / k_s1post: NESTcalc::GetS1 (NEST.cpp:1288-1722; bareNEST.cpp:203) as a stage call on caller-given photons
// and positions.  The hit loop itself is k_hits, unchanged: k_s1pre runs the prelude (binomial hit count, double-phe
// count) into the stage workspace exactly as k_pre does, k_s1post the epilogue exactly as k_post does.
// Nph: photons per record; pos: 6 columns (truth x, y, z, smeared x, y, z) and energy, with column stride `stride` and
// this chunk starting at `off`; out: 9 columns (scintillation[0..8]).
struct S1Args {
  uint64_t stride, off;
  const int32_t* Nph;
  const double *pos, *energy;
  double* out;
};

__global__ void __launch_bounds__(256) k_s1pre(const __grid_constant__ RunParams P, const S1Args a) {
  const Work& W = P.work;
  for (uint64_t idx = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; idx < P.n_events;
       idx += uint64_t(gridDim.x) * blockDim.x) {
    const uint64_t gi = a.off + idx, ev = P.first_event + P.chunk_off + idx;
    const S1Pre pre = s1_prelude(P.rk, ev, P, a.Nph[gi], a.pos[gi], a.pos[a.stride + gi], a.pos[2 * a.stride + gi],
                                 a.pos[3 * a.stride + gi], a.pos[4 * a.stride + gi], a.pos[5 * a.stride + gi], a.energy[gi]);
    W.posDepSm[idx] = pre.posDepSm;
    W.fitSm[idx] = pre.fitSm;
    W.nHits[idx] = pre.full ? pre.nHits : -(pre.nHits + 1);  // negative: Parametric S1 in k_s1post
    W.nphe[idx] = pre.nDP;
  }
}

__global__ void __launch_bounds__(256) k_s1post(const __grid_constant__ RunParams P, const S1Args a) {
  const Work& W = P.work;
  for (uint64_t idx = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; idx < P.n_events;
       idx += uint64_t(gridDim.x) * blockDim.x) {
    const uint64_t gi = a.off + idx, ev = P.first_event + P.chunk_off + idx;
    Stream g;
    g.init(P.rk, ev, STREAM_POST);
    S1Pre pre;
    pre.posDepSm = W.posDepSm[idx];
    pre.fitSm = W.fitSm[idx];
    const int nh = W.nHits[idx];
    pre.full = nh >= 0;
    pre.nHits = pre.full ? nh : -(nh + 1);
    pre.eff = 0.;
    pre.nDP = 0;
    S1Acc acc{0., 0, 0};
    if (pre.full) {
      if (nh > 0) {
        acc.area = double(reinterpret_cast<const unsigned long long*>(W.area)[idx]) * (1. / NB_AREA_FX);
        acc.spike = W.spike[idx];
        acc.nphe = W.nphe[idx];
      }
    } else
      s1_parametric(g, ev, P, pre.nHits, acc);
    double scint[9];
    const PostDraws D = post_draws(g);
    s1_epilogue(g, P, pre, acc, D, scint);
#pragma unroll
    for (int k = 0; k < 9; ++k) a.out[k * a.stride + gi] = scint[k];
  }
}

// ---- band post-processing for batched bands, on the device --------------------------------------------
// A sweep leaves hundreds of band accumulators in device memory; what a fit over the grid wants back from each is a
// handful of numbers, not 28 KB: GetBand's statistics (execNEST.cpp:1582-1618), rootNEST's goodness of fit against a
// target band (examples/rootNEST.cpp:600-643) and the leakage below a centroid line (:880-915).  One block per grid
// point, one thread per S1 bin for the statistics, a block reduction for the sums -- the same formulas, in the same
// order per bin, as nb200_band_finalize / nb200_band_chi2 / nb200_band_leakage on the host.
struct BandPostArgs {
  const unsigned long long* bands;  // [n_points][band_words]
  uint64_t band_words;
  int32_t numBins, logBins, useS2, free_param;
  double border, binWidth, fx_x, logMin, logMax;
  const double* target;    // [numBins][7] band to compare with, or null
  const double* centroid;  // [numBins] leakage line, or null
  double* stats;           // [n_points][numBins][7] or null
  double* gof;             // [n_points][6] or null (needs target)
  double* leak;            // [n_points][3] = events below, all events, ratio; or null (needs centroid)
};
__global__ void __launch_bounds__(256) k_band_post(const __grid_constant__ BandPostArgs A) {
  __shared__ double s_red[6][256];
  const unsigned long long* B = A.bands + uint64_t(blockIdx.x) * A.band_words;
  double acc[6] = {0., 0., 0., 0., 0., 0.};  // chi2 mean, chi2 width, pct mean, pct width, leak below, leak all
  for (int j = threadIdx.x; j < A.numBins; j += blockDim.x) {
    const unsigned long long* b = B + size_t(j) * BAND_WORDS;
    const double N = double(b[0]);
    const double Sx = double((long long)b[2]) / A.fx_x, Sy = double((long long)b[3]) / NB200_FX_Y,
                 Sy2 = double((long long)b[4]) / NB200_FX_Y2, Sy3 = double((long long)b[5]) / NB200_FX_Y3;
    double st[7];
    st[0] = A.border + A.binWidth / 2. + double(j) * A.binWidth;
    st[1] = Sx / N;
    const double mu = Sy / N;
    st[2] = mu;
    const double var = (Sy2 - 2. * mu * Sy + N * mu * mu) / (N - 1.);
    st[3] = sqrt(var);
    st[4] = st[3] / sqrt(N);
    st[5] = N / (N + double(b[1]));
    const double m3 = Sy3 - 3. * mu * Sy2 + 3. * mu * mu * Sy - N * mu * mu * mu;
    st[6] = m3 / (st[3] * st[3] * st[3]) / (N - 2.);
    if (A.stats)
      for (int k = 0; k < 7; ++k) A.stats[(uint64_t(blockIdx.x) * A.numBins + j) * 7 + k] = st[k];
    if (A.target) {
      const double* y = A.target + 7 * size_t(j);
      const bool ok = isfinite(st[2]) && isfinite(y[2]) && isfinite(st[3]) && isfinite(y[3]) && st[4] > 0. && y[4] > 0.;
      if (ok && ((A.useS2 == 0 && fabs(st[2]) < 5.) || A.useS2 != 0)) {
        double error = sqrt(st[4] * st[4] + y[4] * y[4]);
        acc[0] += (y[2] - st[2]) / error * ((y[2] - st[2]) / error);
        error = sqrt(0.5 * (st[4] * st[4] + y[4] * y[4]));
        acc[1] += (y[3] - st[3]) / error * ((y[3] - st[3]) / error);
        acc[2] += 100. * (y[2] - st[2]) / y[2];
        acc[3] += 100. * (y[3] - st[3]) / y[3];
      }
    }
    if (A.centroid && A.logBins > 0) {
      const unsigned long long* row = B + size_t(A.numBins) * BAND_WORDS + size_t(j) * A.logBins;
      const double w = (A.logMax - A.logMin) / double(A.logBins);
      const double pos = (A.centroid[j] - A.logMin) / w;
      double in_hist = 0., below = 0.;
      for (int k = 0; k < A.logBins; ++k) {
        const double c = double(row[k]);
        in_hist += c;
        if (double(k + 1) <= pos)
          below += c;
        else if (double(k) < pos)
          below += c * (pos - double(k));
      }
      if (A.centroid[j] > 0.) below += N - in_hist;
      if (N > 0.) {
        acc[4] += below;
        acc[5] += N;
      }
    }
  }
  for (int k = 0; k < 6; ++k) s_red[k][threadIdx.x] = acc[k];
  __syncthreads();
  if (threadIdx.x == 0) {
    // summed in bin order by one thread: the same order of additions as the host functions
    double t[6] = {0., 0., 0., 0., 0., 0.};
    // thread i holds bins i, i + 256, ...: with numBins <= 256 (every shipped analysis) thread order = bin order
    for (int i = 0; i < int(blockDim.x); ++i)
      for (int k = 0; k < 6; ++k) t[k] += s_red[k][i];
    if (A.gof) {
      const int DoF = A.numBins - abs(A.free_param);
      double* o = A.gof + 6 * uint64_t(blockIdx.x);
      const double c0 = t[0] / double(DoF - 1), c1 = t[1] / double(DoF - 1);
      o[0] = c0;
      o[1] = c1;
      o[2] = 0.5 * (c0 + c1);
      o[3] = sqrt(c0 * c1);
      o[4] = -(t[2] / double(A.numBins));
      o[5] = -(t[3] / double(A.numBins));
    }
    if (A.leak) {
      double* o = A.leak + 3 * uint64_t(blockIdx.x);
      o[0] = t[4];
      o[1] = t[5];
      o[2] = t[4] / t[5];
    }
  }
}

// k_band_fit: the per-bin fits of the band post-processing for batched bands -- GetBand_Gaussian (examples/rootNEST.cpp:
// 1309-1359) or the skew-Gaussian band (:1360-1507) of every S1 slice of every grid point, and the leakage below a line
// read off the fits (:700-738, 848 Gaussian; :700-735, 1587-1602 skew-normal CDF with the reference's Owen's-T
// quadrature).  One WARP per (grid point, S1 bin): the lanes share out the slice's log bins (the sums of the chi^2,
// J^T W r and J^T W J are 15 butterfly all-reduces per evaluation) and walk the minimiser in step -- a sweep has hundreds
// of points x ~100 bins of such slices.  The statements are nb_fit.h's, the very functions the host entry points
// (nb200_band_fit_gauss, nb200_hist_fit_cov, nb200_band_leakage_gauss / _skew) run with a group of one.  (One THREAD per
// slice, the first version, measured 7.7 ms for 24 500 slices: 1.3 warps per scheduler, each a serial chain of exp.)
struct BandFitArgs {
  const unsigned long long* bands;  // [n_points][band_words]
  uint64_t band_words, n_rows;      // n_rows = n_points * numBins
  int32_t numBins, logBins, model, cols;  // model 0: 8 columns, model 1: 12
  double logMin, logMax;
  const double* line;  // [numBins] leakage line, or null
  double* fit;         // [n_points][numBins][cols]
  double* leak;        // [n_points][numBins] or null (needs line)
};
constexpr int kBandFitRows = 4;  // slices (warps) per block
__global__ void __launch_bounds__(kBandFitRows * 32) k_band_fit(const __grid_constant__ BandFitArgs A) {
  const uint64_t t = blockIdx.x * uint64_t(kBandFitRows) + (threadIdx.x >> 5);  // the warp's slice
  if (t >= A.n_rows) return;
  const uint64_t point = t / uint64_t(A.numBins);
  const int j = int(t - point * uint64_t(A.numBins));
  const unsigned long long* B = A.bands + point * A.band_words;
  const unsigned long long* b = B + size_t(j) * BAND_WORDS;
  const unsigned long long* row = B + size_t(A.numBins) * BAND_WORDS + size_t(j) * A.logBins;
  const double w = (A.logMax - A.logMin) / double(A.logBins);
  const double N = double(b[0]);
  double o[12];
  double leak = 0.;
  if (A.model == 0) {
    fit_gauss_row<FitWarp>(row, A.logBins, A.logMin, w, N, o);
    if (A.line && o[7] > 0.) leak = (1. - erf((o[1] - A.line[j]) / o[2] / sqrt(2.))) / 2.;
  } else {
    // GetBand's moments of the slice (execNEST.cpp:1582-1618), as k_band_post
    const double Sy = double((long long)b[3]) / NB200_FX_Y, Sy2 = double((long long)b[4]) / NB200_FX_Y2,
                 Sy3 = double((long long)b[5]) / NB200_FX_Y3;
    const double mu = Sy / N;
    const double sd = sqrt((Sy2 - 2. * mu * Sy + N * mu * mu) / (N - 1.));
    const double m3 = Sy3 - 3. * mu * Sy2 + 3. * mu * mu * Sy - N * mu * mu * mu;
    fit_skew_row<FitWarp>(row, A.logBins, A.logMin, w, mu, sd, m3 / (sd * sd * sd) / N, o);
    if (A.line && o[11] > 0.) {  // (warp-uniform: every lane holds the same o)
      const double c = A.line[j];
      leak = 0.5 + 0.5 * erf((c - o[1]) / o[2] / sqrt(2.)) - 2. * owens_t_ref_by<FitWarp>((c - o[1]) / o[2], o[3]);
    }
  }
  if ((threadIdx.x & 31u) == 0u) {
    for (int k = 0; k < A.cols; ++k) A.fit[t * uint64_t(A.cols) + k] = o[k];
    if (A.leak) A.leak[t] = leak;
  }
}

// ---- the remaining public model methods of NESTcalc as stage calls (NEST.hh:311-333, 345, 404) ---------------------
// k_widths: RecombOmegaNR (NEST.cpp:164-171), RecombOmegaER (:173-225) and FanoER (:227-246) for n records;
// out: 3 columns of n.
struct WidthsParams {
  nb200_model model;
  int32_t inGas, _pad;
  double density;
  uint64_t n;
  const double *efield, *elecFrac, *nq;
  double* out;
};
__global__ void __launch_bounds__(256) k_widths(const __grid_constant__ WidthsParams P) {
  const double* w = P.model.NRERWidthsParam;
  for (uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; i < P.n; i += uint64_t(gridDim.x) * blockDim.x) {
    P.out[i] = recomb_omega_nr(P.elecFrac[i], w);
    P.out[P.n + i] = recomb_omega_er(P.efield[i], P.elecFrac[i], P.nq[i], w, P.model.n_widths, nullptr);
    P.out[2 * P.n + i] = fano_er(P.density, P.nq[i], P.efield[i], w, P.inGas != 0);
  }
}

// k_spike: NESTcalc::GetSpike (NEST.cpp:2243-2268) for n records: pos 3 columns (dx, dy, dz), scint 9 columns
// (scintillation[0..8] of GetS1); out 2 columns (smeared and position-corrected spike count).
struct SpikeArgs {
  uint64_t n;
  const double *pos, *scint;
  double* out;
};
__global__ void __launch_bounds__(256) k_spike(const __grid_constant__ RunParams P, const SpikeArgs a) {
  for (uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; i < a.n; i += uint64_t(gridDim.x) * blockDim.x) {
    const double s4 = a.scint[4 * a.n + i], s5 = a.scint[5 * a.n + i], s6 = a.scint[6 * a.n + i], s7 = a.scint[7 * a.n + i];
    double o0, o1;
    if (s7 > NB_SPIKES_MAXM) {
      o0 = s4;
      o1 = s5;
    } else {
      Stream g;
      g.init(P.rk, P.first_event + i, STREAM_POST);
      o0 = zero_trunc_gauss(g, s6, (P.det.sPEres / 4.) * sqrt(fabs(s6)));
      o1 = o0 / fit_s1(P.det.FitS1, a.pos[i], a.pos[a.n + i], a.pos[2 * a.n + i]) * P.s1_center;
    }
    a.out[i] = o0;
    a.out[a.n + i] = o1;
  }
}

// k_xyres: NESTcalc::xyResolution (NEST.cpp:2699-2736) for n records; out 2 columns (smeared x, y)
struct XyArgs {
  uint64_t n;
  const double *x, *y, *A_top;
  double* out;
};
__global__ void __launch_bounds__(256) k_xyres(const __grid_constant__ RunParams P, const XyArgs a) {
  for (uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; i < a.n; i += uint64_t(gridDim.x) * blockDim.x) {
    Stream g;
    g.init(P.rk, P.first_event + i, STREAM_MAIN);
    double sx, sy;
    xy_resolution(g, P.det, a.x[i], a.y[i], a.A_top[i], sx, sy);
    a.out[i] = sx;
    a.out[a.n + i] = sy;
  }
}

// k_spectrum: the energy samplers alone (TestSpectra::*_spectrum, src/TestSpectra.cpp:31-499; flat / Gauss /
// exponential of execNEST.cpp:769-781): n energies, the same device function and streams as the chain's k_pre
__global__ void __launch_bounds__(256) k_spectrum(const __grid_constant__ RunParams P, double* __restrict__ out) {
  const uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
  const uint64_t n_pad = (P.n_events + 31) / 32 * 32;  // whole warps: the box rejections share candidates across a warp
  for (uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; i < n_pad; i += stride) {
    const bool active = i < P.n_events;
    const uint64_t ie = active ? i : P.n_events - 1;
    Stream g;
    g.init(P.rk, P.first_event + ie, STREAM_MAIN);
    const double keV = sample_energy(g, P, ie);
    if (active) out[i] = keV;
  }
}

// k_lar: LArNEST::FullCalculation (src/LArNEST.cpp:17-33) for species NR/ER/Alpha
// -------------------------------------------------------------------------------------------
struct LArParams {
  int32_t species;
  uint64_t n, first_event, seed;
  uint32_t rk[20];  // Philox round keys of the seed
  double density;
  const double *E, *field;
  double *yields, *fluct;  // [9][n], [4][n]; either may be null
};

template <int SPECIES>
#ifndef NB_LAR_MINB
#define NB_LAR_MINB 3
#endif
__global__ void __launch_bounds__(256, NB_LAR_MINB) k_lar(const __grid_constant__ LArParams P) {
  constexpr bool QUEUE = SPECIES == 1;  // ER: the Lindhard == 1 branch's binomial draws are regrouped (nb_lar.cuh)
  __shared__ __align__(8) LArPending s_q[QUEUE ? 8 : 1][QUEUE ? 64 : 1];
  LArFieldTerms ft;
  ft.field = ft.density = __longlong_as_double(0x7ff8000000000000ll);  // NaN: matches no field
  // a block walks one contiguous range of the list (not a grid stride): lists that sweep the field slowly keep each
  // thread on one field for many events, which is what the memo above needs
  const uint64_t per = ((P.n + gridDim.x - 1) / gridDim.x + blockDim.x - 1) / blockDim.x * blockDim.x;
  const uint64_t lo = blockIdx.x * per, hi = lo + per < P.n ? lo + per : P.n;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const unsigned lt = (1u << lane) - 1u;
  int qn = 0;  // warp-uniform fill of this warp's queue
  // up to 32 pending draws, one candidate each; the rejected go back on the queue
  auto pass = [&]() {
    const int take = qn < 32 ? qn : 32;
    qn -= take;
    const bool live = lane < take;
    LArPending e = live ? s_q[QUEUE ? wid : 0][qn + lane] : LArPending{0., 0., 0u, 0u};
    __syncwarp();
    bool again = false;
    if (live) {
      double Nion;
      const uint64_t i = lo + e.off;
      if (lar_binomial_try(P.rk, P.first_event + i, e.att, e.Nq, Nion)) {
        if (Nion > e.Nq) Nion = e.Nq;
        double f[4];
        lar_fluct_store(e.Nq - Nion, Nion, e.p_r, f);
        double* o = P.fluct + i;
        o[0] = f[0];
        o[P.n] = f[1];
        o[2 * P.n] = f[2];
        o[3 * P.n] = f[3];
      } else {
        again = true;
        ++e.att;
      }
    }
    const unsigned m = __ballot_sync(0xffffffffu, again);
    if (again) s_q[QUEUE ? wid : 0][qn + __popc(m & lt)] = e;
    qn += __popc(m);
    __syncwarp();
  };
  // whole warps walk the range together (the queue's ballots want every lane); the next event's inputs are loaded
  // while this one is computed
  uint64_t i = lo + threadIdx.x;
  double E = i < hi ? P.E[i] : 1., F = i < hi ? P.field[i] : 1.;
  for (uint64_t i0 = lo + (threadIdx.x & ~31u); i0 < hi; i0 += blockDim.x, i += blockDim.x) {
    const bool active = i < hi;
    const uint64_t nx = i + blockDim.x;
    const double En = nx < hi ? P.E[nx] : 1., Fn = nx < hi ? P.field[nx] : 1.;
    LArYields y = lar_yields(SPECIES, E, F, P.density, ft);
    if (P.yields && active) {
      double* o = P.yields + i;
      o[0] = y.TotalYield;
      o[P.n] = y.QuantaYield;
      o[2 * P.n] = y.LightYield;
      o[3 * P.n] = y.Nph;
      o[4 * P.n] = y.Ne;
      o[5 * P.n] = y.Nex;
      o[6 * P.n] = y.Nion;
      o[7 * P.n] = y.Lindhard;
      o[8 * P.n] = y.ElectricField;
    }
    if (P.fluct) {
      double f[4], p_r = 0., pend = -1.;
      if (active) {
        pend = lar_fluct_head(P.rk, P.seed, P.first_event + i, y, QUEUE, p_r, f);
        if (pend < 0.) {
          double* o = P.fluct + i;
          o[0] = f[0];
          o[P.n] = f[1];
          o[2 * P.n] = f[2];
          o[3 * P.n] = f[3];
        }
      }
      if (QUEUE) {
        const unsigned m = __ballot_sync(0xffffffffu, pend >= 0.);
        if (m) {
          if (pend >= 0.) s_q[wid][qn + __popc(m & lt)] = LArPending{pend, p_r, uint32_t(i - lo), 0u};
          qn += __popc(m);
          __syncwarp();
          while (qn >= 32) pass();
        }
      }
    }
    E = En;
    F = Fn;
  }
  if (QUEUE)
    while (qn > 0) pass();
}

}  // namespace nb
