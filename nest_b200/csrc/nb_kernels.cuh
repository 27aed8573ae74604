// nb_kernels.cuh -- the sm_100a kernels: k_event (whole chain, one thread = one event,
// S1 hit loops regrouped by hit count inside the block), k_yields (deterministic means)
// and k_lar (liquid-argon yields + fluctuations).
#pragma once
#include "nb_chain.cuh"
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
  uint64_t n, seed;
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
      const bool nrlike = P.species <= NB200_Cf || P.species == NB200_atmNu;
      if ((nrlike && nearly_equal(P.model.massNum, 0.)) || P.species == NB200_ion) {
        Stream g;
        g.init(P.seed, i, STREAM_MAIN);
        A2 = select_xe_atom(g);
      }
      y = get_yields(P.species, E, P.density, F, P.model.massNum, P.model.atomNum, A2, P.det, P.model,
                     &err);
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
// k_event: the loop body of execNEST() (execNEST.cpp:676-1414) / runNESTvec() (:357-406).
//
// Thread t of a block owns event (batch*BLOCK + t) and keeps its state in registers through
// three phases:
//   A  energy, vertex (drift-window retry), yields, quanta, xy smearing, S1 prelude -> nHits
//   B  the per-hit S1 loop.  Hit counts differ by orders of magnitude between events, so the
//      block counting-sorts its events by nHits in shared memory and thread t runs the loop
//      of the event of rank t: the 32 lanes of a warp then carry near-equal trip counts.
//      Hits draw from Philox blocks keyed (seed; event, hit index), so which thread runs a
//      loop cannot change its result.
//   C  S1 epilogue, S2, signal selection, energy reconstruction, band accumulation
//      (shared-memory integer accumulators, flushed once per block), optional SoA record.
// -------------------------------------------------------------------------------------------
template <int BLOCK>
__device__ __forceinline__ int block_exclusive_scan(int v, int* s_warp) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += t;
  }
  if (lane == 31) s_warp[wid] = x;
  __syncthreads();
  if (wid == 0) {
    int w = (lane < BLOCK / 32) ? s_warp[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += t;
    }
    if (lane < BLOCK / 32) s_warp[lane] = w;
  }
  __syncthreads();
  int base = wid ? s_warp[wid - 1] : 0;
  return base + x - v;
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

template <int BLOCK>
__global__ void __launch_bounds__(BLOCK, 1)
    k_event(const __grid_constant__ RunParams P, int* __restrict__ err_flag) {
  extern __shared__ __align__(16) unsigned long long s_band[];  // [numBins*6] then u32 hist2d
  __shared__ int s_key[BLOCK];
  __shared__ int s_perm[BLOCK];
  __shared__ int s_cnt[BLOCK];
  __shared__ int s_warp[32];
  __shared__ double s_area[BLOCK];
  __shared__ int s_spike[BLOCK];
  __shared__ int s_nphe[BLOCK];

  const int tid = threadIdx.x;
  const nb200_detector& det = P.det;
  const int nb = P.ana.numBins, lb = P.ana.logBins;
  const bool band_smem = P.do_band && P.band_in_smem;
  unsigned int* s_h2 = reinterpret_cast<unsigned int*>(s_band + size_t(nb) * BAND_WORDS);
  if (band_smem) {
    for (int i = tid; i < nb * BAND_WORDS; i += BLOCK) s_band[i] = 0ull;
    for (int i = tid; i < nb * lb; i += BLOCK) s_h2[i] = 0u;
  }
  __syncthreads();

  const uint64_t n_batches = (P.n_events + BLOCK - 1) / BLOCK;
  const double spe_mean = 1. / P.rc.NewMean;
  const bool cli = !P.vec_mode;

  for (uint64_t batch = blockIdx.x; batch < n_batches; batch += gridDim.x) {
    const uint64_t idx = batch * BLOCK + tid;
    const bool active = idx < P.n_events;
    const uint64_t ev = P.first_event + idx;

    // ---------------- phase A ----------------
    Stream g;
    g.init(P.seed, ev, STREAM_MAIN);
    double keV = 0., px = 0., py = 0., pz = 0., sx = 0., sy = 0., field = 0., dt = 0., vD = P.rc.vD;
    Yields y{0., 0., 0., 0., 0., 0.};
    Quanta q{0, 0, 0, 0, 0., 0.};
    S1Pre pre{1., 1., 1., 0, false};
    int err = 0;
    if (active) {
      keV = sample_energy(g, P, idx);
      const nb200_source& S = P.src;
      const bool nonuni = (S.inField == -1.);
      if (S.kind == NB200_SRC_LIST && S.x != nullptr) {
        px = S.x[idx];
        py = S.y[idx];
        pz = S.z[idx];
      } else {
        px = S.xyz[0];
        py = S.xyz[1];
        pz = S.xyz[2];
      }
      const bool ranXY = (S.fixed_pos == 0 || S.fixed_pos == 3) && !(S.kind == NB200_SRC_LIST && S.x);
      const bool ranZ = (S.fixed_pos == 0 || S.fixed_pos == 2) && !(S.kind == NB200_SRC_LIST && S.x);
      for (int guard = 0; guard < 10000; ++guard) {  // Z_NEW loop execNEST.cpp:806-967
        if (ranZ) pz = 0. + (det.TopDrift - 0.) * g.uniform();
        if (ranXY && (guard == 0 || S.fixed_pos == 0)) {
          double r = det.radius * sqrt(g.uniform());
          double sn, cs;
          sincospi(2. * g.uniform(), &sn, &cs);
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

      // yields + quanta: execNEST.cpp:1031-1077 / FullCalculation NEST.cpp:22-42
      const int sp = P.species;
      const bool nrlike = sp <= NB200_Cf || sp == NB200_atmNu;
      if (!cli || keV > 0.001 * P.rc.Wq_eV) {
        int A2 = 131;
        if (P.model.er_model < 0 && ((nrlike && nearly_equal(P.model.massNum, 0.)) || sp == NB200_ion))
          A2 = select_xe_atom(g);
        if (P.model.er_model == 0)
          y = yield_beta(keV, P.rc.rho, field, det, P.model.ERYieldsParam, P.model.multFact);
        else if (P.model.er_model == 1)
          y = yield_gamma(keV, P.rc.rho, field, det, P.model.multFact);
        else if (P.model.er_model == 2)
          y = yield_er_weighted(keV, P.rc.rho, field, det, P.model.ERYieldsParam);
        else
          y = get_yields(sp, keV, P.rc.rho, field, P.model.massNum, P.model.atomNum, A2, det, P.model, &err);
        q = get_quanta(g, y, P.rc.rho, det, P.model, &err);
      }
      if (cli && (det.noiseBaseline[2] != 0. || det.noiseBaseline[3] != 0.))
        q.electrons += to_int_round(gauss(g, det.noiseBaseline[2], det.noiseBaseline[3], false));

      // position smearing execNEST.cpp:1090-1099
      sx = px;
      sy = py;
      if (cli) {
        double Nphd_S2 = fabs(P.rc.g2_params[3]) * q.electrons * exp(-dt / det.eLife_us);
        if (!P.ana.MCtruthPos && Nphd_S2 > NB_PHE_MIN) xy_resolution(g, det, px, py, Nphd_S2, sx, sy);
      }
      pre = s1_prelude(g, P, q.photons, px, py, pz, sx, sy, pz, keV);
    }

    // ---------------- phase B: hit loops regrouped by trip count ----------------
    const int myHits = (active && pre.full) ? pre.nHits : 0;
    const int key = min(myHits, BLOCK - 1);
    s_key[tid] = myHits;
    s_cnt[tid] = 0;
    __syncthreads();
    atomicAdd(&s_cnt[key], 1);
    __syncthreads();
    {
      int c = s_cnt[tid];
      int off = block_exclusive_scan<BLOCK>(c, s_warp);
      s_cnt[tid] = off;
    }
    __syncthreads();
    s_perm[atomicAdd(&s_cnt[key], 1)] = tid;
    __syncthreads();
    {
      // descending: warp 0 takes the longest loops so that they start first
      const int slot = s_perm[BLOCK - 1 - tid];
      const int n = s_key[slot];
      if (n > 0) {
        const uint64_t hev = P.first_event + batch * BLOCK + slot;
        const double eff = s1_eff(det, n);
        const bool need_coin = !(n > det.coinLevel);
        S1Acc acc{0., 0, 0};
        for (int k = 0; k < n; ++k) s1_hit(P, P.seed, hev, uint32_t(k), eff, need_coin, spe_mean, acc);
        s_area[slot] = acc.area;
        s_spike[slot] = acc.spike;
        s_nphe[slot] = acc.nphe;
      }
    }
    __syncthreads();

    // ---------------- phase C ----------------
    if (active) {
      S1Acc acc{0., 0, 0};
      if (pre.full) {
        if (myHits > 0) {
          acc.area = s_area[tid];
          acc.spike = s_spike[tid];
          acc.nphe = s_nphe[tid];
        }
      } else
        s1_parametric(g, P, pre.nHits, acc);
      double scint[9], scint2[9];
      s1_epilogue(g, P, pre, acc, scint);
      int Ne = q.electrons;
      if (cli && pz < det.cathode) Ne = 0;  // execNEST.cpp:1107
      get_s2(g, P, Ne, px, py, pz, sx, sy, dt, field, scint2);
      double signal1 = 0., signal2 = 0., signalE = 0.;
      const double keV_true = keV;
      if (cli) {
        select_signals(P.ana, scint, scint2, signal1, signal2);
        signalE = reconstruct_energy(P, y, scint, scint2, signal1, signal2, field, keV);
        if (P.do_band) {
          bool acc_ok = false;
          double xv = 0., yv = 0.;
          int bin = band_bin(P, signal1, signal2, acc_ok, xv, yv);
          if (bin >= 0) {
            unsigned long long* B = band_smem ? s_band : P.band;
            unsigned long long* b = B + size_t(bin) * BAND_WORDS;
            if (acc_ok) {
              atomicAdd(&b[0], 1ull);
              atomicAdd(&b[2], (unsigned long long)(long long)llrint(xv * NB200_FX_S1));
              atomicAdd(&b[3], (unsigned long long)(long long)llrint(yv * NB200_FX_Y));
              atomicAdd(&b[4], (unsigned long long)(long long)llrint(yv * yv * NB200_FX_Y2));
              atomicAdd(&b[5], (unsigned long long)(long long)llrint(yv * yv * yv * NB200_FX_Y3));
              if (yv != 0. && lb > 0) {
                int jb = int((yv - P.ana.logMin) / (P.ana.logMax - P.ana.logMin) * double(lb));
                jb = min(max(jb, 0), lb - 1);
                if (band_smem)
                  atomicAdd(&s_h2[bin * lb + jb], 1u);
                else
                  atomicAdd(&P.band[size_t(nb) * BAND_WORDS + size_t(bin) * lb + jb], 1ull);
              }
            } else
              atomicAdd(&b[1], 1ull);
          }
        }
      }
      if (P.write_soa) {
        const SoaDev& o = P.soa;
        if (o.energy_kev) o.energy_kev[idx] = keV_true;
        if (o.x_mm) o.x_mm[idx] = px;
        if (o.y_mm) o.y_mm[idx] = py;
        if (o.z_mm) o.z_mm[idx] = pz;
        if (o.field) o.field[idx] = field;
        if (o.driftTime) o.driftTime[idx] = dt;
        if (o.smear_x) o.smear_x[idx] = sx;
        if (o.smear_y) o.smear_y[idx] = sy;
        if (o.photons) o.photons[idx] = q.photons;
        if (o.electrons) o.electrons[idx] = Ne;
        if (o.ions) o.ions[idx] = q.ions;
        if (o.excitons) o.excitons[idx] = q.excitons;
#pragma unroll
        for (int k = 0; k < 9; ++k) {
          if (o.s1[k]) o.s1[k][idx] = scint[k];
          if (o.s2[k]) o.s2[k][idx] = scint2[k];
        }
        if (o.signal1) o.signal1[idx] = signal1;
        if (o.signal2) o.signal2[idx] = signal2;
        if (o.signalE) o.signalE[idx] = cli ? signalE : keV;
        if (o.Nph_mean) o.Nph_mean[idx] = y.Nph;
        if (o.Ne_mean) o.Ne_mean[idx] = y.Ne;
      }
      if (err) atomicOr(err_flag, err);
    }
  }

  if (band_smem) {
    __syncthreads();
    for (int i = tid; i < nb * BAND_WORDS; i += BLOCK) {
      unsigned long long v = s_band[i];
      if (v) atomicAdd(&P.band[i], v);
    }
    for (int i = tid; i < nb * lb; i += BLOCK) {
      unsigned int v = s_h2[i];
      if (v) atomicAdd(&P.band[size_t(nb) * BAND_WORDS + i], (unsigned long long)v);
    }
  }
}

// -------------------------------------------------------------------------------------------
// k_lar: LArNEST::FullCalculation (src/LArNEST.cpp:17-33) for species NR/ER/Alpha
// -------------------------------------------------------------------------------------------
struct LArParams {
  int32_t species;
  uint64_t n, seed, first_event;
  double density;
  const double *E, *field;
  double *yields, *fluct;  // [9][n], [4][n]; either may be null
};

__global__ void __launch_bounds__(256) k_lar(const __grid_constant__ LArParams P) {
  for (uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; i < P.n;
       i += uint64_t(gridDim.x) * blockDim.x) {
    LArYields y = lar_yields(P.species, P.E[i], P.field[i], P.density);
    if (P.yields) {
      P.yields[i] = y.TotalYield;
      P.yields[P.n + i] = y.QuantaYield;
      P.yields[2 * P.n + i] = y.LightYield;
      P.yields[3 * P.n + i] = y.Nph;
      P.yields[4 * P.n + i] = y.Ne;
      P.yields[5 * P.n + i] = y.Nex;
      P.yields[6 * P.n + i] = y.Nion;
      P.yields[7 * P.n + i] = y.Lindhard;
      P.yields[8 * P.n + i] = y.ElectricField;
    }
    if (P.fluct) {
      Stream g;
      g.init(P.seed, P.first_event + i, STREAM_LAR);
      double f[4];
      lar_fluctuations(g, y, f);
      P.fluct[i] = f[0];
      P.fluct[P.n + i] = f[1];
      P.fluct[2 * P.n + i] = f[2];
      P.fluct[3 * P.n + i] = f[3];
    }
  }
}

}  // namespace nb
