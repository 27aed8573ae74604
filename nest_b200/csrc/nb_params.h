// nb_params.h -- the kernel-parameter block: the C-ABI PODs of include/nest_b200.h
// with device pointers substituted, plus per-run constants hoisted out of the event loop
// on the host.  Passed by value as a __grid_constant__ kernel argument (constant bank),
// the B200 stand-in for the reference's VDetector* getter calls inside the hot loop
// (src/NEST.cpp:1288-2063).
#pragma once
#include <stdint.h>

#include "../../include/nest_b200.h"

#ifndef __CUDACC__
struct uint2 {  // plain C++ builds of these headers (tests/host_checks.cpp): the one CUDA vector type named below
  unsigned int x, y;
};
#endif

namespace nb {

// packed band accumulator in device memory (u64 words), per S1 bin
//   [0] accepted  [1] rejected  [2] sumS1  [3] sumY  [4] sumY2  [5] sumY3
// followed by numBins*logBins u64 of the 2-D histogram
enum { BAND_WORDS = 6 };

struct SoaDev {  // nb200_soa_out with device pointers (null = column not written)
  double *energy_kev, *x_mm, *y_mm, *z_mm, *field, *driftTime, *smear_x, *smear_y;
  int32_t *photons, *electrons, *ions, *excitons;
  double* s1[9];
  double* s2[9];
  double *signal1, *signal2, *signalE, *Nph_mean, *Ne_mean, *deltaT_ns, *keVee;
};

// per-event state handed from k_pre to k_hits to k_post (device workspace, one chunk)
struct Work {
  double *keV, *px, *py, *pz, *sx, *sy, *field, *dt, *yNph, *yNe, *yL, *posDepSm, *fitSm, *area;
  int32_t *photons, *electrons, *ions, *excitons, *nHits, *spike, *nphe;
};
enum { WORK_F64 = 14, WORK_I32 = 7 };

// field/density-only factors of the mean-yield models, see nb_yields.cuh
struct YieldHoist {
  double field, density;  // what the factors were evaluated for (field < 0: none)
  double nr_ti;           // NR: ThomasImel
  double b[6];            // beta: m01, m03, m04, m07, m10, density correction of Qy
  double g[3];            // gamma: m1, m5, m7
  double w[3];            // RecombOmegaER: amplitude(field), the two skew-Gaussian normalisations
};

struct RunParams {
  nb200_detector det;  // map.lut -> device memory
  nb200_model model;
  nb200_analysis ana;
  nb200_source src;  // E/x/y/z/wimp_* -> device memory
  nb200_run_consts rc;

  // hoisted per-run constants
  double s1_center;      // FitS1(0,0,TopDrift - vD_middle*dtCntr)   NEST.cpp:1317-1321
  double s2_center;      // FitS2(0,0)                               NEST.cpp:1762-1763
  unsigned long long dphe_thr;  // double-phe decision on a 32-bit word w: w < dphe_thr
  // single photo-electron area phe = T + N, T ~ N(1/NewMean, sPEres^2) truncated to T > 0 and
  // N ~ N(noiseBaseline[0], noiseBaseline[1]^2) (NEST.cpp:1370-1376), sampled through its sum:
  // S ~ N(pe_mu, pe_sig^2) accepted with P(T > 0 | S) where T | S ~ N(pe_m0 + pe_rho S, pe_tau^2);
  // above pe_cut that probability is 1 to within 2^-64 and no second normal is drawn
  double pe_mu, pe_sig, pe_rho, pe_m0, pe_tau, pe_cut;
  // below pe_cut the acceptance P(T > 0 | S) = Phi(m/tau), m = pe_m0 + pe_rho S, is first tried
  // against a 4-bit uniform: for m >= pe_fast = 1.5342 tau (Phi >= 15/16) the draw u4 < 15 accepts
  // outright; only u4 == 15 or m < pe_fast needs Phi itself (exact residual test in the queue)
  double pe_fast;
  double pe_sfast;       // the same threshold on S: m < pe_fast <=> S < pe_sfast (pe_rho > 0)
  double pe_inv_tau16;   // 16 / pe_tau: index scale of the Phi bounds table (pe_accept_exact)
  const uint2* dpe_alias;  // alias tables of Binomial(2^b, P_dphe) (nb_rng.cuh: binomial_fixed_p), or null
  double coin_u_min;     // exp(-coinWind/20): -20 ln u < coinWind <=> u > this  NEST.cpp:1422
  double coin_pow[11];   // pow(numPMTs, 1 - k), k = 0..10: the 2-fold coincidence probability  NEST.cpp:1655
  double below_thr_pct;  // Parametric S1 threshold percentile       NEST.cpp:1437-1457
  double recon_mult;     // MultFact prefactor, execNEST.cpp:1179-1182
  double wimp_dr0[9];    // WIMP_dRate(0) per Xe isotope             TestSpectra.cpp:459,466
  YieldHoist hoist;      // yield-model factors of (inField, rho), uniform field only
  // uniform field + random z: the drift-window retry (execNEST.cpp:962-967) keeps exactly the z with
  // dt = (TopDrift - z)/vD in [dt_min, dt_max], so z is drawn uniformly on that interval instead of
  // redrawing 26 % of the LUX vertices (z_direct = 0: keep the retry loop)
  double z_lo, z_hi;
  int32_t z_direct, _pad_z;
  double band_width, band_border;  // GetBand bin geometry           execNEST.cpp:1521-1527
  double band_fx_x;                // fixed-point scale of the band's sum of X (nb200_band_scale_x)
  // GetDriftVelocity_Liquid rows bracketing T_Kelvin (NEST.cpp:2429-2502); dv_mode 0 blend,
  // 1 = T on the lower node, 2 = T on the upper node
  double dv_row[2][7];
  double dv_Ti, dv_Tf, dv_T;
  int32_t dv_mode;
  int32_t vtable_n;      // SetDriftVelocity_NonUniform table (NEST.cpp:2664-2697), inField == -1
  const double* vtable;

  uint64_t seed, first_event, n_events;
  uint32_t rk[20];     // Philox round keys of `seed` (nb_rng.cuh: philox_round_keys)
  uint64_t chunk_off;  // index of this chunk's first event within the call (lists, SoA columns)
  int32_t species;
  int32_t vec_mode;   // 1 = runNESTvec semantics (execNEST.cpp:357-406)
  int32_t write_soa;
  int32_t do_band;
  int32_t band_in_smem;  // accumulate the band in shared memory (fits) or straight in global
  int32_t _pad;
  SoaDev soa;
  Work work;
  unsigned long long* band;  // device accumulator or null
};

}  // namespace nb
