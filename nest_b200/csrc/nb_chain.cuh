// nb_chain.cuh -- the stochastic stages of the per-event chain (device).
//
// GetQuanta, GetS1 (split into prelude / per-hit / epilogue so that the hit loop can be
// regrouped across threads), GetS2, xyResolution, the TestSpectra samplers, the
// execNEST() signal selection + energy reconstruction and the GetBand binning.
// Reference lines are cited per function (src/NEST.cpp, src/execNEST.cpp,
// src/TestSpectra.cpp of NESTCollaboration/nest).
#pragma once
#include "nb_yields.cuh"

namespace nb {
#ifdef __CUDACC__

struct Quanta {  // QuantaResult NEST.hh:171-178
  int photons, electrons, ions, excitons;
  double recombProb, Variance;
};

// NESTcalc::RecombOmegaNR NEST.cpp:164-171
NB_D double recomb_omega_nr(double elecFrac, const double* w) {
  double d = elecFrac - w[3];
  double omega = w[2] * exp(-0.5 * (d * d) / (w[4] * w[4]));
  if (omega < 0.) omega = 0;
  return omega;
}

// NESTcalc::RecombOmegaER NEST.cpp:173-225 (the wide > cntr throw is checked on the host).
// omega_er_consts: the factors that depend on the field and the width vector only -- amplitude
// (two pow) and the normalisations of the two skew-Gaussians (exp, erf, sqrt each); hoisted to the
// host for a uniform field like the yield factors (YieldHoist::w).
NB_HD void omega_er_consts(double efield, const double* w, int nw, double* o) {
  double ampl = 0.07 + (w[7] - 0.07) / pow(1. + pow(efield / 295.2, 251.6), 0.0069114);
  if (ampl < 0.) ampl = 0.;
  const double cntr = w[8], wide = w[9], skew = w[10];
  const double mode = cntr + 2. * NB_INV_SQRT2PI * skew * wide / sqrt(1. + skew * skew);
  const double dm = mode - cntr;
  o[0] = ampl;
  o[1] = 1. / (exp(-0.5 * (dm * dm) / (wide * wide)) * (1. + erf(skew * (mode - cntr) / (wide * NB_SQRT2))));
  double cntr2 = 5.207, wide2 = 1.361, skew2 = -2.495;
  if (nw > 15) {
    cntr2 = w[14];
    wide2 = w[15];
  }
  if (nw > 16) skew2 = w[16];
  const double mode2 = cntr2 + 2. * NB_INV_SQRT2PI * skew2 * wide2 / sqrt(1. + skew2 * skew2);
  const double dm2 = mode2 - cntr2;
  o[2] = 1. / (exp(-0.5 * (dm2 * dm2) / (wide2 * wide2)) * (1. + erf(skew2 * (mode2 - cntr2) / (wide2 * NB_SQRT2))));
}
NB_D double recomb_omega_er(double efield, double elecFrac, double numQuanta, const double* w, int nw,
                            const YieldHoist* h) {
  double cl[3];
  const double* c = cl;
  if (h && h->field == efield)
    c = h->w;
  else
    omega_er_consts(efield, w, nw, cl);
  const double ampl = c[0], norm = c[1];
  const double cntr = w[8], wide = w[9], skew = w[10];
  double omega;
  if (cntr < 1.) {
    double d = elecFrac - cntr;
    omega = norm * ampl * exp(-0.5 * (d * d) / (wide * wide)) *
            (1. + erf(skew * (elecFrac - cntr) / (wide * NB_SQRT2)));
  } else {
    double ampl2 = 0.06103, cntr2 = 5.207, wide2 = 1.361, skew2 = -2.495;
    if (nw > 15) {
      ampl2 = w[13];
      cntr2 = w[14];
      wide2 = w[15];
    }
    if (nw > 16) skew2 = w[16];
    const double norm2 = c[2];
    double logNQ = log10(numQuanta);
    double d1 = logNQ - cntr, d2 = logNQ - cntr2;
    omega = 0.00 +
            norm * ampl * exp(-0.5 * (d1 * d1) / (wide * wide)) *
                (1. + erf(skew * (logNQ - cntr) / (wide * NB_SQRT2))) +
            norm2 * ampl2 * exp(-0.5 * (d2 * d2) / (wide2 * wide2)) *
                (1. + erf(skew2 * (logNQ - cntr2) / (wide2 * NB_SQRT2)));
  }
  if (omega < 0.) omega = 0;
  return omega;
}

// NESTcalc::FanoER NEST.cpp:227-246
NB_D double fano_er(double density, double Nq_mean, double efield, const double* w, bool inGas) {
  double Fano = 0.12707 - 0.029623 * density - 0.0057042 * (density * density) +
                0.0015957 * (density * density * density);
  if (!inGas) {
    if (w[6] < 0.)
      Fano += fabs(w[6]) * sqrt(Nq_mean) * sqrt(efield);
    else
      Fano = w[6];
  }
  return Fano;
}

NB_D int to_int_round(double x) { return int(floor(x + 0.5)); }

// NESTcalc::GetQuanta NEST.cpp:248-432
NB_D Quanta get_quanta(Stream& g, uint64_t ev, const Yields& y, double density,
                       const nb200_detector& det, const nb200_model& mo, int* err, const YieldHoist* h = nullptr) {
  const double* w = mo.NRERWidthsParam;
  Quanta q{0, 0, 0, 0, 0., 0.};
  bool HighE;
  int Nq_actual, Ne, Nph, Ni, Nex;
  double excitonRatio = y.NexONi;
  double Nq_mean = y.Nph + y.Ne;
  double elecFrac = smax(0., smin(y.Ne / Nq_mean, 1.));
  if (excitonRatio < 0.) {
    excitonRatio = 0.;
    HighE = true;
  } else
    HighE = false;
  double alf = 1. / (1. + excitonRatio);
  double recombProb = 1. - (excitonRatio + 1.) * elecFrac;
  if (recombProb < 0.) excitonRatio = 1. / elecFrac - 1.;

  const bool isER = nearly_equal(y.L, 1.);
  if (isER) {
    if (nearly_equal(Nq_mean, 0.))
      Nq_actual = 0;
    else {
      double Fano = fano_er(density, Nq_mean, y.field, w, det.inGas != 0);
      Nq_actual = to_int_round(gauss(g, Nq_mean, sqrt(Fano * Nq_mean), true));
    }
    Ni = int(binomial(g.rk, ev, STREAM_BIN_NI, Nq_actual, alf));
    Nex = Nq_actual - Ni;
  } else {
    // two independent Gaussians (:294-311); both come from one Box-Muller pair
    double z0, z1;
    normal_pair(g, z0, z1);
    double Ni_mean = Nq_mean * alf;
    double Fano = w[0] + w[11] * Ni_mean;
    if (Fano < 0.) Fano = 0.;
    double d = Ni_mean + sqrt(Fano * Ni_mean) * z0;
    Ni = to_int_round((d < 0.) ? 0. : d);
    double Nex_mean = Nq_mean * excitonRatio * alf;
    Fano = w[1] + w[12] * Nex_mean;
    if (Fano < 0.) Fano = 0.;
    d = Nex_mean + sqrt(Fano * Nex_mean) * z1;
    Nex = to_int_round((d < 0.) ? 0. : d);
    Nq_actual = Nex + Ni;
  }
  if (Nq_actual == 0) return q;

  if (Nex < 0) Nex = 0;
  if (Ni < 0) Ni = 0;
  if (Nex > Nq_actual) Nex = Nq_actual;
  if (Ni > Nq_actual) Ni = Nq_actual;
  q.ions = Ni;
  q.excitons = Nex;

  if (Nex <= 0 && HighE) recombProb = y.Nph / double(Ni);
  recombProb = smax(0., smin(recombProb, 1.));
  if (isnan(recombProb) || isnan(elecFrac) || Ni == 0 || nearly_equal(recombProb, 0.0)) {
    q.photons = Nex;
    q.electrons = Ni;
    q.recombProb = 0.;
    q.Variance = 0.;
    return q;
  }

  double omega = y.L < 1 ? recomb_omega_nr(elecFrac, w)
                         : recomb_omega_er(y.field, elecFrac, Nq_mean, w, mo.n_widths, h);
  double lambdaChen;
  if (w[2] < 0. && y.L < 1)
    lambdaChen = 1.0 + w[2];
  else
    lambdaChen = 1.0;
  if (lambdaChen <= 0.) lambdaChen = 1E-6;
  double Variance = lambdaChen * recombProb * (1. - recombProb) * Ni + omega * omega * Ni * Ni;

  double skewness;
  if ((y.Nph + y.Ne) > 1e4 || y.field > 4e3 || y.field < 50.) {
    skewness = 0.00;
  } else {
    skewness = 0.;
    if (isER) {
      if (mo.SkewnessER != -999.) {
        skewness = mo.SkewnessER;
      } else {  // LUX skewness model :366-388
        Wvalue wv = work_function(density, det.molarMass, det.OldW13eV != 0);
        double engy = 1e-3 * wv.Wq_eV * (y.Nph + y.Ne);
        double fld = y.field;
        const double alpha0 = 1.39, cc0 = 4.0, cc1 = 22.1, E0 = 7.7, E1 = 54., E2 = 26.7, E3 = 6.4,
                     F0 = 225., F1 = 71.;
        skewness = 1. / (1. + exp((engy - E2) / E3)) *
                       (alpha0 + cc0 * exp(-1. * fld / F0) * (1. - exp(-1. * engy / E0))) +
                   1. / (1. + exp(-1. * (engy - E2) / E3)) * cc1 * exp(-1. * engy / E1) *
                       exp(-1. * sqrt(fld) / sqrt(F1));
      }
    } else {
      skewness = w[5];
    }
  }

  double widthCorrection = sqrt(1. - (2. / NB_PI) * skewness * skewness / (1. + skewness * skewness));
  double muCorrection = (sqrt(Variance) / widthCorrection) *
                        (skewness / sqrt(1. + skewness * skewness)) * 2. * NB_INV_SQRT2PI;
  if (fabs(skewness) > DBL_MIN)
    Ne = to_int_round(skew_gauss(g, (1. - recombProb) * Ni - muCorrection,
                                 sqrt(Variance) / widthCorrection, skewness));
  else
    Ne = to_int_round(gauss(g, (1. - recombProb) * Ni, sqrt(Variance), true));

  Ne = min(max(Ne, 0), Ni);  // NESTcalc::clamp
  Nph = Nq_actual - Ne;
  if (Nph > Nq_actual) Nph = Nq_actual;
  if (Nph < Nex) Nph = Nex;
  if ((Nph + Ne) != (Nex + Ni)) *err = 1;
  q.Variance = Variance;
  q.recombProb = recombProb;
  q.photons = Nph;
  q.electrons = Ne;
  return q;
}

// NESTcalc::xyResolution NEST.cpp:2699-2736
NB_D void xy_resolution(Stream& g, const nb200_detector& det, double x, double y, double A_top,
                        double& sx, double& sy) {
  A_top *= 1. - fit_tba(det.FitTBA, x, y, det.TopDrift / 2.);
  double kappa = det.PosResBase;
  double sigmaR = kappa / sqrt(A_top);
  sigmaR = sqrt(sigmaR * sigmaR + det.PosResFlat * det.PosResFlat);
  uint32_t pa, pb;
  g.next2(pa, pb);  // phi = 2 pi u, u from 32 bits
  sigmaR = fabs(gauss(g, 0.0, sigmaR, false));
  double s, c;
  sincos_bits(pa, s, c);
  double sigmaX = sigmaR * c, sigmaY = sigmaR * s;
  if (sigmaR > 1e2 || isnan(sigmaR) || sigmaR < 0. || fabs(sigmaX) > 1e2 || fabs(sigmaY) > 1e2) {
    if (A_top > 40.) {
      sigmaX = 0.;
      sigmaY = 0.;
    }
  }
  sx = x + sigmaX;
  sy = y + sigmaY;
}

// ---- S1 ---------------------------------------------------------------------------------
struct S1Pre {
  double posDepSm, fitSm;  // FitS1(smear)/centre, and FitS1(smear) itself (GetSpike)
  double eff;
  int nHits;
  int nDP;    // Full mode: how many of the hits carry a second photo-electron (see k_hits)
  bool full;  // Full per-hit loop (true) or Parametric
};

// eff ramp NEST.cpp:1339-1347
NB_D double s1_eff(const nb200_detector& det, int nHits) {
  double eff = det.sPEeff;
  if (eff < 1.) eff += ((1. - eff) / (2. * double(det.numPMTs))) * double(nHits);
  return smax(0., smin(eff, 1.));
}

// GetS1 up to the per-hit loop: NEST.cpp:1288-1362
NB_D S1Pre s1_prelude(const uint32_t* rk, uint64_t ev, const RunParams& P, int Nph, double tx, double ty, double tz,
                      double sx, double sy, double sz, double energy) {
  const nb200_detector& det = P.det;
  S1Pre r;
  double posDep = fit_s1(det.FitS1, tx, ty, tz);
  r.fitSm = fit_s1(det.FitS1, sx, sy, sz);
  posDep /= P.s1_center;
  r.posDepSm = r.fitSm / P.s1_center;
  double g1_XYZ = smax(0., smin(det.g1 * posDep, 1.));
  if (nearly_equal(det.g1, 1.)) {
    g1_XYZ = 1.;
    r.posDepSm = 1.;
  }
  if (nearly_equal(det.g1, 0.)) {
    g1_XYZ = 0.;
    r.posDepSm = 0.;
  }
  r.nHits = int(binomial(rk, ev, STREAM_BIN_S1, Nph, g1_XYZ));
  r.eff = s1_eff(det, r.nHits);
  int mode = P.ana.s1mode;
  if (mode == NB200_S1_HYBRID) mode = energy < 100. ? NB200_S1_FULL : NB200_S1_PARAMETRIC;
  r.full = (mode == NB200_S1_FULL);
  // NEST.cpp:1391 decides hit by hit, with probability P_dphe, whether a second photo-electron
  // follows; the hits are exchangeable, so only their number matters: one binomial here, and k_hits
  // makes the first nDP hits the double ones
  const int nh = r.full ? r.nHits : 0;
  const bool tab = P.dpe_alias != nullptr && nh < NB_ALIAS_MAXN;  // P_dphe in (0,1) and a moderate hit count
  if (tab)
    r.nDP = binomial_fixed_p(rk, ev, STREAM_BIN_S1DPE, nh, P.dpe_alias);
  else  // large events (or P_dphe of 0 / 1): the general sampler
    r.nDP = int(binomial(rk, ev, STREAM_BIN_S1DPE + 1u, nh, det.P_dphe));
  return r;
}

struct S1Acc {
  double area;
  int spike, nphe;
};

// ---- the per-hit S1 loop, NEST.cpp:1367-1429 -------------------------------------------------
// Per PMT hit the reference draws phe = T + N with T a zero-truncated Gaussian(1/NewMean, sPEres)
// and N the Gaussian baseline noise, kills it with probability 1 - eff, adds a second such
// photo-electron with probability P_dphe, and counts the hit (spike + area) when the summed area
// clears sPEthr and, for events with nHits <= coinLevel, a coincidence-window draw.
//
// Same distributions, drawn differently (all exact to within the 2^-64 resolution of the
// reference's own generator):
//   * a photo-electron area is sampled through its SUM S = T + N ~ N(pe_mu, pe_sig^2), accepted
//     with probability P(T > 0 | S): one normal per photo-electron instead of two.  Above pe_cut
//     that probability is 1 - 2^-64 and nothing else is drawn; below it (22 % of LUX hits) T | S is
//     drawn and the pair redrawn when T <= 0 (0.36 %);
//     The accept test itself never needs T: it is a Bernoulli(Phi(m/tau)) draw, tried first
//     against a 4-bit uniform (u4 < 15 accepts when Phi >= 15/16) with the exact residual
//     probability evaluated (erfc) only when that fails -- 2 % of hits;
//   * that normal comes from a 2048-layer ziggurat (nb_rng.cuh) instead of Box-Muller;
//   * the hits of an event are exchangeable, so the number of double-phe hits is drawn once per
//     event (k_pre: Binomial(nHits, P_dphe) from alias tables) instead of a Bernoulli per hit, and one
//     Philox block serves a SLOT of two photo-electrons -- one double hit or two single hits:
//     2 x (28 bits of abscissa, 11 of layer, 4 for the truncation pre-test);
//   * slots needing anything more (wedge / tail of the ziggurat, exact truncation test, redraw,
//     efficiency or coincidence draws) are queued per warp and redone in full 32 at a time (k_hits);
//   * -20 ln u < coinWind is tested as u > exp(-coinWind/20) and only when nHits <= coinLevel.

// Phi(z) at z = -8 + k/16, k = 0..304, filled by nb200_create (host erfc).  Linear interpolation
// between neighbours is within 1.18e-4 of Phi (h^2/8 max|Phi''| = (1/16)^2/8 * 0.242).
// (up to z = 11: the pre-test sends everything below m/tau = 9.1 here, and arguments past the table
// end evaluate erfc -- with a table ending at 9 that was 1 % of the tests, and the reason erfc ran
// in nearly every finisher pass)
#define NB_PHI_N 305
#define NB_PHI_EPS 1.25e-4
__device__ double g_phitab[NB_PHI_N];

// exact residual acceptance of a photo-electron sum s that failed the 4-bit pre-test:
// P(accept) = Phi(m/tau) overall, of which 15/16 was already granted when m >= pe_fast.
// The comparison u < p is first tried against rigorous bounds of Phi from the table (decides all
// but ~0.4 % of the cases with two loads and a handful of FMAs); only a u inside the bracket
// evaluates erfc.  The outcome is that of the erfc test in every case.
NB_D bool pe_accept_exact(const RunParams& P, double sv, double u) {
  const double m = fma(P.pe_rho, sv, P.pe_m0);
  const bool scaled = !(sv < P.pe_sfast);  // the pre-test's own comparison (m >= pe_fast)
#ifndef NB_PE_NO_BOUNDS
  if (P.pe_tau > 0.) {
    const double t = fma(m, P.pe_inv_tau16, 128.);  // (z + 8) * 16
    if (t >= 0. && t < double(NB_PHI_N - 2)) {
      const int k = int(t);
      const double f = t - double(k);
      const double a = __ldg(&g_phitab[k]), b = __ldg(&g_phitab[k + 1]);
      const double L = fma(f, b - a, a);
      double lo = L - NB_PHI_EPS, hi = L + NB_PHI_EPS;
      if (scaled) {
        lo = fma(16., lo, -15.);
        hi = fma(16., hi, -15.);
      }
      if (u < lo) return true;
      if (u >= hi) return false;
    }
  }
#endif
  double phi;
  if (P.pe_tau > 0.)
    phi = 0.5 * erfc(-(m / P.pe_tau) * 0.70710678118654752440);
  else
    phi = (m > 0.) ? 1. : 0.;
  const double p = scaled ? fma(16., phi, -15.) : phi;
  return u < p;
}

// joint redraw of (S, T) after a failed truncation check (0.35 % of photo-electrons): one block per
// attempt, both normals from the hit loop's ziggurat (tables read from global memory here)
NB_D double zig_normal_g(uint32_t w0, uint32_t w1, uint32_t e0, uint32_t e1, uint32_t k, uint32_t stream, uint32_t k0,
                         uint32_t k1) {
  const int32_t t = zig_t(w0);
  const uint32_t i = w1 >> (32 - NB_ZIG_BITS);
  if (uint32_t(abs(t)) < __ldg(&g_zig_k[i])) return zig_td(t) * __ldg(&g_zig_w[i]);
  return zig_finish(w0, i, 0u, false, e0, e1, k, stream, k0, k1);
}
__device__ __noinline__ double pe_redraw(uint64_t seed, uint64_t event, uint32_t blk, double pe_mu, double pe_sig,
                                         double pe_cut, double pe_tau, double pe_rho, double pe_m0) {
  const uint32_t k0 = uint32_t(seed), k1 = uint32_t(seed >> 32), e0 = uint32_t(event), e1 = uint32_t(event >> 32);
  double sv = pe_mu;
#pragma unroll 1
  for (uint32_t a = 0; a < 32u; ++a) {  // callers space blk by 32
    const u128 b = philox_block(e0, e1, blk + a, STREAM_HITX, k0, k1);
    sv = fma(pe_sig, zig_normal_g(b.x, b.y, e0, e1, blk + a, STREAM_ZIGX, k0, k1), pe_mu);
    if (sv > pe_cut) break;
    const double zt = zig_normal_g(b.z, b.w, e0, e1, blk + a, STREAM_ZIGX + 1u, k0, k1);
    if (fma(pe_tau, zt, fma(pe_rho, sv, pe_m0)) > 0.) break;
  }
  return sv;
}

// Parametric S1, NEST.cpp:1432-1481
NB_D void s1_parametric(Stream& g, uint64_t ev, const RunParams& P, int nHits, S1Acc& acc) {
  const nb200_detector& det = P.det;
  int Nphe = nHits + int(binomial(g.rk, ev, STREAM_BIN_PAR0, nHits, det.P_dphe));
  double eff = s1_eff(det, nHits);
  double Nphe_det = double(binomial(g.rk, ev, STREAM_BIN_PAR1, Nphe, 1. - (1. - eff) / (1. + det.P_dphe)));
  acc.area = gauss(g, Nphe_det * (1. / P.rc.NewMean), det.sPEres * sqrt(Nphe_det / P.rc.NewMean), true);
  acc.spike = int(binomial(g.rk, ev, STREAM_BIN_PAR2, nHits, eff * (1. - P.below_thr_pct)));
  acc.nphe = Nphe;
}

// n-choose-r as the ratio of factorials the reference forms with tgammal (NEST.cpp:2237-2241)
NB_D double n_choose_r(double n, int r) {
  double c = 1.;
  for (int j = 1; j <= r; ++j) c = c * (n - double(r) + double(j)) / double(j);
  return c;
}

// The normals and the yes/no word that every event consumes after the hit loop, drawn together at the top of the
// stage -- two Philox blocks at full lane occupancy instead of three to four calls from inside divergent branches
// (ncu: philox_block was 15 % of k_post's warp instructions at 15 active lanes).  All independent:
//   z_spike  GetSpike's smearing (NEST.cpp:2249-2256)      z_s2noise  S2 linear/quadratic noise (:1994-2003)
//   z_s2ph   S2 photon number (:1976-1980)                 z_s2area   S2 pulse area (:1986)
//   u_coin   the n-fold coincidence draw of GetS1 (:1689): a 32-bit uniform, like the chain's other yes/no draws
struct PostDraws {
  double z_spike, z_s2noise, z_s2ph, z_s2area;
  uint32_t u_coin;
};
NB_D PostDraws post_draws(Stream& g) {
  PostDraws d;
  uint32_t spare;
  normal_pair_spare(g, d.z_spike, d.z_s2noise, d.u_coin);
  normal_pair_spare(g, d.z_s2ph, d.z_s2area, spare);
  return d;
}

// GetS1 after the hit loop: noise, thresholds, corrections, n-fold coincidence, GetSpike,
// below-threshold sign flag.  NEST.cpp:1608-1721, 2243-2268.  out[9] = scintillation.
NB_D void s1_epilogue(Stream& g, const RunParams& P, const S1Pre& pre, const S1Acc& acc, const PostDraws& D,
                      double* out) {
  const nb200_detector& det = P.det;
  double pulseArea = acc.area;
  double spike = double(acc.spike);
  double sig;
  if (det.noiseQuadratic[0] != 0) {
    double a = det.noiseQuadratic[0] * pulseArea * pulseArea, b = det.noiseLinear[0] * pulseArea;
    sig = sqrt(a * a + b * b);
  } else
    sig = det.noiseLinear[0] * pulseArea;
  // gauss(mu, 0, true) == max(mu, 0): skip the draw when the width is exactly zero
  if (sig != 0.) pulseArea = pulseArea + sig * normal(g);
  pulseArea = (pulseArea < 0.) ? 0. : pulseArea;
  if (pulseArea < det.sPEthr) pulseArea = 0.;
  if (spike < 0) spike = 0;
  double pulseAreaC = div_z(pulseArea, pre.posDepSm);
  double Nphd = div_z(pulseArea, 1. + det.P_dphe);
  double NphdC = div_z(pulseAreaC, 1. + det.P_dphe);
  double spikeC = div_z(spike, pre.posDepSm);
  out[0] = pre.nHits;
  out[1] = acc.nphe;
  out[2] = pulseArea;
  out[3] = pulseAreaC;
  out[4] = Nphd;
  out[5] = NphdC;
  out[6] = spike;
  out[7] = spikeC;
  out[8] = spike;

  double prob;
  if (spike < det.coinLevel)
    prob = 0.;
  else if (spike > 10.)
    prob = 1.;
  else if (det.coinLevel == 0)
    prob = 1.;
  else if (det.coinLevel == 1)
    prob = (spike >= 1.) ? 1. : 0.;
  else if (det.coinLevel == 2)
    prob = 1. - P.coin_pow[int(spike)];  // pow(numPMTs, 1 - spike), 2 <= spike <= 10: tabulated by the host's libm
  else {
    double numer = 0., denom = 0.;
    for (int i = int(spike); i > 0; --i) {
      double c = n_choose_r(double(det.numPMTs), i);
      denom += c;
      if (i >= det.coinLevel) numer += c;
    }
    prob = numer / denom;
  }

  if (spike >= 1. && spikeC > 0.) {  // GetSpike NEST.cpp:2243-2268
    if (spikeC > NB_SPIKES_MAXM) {
      out[6] = Nphd;
      out[7] = NphdC;
    } else {
      // rand_zero_trunc_gauss RandomGen.cpp:55-61 (redraw while <= 0)
      const double sw = (det.sPEres / 4.) * sqrt(fabs(spike));
      double ns = spike + sw * D.z_spike;
      for (int guard = 0; ns <= 0. && guard < 100000; ++guard) ns = spike + sw * normal(g);
      out[6] = ns;
      out[7] = ns / pre.fitSm * P.s1_center;
    }
  }

  // rand_uniform() > prob && prob < 1: the draw is irrelevant when prob >= 1
  if (prob < 1. && u32d(D.u_coin) > prob) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      double v = out[i];
      if (nearly_equal(v, 0.)) v = NB_PHE_MIN;
      out[i] = -1. * fabs(v);
    }
  }
}

// NESTcalc::GetS2, S2CalculationMode::Full: NEST.cpp:1724-1776, 1975-2063.  out[9] = ionization
NB_D void get_s2(Stream& g, uint64_t ev, const RunParams& P, const PostDraws& D, int Ne, double tx, double ty, double tz,
                 double sx, double sy, double dt, double dfield, double* out) {
  const nb200_detector& det = P.det;
  const double elYield = P.rc.g2_params[0], ExtEff = P.rc.g2_params[1], SE = P.rc.g2_params[2],
               g2 = P.rc.g2_params[3], gasGap = P.rc.g2_params[4];
#pragma unroll
  for (int i = 0; i < 9; ++i) out[i] = 0.;
  if (dfield < NB_FIELD_MIN || elYield <= 0. || ExtEff <= 0. || SE <= 0. || g2 <= 0. || gasGap <= 0.)
    return;
  double posDep = fit_s2(det.FitS2, tx, ty) / P.s2_center;
  double posDepSm = fit_s2(det.FitS2, sx, sy) / P.s2_center;
  if (nearly_equal(det.g1_gas, 1.)) {
    posDep = 1.;
    posDepSm = 1.;
  }
  if (nearly_equal(det.g1_gas, 0.)) {
    posDep = 0.;
    posDepSm = 0.;
  }
  const double life = exp(-dt / det.eLife_us);
  long long Nee = binomial(g.rk, ev, STREAM_BIN_NEE, Ne, ExtEff * life);
  double mu = elYield * double(Nee);
  const double z0 = D.z_s2ph, z1 = D.z_s2area;  // one Box-Muller pair: photon number here, pulse area below
  double nraw = mu + sqrt(det.s2Fano * elYield * double(Nee)) * z0;
  double nphd = floor(((nraw < 0.) ? 0. : nraw) + 0.5);
  long long Nph = (long long)nphd;
  long long nHits = binomial(g.rk, ev, STREAM_BIN_S2HIT, Nph, det.g1_gas * posDep);
  // second photo-electrons: Binomial(nHits, P_dphe), fixed probability -> alias tables (nb_rng.cuh)
  const bool tab = P.dpe_alias != nullptr && nHits < NB_ALIAS_MAXN;
  long long Nphe = nHits;
  if (tab)
    Nphe += binomial_fixed_p(g.rk, ev, STREAM_BIN_S2DPE, int(nHits), P.dpe_alias);
  else  // very large S2 (or P_dphe of 0 / 1): the general sampler
    Nphe += binomial(g.rk, ev, STREAM_BIN_S2DPE + 8u, nHits, det.P_dphe);
  double m = div_z(double(Nphe), P.rc.NewMean);
  double pulseArea = m + det.sPEres * sqrt(m) * z1;
  pulseArea = (pulseArea < 0.) ? 0. : pulseArea;
  double sig;
  if (det.noiseQuadratic[1] != 0) {
    double a = det.noiseQuadratic[1] * (pulseArea * pulseArea), b = det.noiseLinear[1] * pulseArea;
    sig = sqrt(a * a + b * b);
  } else
    sig = det.noiseLinear[1] * pulseArea;
  if (sig != 0.) pulseArea = pulseArea + sig * D.z_s2noise;  // rand_gauss(pulseArea, sig, true)
  pulseArea = (pulseArea < 0.) ? 0. : pulseArea;
  double pulseAreaC = div_z(div_z(pulseArea, life), posDepSm);
  double Nphd = div_z(pulseArea, 1. + det.P_dphe);
  double NphdC = div_z(pulseAreaC, 1. + det.P_dphe);
  out[0] = double(Nee);
  out[1] = double(Nph);
  out[2] = double(nHits);
  out[3] = double(Nphe);
  if (det.s2_thr >= 0) {
    out[4] = pulseArea;
    out[5] = pulseAreaC;
    out[6] = Nphd;
    out[7] = NphdC;
  } else {  // bottom-array S2 (:2010-2049); the S2b draw only matters on this branch
    double tba = fit_tba(det.FitTBA, tx, ty, tz);
    double S2b = gauss(g, tba * pulseArea, sqrt(tba * pulseArea * (1. - tba)), true);
    double S2bc = S2b / life / posDepSm;
    out[4] = S2b;
    out[5] = S2bc;
    out[6] = S2b / (1. + det.P_dphe);
    out[7] = S2bc / (1. + det.P_dphe);
  }
  if (pulseArea < fabs(det.s2_thr)) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      double v = out[i];
      if (nearly_equal(v, 0.)) v = NB_PHE_MIN;
      out[i] = -1. * fabs(v);
    }
  }
  out[8] = g2;
}

// ---- energy sources: src/TestSpectra.cpp, execNEST.cpp:681-790 -----------------------------

// The reference's spectra are box rejections (RandomGen::VonNeumann RandomGen.cpp:140-157):
// x ~ U(xMin, xMax), y ~ U(0, yMax), keep the first pair with y <= f(x).  Candidate k of an event is
// Philox block k of its stream, so WHICH candidate an event ends up with does not depend on who
// evaluates it: when a full warp is sampling (the chain kernels), the lanes that already hold an
// accepted energy evaluate further candidates of the lanes that do not -- floor(32 / needy) candidates
// per needy lane and pass, the first accepted one (in stream order) wins.  A warp then finishes in
// ~3 passes instead of the ~5 sequential tries of its unluckiest lane, and every event gets
// bit for bit the energy the sequential loop gives (NB_COOP_REJECT=0 builds that loop;
// tools/cmp_variants.py compares the two).
#ifndef NB_COOP_REJECT
#define NB_COOP_REJECT 1
#endif
template <class F>
NB_D double box_reject(Stream& g, double xMin, double xMax, double yMax, const F& f) {
#if NB_COOP_REJECT
  const unsigned grp = __activemask();
  if (grp == 0xffffffffu && __all_sync(grp, !g.have)) {
    const int lane = int(threadIdx.x & 31u);
    const unsigned lt = (1u << lane) - 1u;
    uint32_t kdone = g.blk;  // next untried block of this lane's stream
    double x = xMin;
    bool need = true;
    {
      const u128 r = philox4x32_10_rk(g.e0, g.e1, kdone, g.sid, g.rk);
      const double x0 = xMin + (xMax - xMin) * u53(r.x, r.y);
      const double y0 = yMax * u53(r.z, r.w);
      ++kdone;
      if (!(y0 > f(x0))) {
        x = x0;
        need = false;
      }
    }
    unsigned nm = __ballot_sync(grp, need);
    for (int pass = 0; nm && pass < 100000; ++pass) {
      const int nneed = __popc(nm);
      const int per = 32 / nneed;
      const int j = lane / per, t = lane - j * per;
      const bool serve = j < nneed;
      const int target = serve ? nth_set_bit(nm, j) : lane;
      const uint32_t te0 = __shfl_sync(grp, g.e0, target), te1 = __shfl_sync(grp, g.e1, target);
      const uint32_t tk = __shfl_sync(grp, kdone, target);
      bool ok = false;
      double xc = 0.;
      if (serve) {
        const u128 r = philox4x32_10_rk(te0, te1, tk + uint32_t(t), g.sid, g.rk);
        xc = xMin + (xMax - xMin) * u53(r.x, r.y);
        const double y0 = yMax * u53(r.z, r.w);
        ok = !(y0 > f(xc));
      }
      const unsigned okm = __ballot_sync(grp, ok);
      int src = lane;
      bool got = false;
      if (need) {
        const int myj = __popc(nm & lt);
        const unsigned seg = (per == 32) ? okm : ((okm >> (myj * per)) & ((1u << per) - 1u));
        if (seg) {
          const int tt = __ffs(int(seg)) - 1;
          src = myj * per + tt;
          got = true;
          kdone += uint32_t(tt + 1);
        } else
          kdone += uint32_t(per);
      }
      const double xs = __shfl_sync(grp, xc, src);
      if (got) {
        x = xs;
        need = false;
      }
      nm = __ballot_sync(grp, need);
    }
    g.blk = kdone;
    g.have = false;
    return x;
  }
#endif
  for (int it = 0; it < 100000; ++it) {
    const double x0 = xMin + (xMax - xMin) * g.uniform();
    const double y0 = yMax * g.uniform();
    if (!(y0 > f(x0))) return x0;
  }
  return xMin;
}

// TestSpectra::CH3T_spectrum TestSpectra.cpp:31-58 (yMax 1.1e7)
struct Ch3tShape {
  NB_D double operator()(double x0) const {
    const double m_e = 510.9989461, aa = 0.0072973525664, ZZ = 2., qValue = 18.5898;
    double B = sqrt(x0 * x0 + 2. * x0 * m_e) / (x0 + m_e);
    double x = (2. * NB_PI * ZZ * aa) * (x0 + m_e) / sqrt(x0 * x0 + 2. * x0 * m_e);
    return (sqrt(2. * x0 * m_e) * (x0 + m_e) * (qValue - x0) * (qValue - x0) * x * (1. / (1. - exp(-x))) *
            (1.002037 - 0.001427 * (B)));
  }
};
NB_D double ch3t_spectrum(Stream& g, double xMin, double xMax) {
  const double qValue = 18.5898, yMax = 1.1e7;
  if (xMax > qValue) xMax = qValue;
  if (xMin < 0.) xMin = 0.;
  return box_reject(g, xMin, xMax, yMax, Ch3tShape());
}
// TestSpectra::B8_spectrum TestSpectra.cpp:110-128 (range forced to [0,5] keV)
struct B8Shape {
  NB_D double operator()(double x0) const {
    const double m1 = -2.6455e-10, m2 = 5.0843;
    double f = 3.4155 - 1.2184 * x0 + 0.32849 * (x0 * x0) - 0.12441 * (x0 * x0 * x0) + m1 * exp(m2 * x0);
    return exp(2.302585092994045684 * f);  // 10^f; a rejection bound, 1e-15 is ample
  }
};
NB_D double b8_spectrum(Stream& g) { return box_reject(g, 0., 5., pow(10., 3.4155), B8Shape()); }
// TestSpectra::DD_spectrum TestSpectra.cpp:192-211
struct DdShape {
  double expFall, peakFrac, peakMu, peakSig, peakSkew;
  NB_D double operator()(double x0) const {
    double d = (x0 - peakMu) / peakSig;
    return exp(-x0 / expFall) + peakFrac * exp(-(d * d)) * (erf(peakSkew * (x0 - peakMu) / peakSig) + 1);
  }
};
NB_D double dd_spectrum(Stream& g, double xMin, double xMax, double expFall, double peakFrac,
                        double peakMu, double peakSig, double peakSkew) {
  if (xMax > 80.) xMax = 80.;
  if (xMin < 0.000) xMin = 0.000;
  double yMax = 1.;
  if (peakMu < xMax) yMax += peakFrac;
  return box_reject(g, xMin, xMax, yMax, DdShape{expFall, peakFrac, peakMu, peakSig, peakSkew});
}
// TestSpectra::PowLawFit_spectrum TestSpectra.cpp:151-167 (AmBe, expo -0.95351)
NB_D double powlaw_spectrum(Stream& g, double xMin, double xMax, double expo) {
  double yMax, yMin;
  expo += 1.;
  if (expo < 0.) {
    if (xMin <= 0.) xMin = 0.1;
    yMax = pow(xMin, expo);
    yMin = pow(xMax, expo);
  } else {
    yMax = pow(xMax, expo);
    yMin = pow(xMin, expo);
  }
  double xTry = pow(yMin + (yMax - yMin) * g.uniform(), 1. / expo);
  if (xTry < xMin) xTry = xMin;
  if (xTry > xMax) xTry = xMax;
  return xTry;
}
// The less common spectra, one out-of-line subroutine (the chain kernels are instruction-fetch
// sensitive): C14_spectrum TestSpectra.cpp:60-108, Cf_spectrum :169-190, ppSolar_spectrum
// :213-227, atmNu_spectrum :229-245.  Each is the reference's box rejection
// (RandomGen::VonNeumann RandomGen.cpp:140-157): x ~ U(xMin,xMax), y ~ U(0,yMax), keep if y <= f(x).
struct RareShape {
  int kind;
  NB_D double operator()(double x0) const {
    double f;
    if (kind == NB200_SRC_C14) {
      const double m_e = 510.9989461, aa = 0.0072973525664, ZZ = 7., V0 = 0.495, qValue = 156.;
      double Ee = x0 + m_e;
      double pe = sqrt(Ee * Ee - m_e * m_e);
      double dNdE_phasespace = pe * Ee * (qValue - x0) * (qValue - x0);
      double Ee_screen = Ee - V0;
      double W_screen = (Ee_screen) / m_e;
      double p_screen = sqrt(W_screen * W_screen - 1);
      double WW = (Ee) / m_e;
      double pp = sqrt(WW * WW - 1);
      double G_screen = (Ee_screen) / (m_e);
      double B_screen = sqrt((G_screen * G_screen - 1) / (G_screen * G_screen));
      double x_screen = (2 * NB_PI * ZZ * aa) / B_screen;
      double F_nr_screen = W_screen * p_screen / (WW * pp) * x_screen * (1 / (1 - exp(-x_screen)));
      double F_bb_screen = F_nr_screen * pow(W_screen * W_screen * (1 + 4 * (aa * ZZ) * (aa * ZZ)) - 1,
                                             sqrt(1 - aa * aa * ZZ * ZZ) - 1);
      f = dNdE_phasespace * F_bb_screen;
    } else if (kind == NB200_SRC_CF) {
      const double l = log10(x0);
      f = 3.7488 - 0.77942 * l + 1.30300 * (l * l) - 2.75280 * (l * l * l) + 1.57310 * ((l * l) * (l * l)) -
          0.30072 * ((l * l) * (l * l) * l);
      f = exp(2.302585092994045684 * f);  // 10^f
      f *= 1.9929 - .033214 * x0 + .00032857 * (x0 * x0) - 1.000e-6 * (x0 * x0 * x0);
    } else if (kind == NB200_SRC_PPSOLAR) {
      f = 9.2759e-6 - 3.5556e-8 * x0 - 5.4608e-12 * x0 * x0;
    } else {
      f = (1. + 0.0041482 * x0 + 0.00079972 * x0 * x0 - 1.0201e-5 * x0 * x0 * x0) * exp(-x0 / 12.355);
    }
    return f;
  }
};
__device__ __noinline__ double rare_spectrum(Stream& g, int kind, double xMin, double xMax) {
  double yMax = 1.;
  if (kind == NB200_SRC_C14) {
    if (xMax > 156.) xMax = 156.;
    if (xMin < 0.) xMin = 0.;
    yMax = 2.5e9;
  } else if (kind == NB200_SRC_CF) {
    if (xMax > 200.) xMax = 200.;
    if (xMin < DBL_MIN) xMin = DBL_MIN;
    yMax = 2. * pow(10., 3.7488);
  } else if (kind == NB200_SRC_PPSOLAR) {
    if (xMax > 250.) xMax = 250.;
    if (xMin < 0.00) xMin = 0.00;
    yMax = 0.000594;
  } else {
    if (xMax > 85.) xMax = 85.;
    if (xMin < 0.0) xMin = 0.0;
  }
  return box_reject(g, xMin, xMax, yMax, RareShape{kind});
}
NB_D int xe_isotope_index(int A) {
  switch (A) {
    case 124: return 0;
    case 126: return 1;
    case 128: return 2;
    case 129: return 3;
    case 130: return 4;
    case 131: return 5;
    case 132: return 6;
    case 134: return 7;
    default: return 8;
  }
}
// TestSpectra::WIMP_spectrum TestSpectra.cpp:456-499.  WIMP_dRate(0, m, day) depends only on
// the random isotope it draws (:278), so its 9 possible values are tabulated on the host
// (P.wimp_dr0) and every call of the reference becomes an isotope draw + lookup.
//
// The reference's nested loops are one flat sequence of ATTEMPTS: a point (x, y), two fresh
// isotope draws a, b for the triangle pre-cut y <= -a/xMax x + b (:465-475), and -- only if the
// point survives it -- the table lookup, the Von Neumann test y <= f(x) and one tick of the
// 100-try counter (:476-495).  Stream layout: block 0 = the isotope of yMax; attempt k = blocks
// 1+2k (x, y) and 2+2k (a, b).  With that layout an attempt can be evaluated by any lane, and a
// full warp shares the work as box_reject does: the spectrum is steep (a few per cent of the box), a
// lane needs ~10 attempts and a warp would wait for its unluckiest lane's ~30.
NB_D int xe_isotope_of(double u01) {
  double iso = u01 * 100.;
  if (iso <= 0.090) return 124;
  if (iso <= 0.180) return 126;
  if (iso <= 2.100) return 128;
  if (iso <= 28.54) return 129;
  if (iso <= 32.62) return 130;
  if (iso <= 53.80) return 131;
  if (iso <= 80.69) return 132;
  if (iso <= 91.13) return 134;
  return 136;
}
struct WimpAttempt {
  double x0;
  bool passed;  // survived the triangle pre-cut: the Von Neumann test ran (one tick of the counter)
  bool acc;     // ... and accepted
  bool odd;     // no table interval holds x0: the reference would reuse the previous f (:458) -- replay
};
NB_D WimpAttempt wimp_attempt(const RunParams& P, uint32_t e0, uint32_t e1, uint32_t sid, const uint32_t* rk,
                              uint32_t k, double yMax) {
  const nb200_source& S = P.src;
  const double xMax = S.wimp_xMax, step = 1. / S.wimp_divisor;
  const u128 r = philox4x32_10_rk(e0, e1, 1u + 2u * k, sid, rk);
  const u128 q = philox4x32_10_rk(e0, e1, 2u + 2u * k, sid, rk);
  WimpAttempt w;
  w.x0 = xMax * u53(r.x, r.y);
  const double y0 = yMax * u53(r.z, r.w);
  const double a = P.wimp_dr0[xe_isotope_index(xe_isotope_of(u53(q.x, q.y)))];
  const double b = P.wimp_dr0[xe_isotope_index(xe_isotope_of(u53(q.z, q.w)))];
  w.passed = !(y0 > (-a / xMax * w.x0 + b));
  w.acc = false;
  w.odd = false;
  if (w.passed) {
    const int idx = int(w.x0 * S.wimp_divisor);
    const double xl = double(idx) * step;
    if (idx >= 0 && idx < 100 && xl < xMax && w.x0 > xl && w.x0 < xl + step)
      w.acc = !(y0 > S.wimp_base[idx] * exp(-S.wimp_exponent[idx] * w.x0));
    else
      w.odd = true;
  }
  return w;
}
// the literal sequential algorithm (persisting f), from attempt 0
NB_D double wimp_sequential(const RunParams& P, const Stream& g, double yMax, uint32_t& attempts) {
  const nb200_source& S = P.src;
  const double xMax = S.wimp_xMax, step = 1. / S.wimp_divisor;
  int count = 0;
  double f = 0.00;  // FuncValue persists across tries (:458)
  for (uint32_t k = 0; k < 1000000u; ++k) {
    const u128 r = philox4x32_10_rk(g.e0, g.e1, 1u + 2u * k, g.sid, g.rk);
    const u128 q = philox4x32_10_rk(g.e0, g.e1, 2u + 2u * k, g.sid, g.rk);
    const double x0 = xMax * u53(r.x, r.y), y0 = yMax * u53(r.z, r.w);
    const double a = P.wimp_dr0[xe_isotope_index(xe_isotope_of(u53(q.x, q.y)))];
    const double b = P.wimp_dr0[xe_isotope_index(xe_isotope_of(u53(q.z, q.w)))];
    attempts = k + 1u;
    if (y0 > (-a / xMax * x0 + b)) continue;  // triangle pre-cut (:465-475)
    const int idx = int(x0 * S.wimp_divisor);  // open interval (x, x + 1/divisor), x < xMax (:476-487)
    const double xl = double(idx) * step;
    if (idx >= 0 && idx < 100 && xl < xMax && x0 > xl && x0 < xl + step)
      f = S.wimp_base[idx] * exp(-S.wimp_exponent[idx] * x0);
    const bool rejected = y0 > f;           // RandomGen::VonNeumann RandomGen.cpp:140-157
    if (++count >= 100) return 0.;           // :491-495, taken even when the 100th try was accepted
    if (!rejected) return x0;
  }
  return 0.;
}
NB_D double wimp_spectrum(Stream& g, const RunParams& P) {
  // requires a fresh stream (the first draw of the event)
  const u128 r0 = philox4x32_10_rk(g.e0, g.e1, 0u, g.sid, g.rk);
  const double yMax = P.wimp_dr0[xe_isotope_index(xe_isotope_of(u53(r0.x, r0.y)))];
  uint32_t attempts = 0;
  double x = 0.;
#if NB_COOP_REJECT
  const unsigned grp = __activemask();
  if (grp == 0xffffffffu) {
    const int lane = int(threadIdx.x & 31u);
    const unsigned lt = (1u << lane) - 1u;
    int count = 0;
    bool need = true, replay = false;
    {
      const WimpAttempt w = wimp_attempt(P, g.e0, g.e1, g.sid, g.rk, 0u, yMax);
      attempts = 1u;
      if (w.odd)
        replay = true, need = false;
      else if (w.passed) {
        count = 1;
        if (w.acc) x = w.x0, need = false;
      }
    }
    unsigned nm = __ballot_sync(grp, need);
    for (int pass = 0; nm && pass < 100000; ++pass) {
      const int nneed = __popc(nm);
      const int per = 32 / nneed;
      const int j = lane / per, t = lane - j * per;
      const bool serve = j < nneed;
      const int target = serve ? nth_set_bit(nm, j) : lane;
      const uint32_t te0 = __shfl_sync(grp, g.e0, target), te1 = __shfl_sync(grp, g.e1, target);
      const uint32_t tk = __shfl_sync(grp, attempts, target);
      const double tyMax = __shfl_sync(grp, yMax, target);
      WimpAttempt w{0., false, false, false};
      if (serve) w = wimp_attempt(P, te0, te1, g.sid, g.rk, tk + uint32_t(t), tyMax);
      const unsigned pm = __ballot_sync(grp, w.passed), am = __ballot_sync(grp, w.acc), om = __ballot_sync(grp, w.odd);
      int src = lane;
      bool got = false;
      if (need) {
        const int myj = __popc(nm & lt);
        const unsigned segmask = (per == 32) ? 0xffffffffu : ((1u << per) - 1u);
        const unsigned ps = (pm >> (myj * per)) & segmask, as = (am >> (myj * per)) & segmask,
                       os = (om >> (myj * per)) & segmask;
        const int t_acc = as ? __ffs(int(as)) - 1 : 64;
        const int t_odd = os ? __ffs(int(os)) - 1 : 64;
        const int to100 = 100 - count;  // >= 1
        const int t_bail = (__popc(ps) >= to100) ? nth_set_bit(ps, to100 - 1) : 64;
        const int t_first = min(t_acc, min(t_odd, t_bail));
        if (t_first == 64) {  // nothing decisive in this batch
          count += __popc(ps);
          attempts += uint32_t(per);
        } else if (t_odd == t_first) {
          replay = true;
          need = false;
        } else if (t_bail == t_first) {  // the counter runs out first (bail-out wins over acceptance)
          x = 0.;
          attempts += uint32_t(t_bail + 1);
          need = false;
        } else {
          src = myj * per + t_acc;
          got = true;
          attempts += uint32_t(t_acc + 1);
        }
      }
      const double xs = __shfl_sync(grp, w.x0, src);
      if (got) {
        x = xs;
        need = false;
      }
      nm = __ballot_sync(grp, need);
    }
    if (replay) x = wimp_sequential(P, g, yMax, attempts);
  } else
#endif
    x = wimp_sequential(P, g, yMax, attempts);
  g.blk = 1u + 2u * attempts;
  g.have = false;
  return x;
}

// energy per event: execNEST.cpp:681-790
NB_D double sample_energy(Stream& g, const RunParams& P, uint64_t idx) {
  const nb200_source& S = P.src;
  double keV;
  switch (S.kind) {
    case NB200_SRC_MONO: keV = S.eMin; break;
    case NB200_SRC_LIST: return S.E[idx];
    case NB200_SRC_CH3T: keV = ch3t_spectrum(g, S.eMin, S.eMax); break;
    case NB200_SRC_B8: keV = b8_spectrum(g); break;
    case NB200_SRC_DD:
      keV = dd_spectrum(g, S.eMin, S.eMax, S.par[0], S.par[1], S.par[2], S.par[3], S.par[4]);
      break;
    case NB200_SRC_AMBE: keV = powlaw_spectrum(g, S.eMin, S.eMax, S.par[0]); break;
    case NB200_SRC_WIMP: return wimp_spectrum(g, P);
    case NB200_SRC_C14:
    case NB200_SRC_CF: keV = rare_spectrum(g, S.kind, S.eMin, S.eMax); break;
    case NB200_SRC_PPSOLAR:
    case NB200_SRC_ATMNU: return rare_spectrum(g, S.kind, S.eMin, S.eMax);  // not clamped (:786-787)
    case NB200_SRC_GAUSS: keV = zero_trunc_gauss(g, S.eMin, S.eMax); break;
    case NB200_SRC_EXP: keV = exponential(g, -S.eMax * log(2.), 0., S.eMin); break;
    default: keV = S.eMin + (S.eMax - S.eMin) * g.uniform(); break;
  }
  if (S.kind != NB200_SRC_B8 && S.eMax > 0. && S.eMin < S.eMax) {  // :786-790
    if (keV > S.eMax) keV = S.eMax;
    if (keV < S.eMin) keV = S.eMin;
  }
  return keV;
}

// signal selection execNEST.cpp:1121-1158
NB_D void select_signals(const nb200_analysis& A, const double* scint, const double* scint2,
                         double& signal1, double& signal2) {
  if (A.usePD == 0 && fabs(scint[3]) > A.minS1 && scint[3] < A.maxS1)
    signal1 = scint[3];
  else if (A.usePD == 1 && fabs(scint[5]) > A.minS1 && scint[5] < A.maxS1)
    signal1 = scint[5];
  else if ((A.usePD >= 2 && fabs(scint[7]) > A.minS1 && scint[7] < A.maxS1) || A.maxS1 >= 998.5)
    signal1 = scint[7];
  else
    signal1 = -999.;
  if (A.usePD == 0 && fabs(scint2[5]) > A.minS2 && scint2[5] < A.maxS2)
    signal2 = scint2[5];
  else if (A.usePD >= 1 && fabs(scint2[7]) > A.minS2 && scint2[7] < A.maxS2)
    signal2 = scint2[7];
  else
    signal2 = -999.;
}

// energy reconstruction execNEST.cpp:1170-1250; returns signalE, keV updated in place
NB_D double reconstruct_energy(const RunParams& P, const Yields& y, const double* scint,
                               const double* scint2, double signal1, double signal2, double field,
                               double& keV, double& keVee) {
  const nb200_detector& det = P.det;
  keVee = 0.;  // this event's term of the CLI's electron-equivalent energy sum, execNEST.cpp:1239-1240
  const nb200_analysis& A = P.ana;
  const double g1 = det.g1, g2 = fabs(P.rc.g2_params[3]);
  if (!A.MCtruthE) {
    double MultFact = 1., eff = det.sPEeff;
    if (A.s1mode != NB200_S1_PARAMETRIC) {
      if (det.sPEthr >= 0. && det.sPEres > 0. && eff > 0.) {
        MultFact = P.recon_mult;
        if (A.usePD == 2 && scint[7] < NB_SPIKES_MAXM)
          MultFact = 1.04;
        else
          MultFact = 1. + MultFact * det.P_dphe;
        if (eff < 1.) eff += ((1. - eff) / (2. * double(det.numPMTs))) * scint[0];
        eff = smax(0., smin(eff, 1.));
        MultFact /= eff;
      } else
        MultFact = 1.;
    }
    double Nph, Ne;
    if (A.usePD == 0)
      Nph = fabs(scint[3]) / (g1 * (1. + det.P_dphe));
    else if (A.usePD == 1)
      Nph = fabs(scint[5]) / g1;
    else
      Nph = fabs(scint[7]) / g1;
    if (A.usePD == 0)
      Ne = fabs(scint2[5]) / (g2 * (1. + det.P_dphe));
    else
      Ne = fabs(scint2[7]) / g2;
    Nph *= MultFact;
    if (signal1 <= 0.) Nph = 0.;
    if (signal2 <= 0.) Ne = 0.;
    if (det.coinLevel <= 0 && Nph <= NB_PHE_MIN) Nph = DBL_MIN;
    if (y.L > DBL_MIN && Nph > 0. && Ne > 0.) {
      if (nearly_equal(y.L, 1.))
        keV = (Nph + Ne) * P.rc.Wq_eV * 1e-3;
      else if (P.species <= NB200_Cf) {
        const double* n = P.model.NRYieldsParam;
        // x^y = exp(y ln x) with the bit-built logarithm: a reconstructed (smeared) energy, not one of
        // the deterministic yields held to 1e-12 -- its error (~|y ln x| 1e-16) is far below either
        const double inv = 1. / n[1];
        keV = exp(inv * log_pos((Ne + Nph) / n[0]));
        Ne *= 1. - 1. / pow_small(1. + pow_small((keV / n[5]), n[6]), n[10]);
        Nph *= 1. - 1. / pow_small(1. + pow_small((keV / n[7]), n[8]), n[11]);
        keV = exp(inv * log_pos((Ne + Nph) / n[0]));
      } else {
        const double logden = log10(P.rc.rho);
        const double Wq_eV_alpha =
            28.259 + 25.667 * logden - 33.611 * pow(logden, 2.) - 123.73 * pow(logden, 3.) -
            136.47 * pow(logden, 4.) - 74.194 * pow(logden, 5.) - 20.276 * pow(logden, 6.) -
            2.2352 * pow(logden, 7.);
        keV = (Nph + Ne) * Wq_eV_alpha * 1e-3 / y.L;
      }
      keVee = (Nph + Ne) * P.rc.Wq_eV * 1e-3;
    } else
      keV = 0.;
  }
  if ((signal1 <= 0. || signal2 <= 0.) && field >= NB_FIELD_MIN && det.coinLevel != 0) return 0.;
  return keV;
}

// GetBand bin search + value, execNEST.cpp:1508-1580 (resol == false).  Returns the bin or -1;
// accepted/value as the reference pushes them (0 when log10 falls outside (logMin, logMax)).
NB_D int band_bin(const RunParams& P, double signal1, double signal2, bool& accepted, double& xval,
                  double& yval) {
  const nb200_analysis& A = P.ana;
  const int nFold = P.det.coinLevel;
  const double X = (A.useS2 == 2) ? signal2 : signal1;
  const double O = (A.useS2 == 2) ? signal1 : signal2;
  const double binWidth = P.band_width, border = P.band_border;
  const double aX = fabs(X);
  int bin = -1;
  if (!nFold)
    bin = 0;
  else {
    int j0 = int(floor((aX - border) / binWidth)) - 1;
    if (j0 < 0) j0 = 0;
    for (int j = j0; j < j0 + 3 && j < A.numBins; ++j) {
      double s1c = border + binWidth / 2. + double(j) * binWidth;
      if (aX > (s1c - binWidth / 2.) && aX <= (s1c + binWidth / 2.)) {
        bin = j;
        break;
      }
    }
  }
  if (bin < 0) return -1;
  const double ReqS1 = nFold == 0 ? -DBL_MAX : 0.;
  accepted = (X >= ReqS1 && O >= 0.);
  xval = X;
  yval = 0.;
  if (accepted) {
    double v;
    if (A.useS2 == 0)
      v = log10(O / X);
    else if (A.useS2 == 1)
      v = log10(O);
    else
      v = log10(X / O);
    if (X != 0. && O != 0. && v > A.logMin && v < A.logMax) yval = v;
  }
  return bin;
}

#endif  // __CUDACC__
}  // namespace nb
