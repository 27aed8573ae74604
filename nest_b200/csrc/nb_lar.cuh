// nb_lar.cuh -- liquid-argon mean yields and default fluctuations (device).
//
// Restates LArNEST::GetNRYields / GetERYields / GetAlphaYields / GetRecombinationYields /
// GetDefaultFluctuations (reference src/LArNEST.cpp:62-212, 363-406) with the default
// parameter sets of include/NEST/LArParameters.hh:39-128 and LArNEST.hh:432-436
// (W_q = 19.5 eV, Nex/Nion = 0.21).  Compiled with -fmad=false: the exact `Lindhard == 1.`
// comparison of :374 depends on each IEEE operation being rounded separately.
#pragma once
#include "nb_rng.cuh"

namespace nb {
#ifdef __CUDACC__

struct LArYields {  // LArYieldResult LArNEST.hh:19-29
  double TotalYield, QuantaYield, LightYield, Nph, Ne, Nex, Nion, Lindhard, ElectricField;
};

#define NB_LAR_WQ 19.5
#define NB_LAR_NEXONI 0.21

// GetRecombinationYields LArNEST.cpp:62-84
NB_D LArYields lar_recombination(double Total, double Ye, double Yph, double energy, double efield) {
  double Ne = Ye * energy;
  double Nph = Yph * energy;
  double Nq = Ne + Nph;
  double Nion = Nq / (1 + NB_LAR_NEXONI);
  double Nex = Nq - Nion;
  double Lindhard = (Total / energy) * NB_LAR_WQ * 1e-3;
  return LArYields{Total, Ye, Yph, Nph, Ne, Nex, Nion, Lindhard, efield};
}

// The factors of the three yield models that depend on the field (and density) only.  LArNEST's benchmarks sweep a
// few fields over many energies (examples/LArNEST/LArNESTMeanYieldsBenchmarks.cpp:17-94), so a thread keeps the
// factors of the last field it saw and re-evaluates them -- with the same libm calls as before, bit for bit -- only
// when the field changes: ER goes from eight pow per event to three, alpha from five to none.
struct LArFieldTerms {
  double field, density;  // what the terms below were evaluated for (NaN: nothing yet)
  double nr_fd;           // NR: 1 / (gamma F^delta)
  double er_a, er_b, er_g, er_db;
  double al_Ye, al_Yph;
};
NB_D void lar_field_terms(int species, double efield, double density, LArFieldTerms& t) {
  t.field = efield;
  t.density = density;
  if (species == 0) {
    const double gamma = 0.1, delta = -0.0932;
    t.nr_fd = 1.0 / (gamma * pow(efield, delta));
  } else if (species == 1) {
    t.er_a = 32.988 + (-552.988) / (17.2346 + pow(efield / (-4.7 + 0.025115 * exp(density / 0.265360653)), 0.242671));
    t.er_b = 0.778482 + 25.9 * pow(1.105 + pow(efield / 0.4, 4.55), -7.502);
    t.er_g = 0.659509 * (1000 / NB_LAR_WQ + 6.5 * (5.0 + (-0.5) / pow(efield / 1047.408, 0.01851)));
    t.er_db = 1052.264 + (14159350000 - 1652.264) / (-5.0 + pow(efield / 0.157933, 1.83894));
  } else {
    const double eA = 1.0 / 6200.0, eB = 64478398.7663, eC = 0.173553719, eD = 1.21, eE = 0.02852, eF = 0.01,
                 eG = 4.71598, eH = 7.72848, eI = -0.109802, eJ = 3.0;
    double fieldTerm = eF * pow((eG + pow(efield, eH)), eI);
    double Ye = eA * (eB - (eB * eC + (eB / eD) * (1 - (eE * log(1 + (eB / eD) * (fieldTerm / eJ)) /
                                                          ((eB / eD) * fieldTerm)))));
    if (Ye < 0.0) Ye = 0.0;
    const double pA = 1.16, pB = -0.012, pC = 1.0 / 6500.0, pD = 278037.250283, pE = 0.173553719, pF = 1.21,
                 pG = 2, pH = 0.653503, pI = 4.98483, pJ = 10.0822, pK = 1.2076, pL = -0.97977, pM = 3.0;
    double quench = 1.0 / (pA * pow(efield, pB));
    double ft = pH * pow((pI + pow((efield / pJ), pK)), pL);
    double Yph = quench * pC *
                 (pD * pE + (pD / pF) * (1 - (pG * log(1 + (pD / pF) * (ft / pM)) / ((pD / pF) * ft))));
    if (Yph < 0.0) Yph = 0.0;
    t.al_Ye = Ye;
    t.al_Yph = Yph;
  }
}

// x^y for x > 0 as exp(y ln x): the energy-dependent powers of the NR and ER models.  CUDA's exp and log are
// correctly rounded to ~1 ulp, so the result is within |y ln x| 2^-52 (< 4e-15 for every exponent used here) of
// libm's pow -- inside the 1e-12 the yields are held to (tests/test_gpu_lar.py) at a third of pow's cost.
NB_D double lar_pow(double lnx, double y) { return exp(y * lnx); }
// ln x: the bit-built logarithm (nb_rng.cuh log_pos: absolute error ~1e-16, a quarter of the instructions of CUDA's
// log) wherever it applies -- positive normal finite x, i.e. every energy of a sane list -- and libm's otherwise, so
// that zero, negative, subnormal and non-finite energies still give the reference's inf / NaN
NB_D double lar_log(double x) {
  const int hi = __double2hiint(x);
  if (hi >= 0x00100000 && hi < 0x7ff00000) return log_pos(x);
  return log(x);
}

// GetRecombinationYields for the models whose Lindhard factor never meets the exact comparison of :374 (NR ~ 0.03,
// alpha ~ 0.8): the two divisions become products with reciprocals (<= 1 ulp each: inside the 1e-12 gate)
NB_D LArYields lar_recombination_fast(double Total, double Ye, double Yph, double energy, double efield) {
  double Ne = Ye * energy;
  double Nph = Yph * energy;
  double Nq = Ne + Nph;
  double Nion = Nq * (1.0 / (1 + NB_LAR_NEXONI));
  double Nex = Nq - Nion;
  // (1 / energy overflows for subnormal energies, where the quotient itself does not)
  const double TE = fabs(energy) >= 1e-300 ? Total * (1.0 / energy) : Total / energy;
  double Lindhard = TE * NB_LAR_WQ * 1e-3;
  return LArYields{Total, Ye, Yph, Nph, Ne, Nex, Nion, Lindhard, efield};
}

// NR: LArNEST.cpp:86-118, parameters LArParameters.hh:39-47
NB_D LArYields lar_nr(double energy, double efield, const LArFieldTerms& t) {
  const double alpha = 11.10, beta = 0.087, epsilon = 2.998, eta = 2.94;
  const double ln_zeta = -1.2039728043259359926;  // ln 0.3
  const double lnE = lar_log(energy);
  const double Eb = lar_pow(lnE, beta);
  double Total = alpha * Eb;
  const double rs = rsqrt(energy + epsilon);
  double Ye = t.nr_fd * rs * (1.0 - 1.0 / (1.0 + lar_pow(lnE - ln_zeta, eta)));
  if (Ye < 0.0) Ye = 0.0;
  double Yph = alpha * Eb - t.nr_fd * rs;
  if (Yph < 0.0) Yph = 0.0;
  return lar_recombination_fast(Total, Ye, Yph, energy, efield);
}

// ER: LArNEST.cpp:120-170, parameters LArParameters.hh:48-94.  Total, Total / energy and the Lindhard factor keep the
// reference's IEEE operations one by one: `Lindhard == 1.` (:374) is decided by their last bits
NB_D LArYields lar_er(double energy, double efield, const LArFieldTerms& t) {
  double Total = (energy * 1.0e3 / NB_LAR_WQ);
  const double a = t.er_a, b = t.er_b, g = t.er_g, db = t.er_db;
  const double p1 = 1.0, p2 = 10.304, p3 = 13.0654, p4 = 0.10535, p5 = 0.7, dl = 15.7489, let = -2.07763;
  double Ye = a * b + (g - a * b) / lar_pow(lar_log(p1 + p2 * lar_pow(lar_log(energy + 0.5), p3)), p4) +
              dl / (p5 + db * lar_pow(lar_log(energy), let));
  if (Ye < 0.0) Ye = 0.0;
  const double TE = Total / energy;
  double Yph = TE - Ye;
  if (Yph < 0.0) Yph = 0.0;
  double Ne = Ye * energy;
  double Nph = Yph * energy;
  double Nq = Ne + Nph;
  double Nion = Nq * (1.0 / (1 + NB_LAR_NEXONI));
  double Nex = Nq - Nion;
  double Lindhard = TE * NB_LAR_WQ * 1e-3;
  return LArYields{Total, Ye, Yph, Nph, Ne, Nex, Nion, Lindhard, efield};
}

// Alpha: LArNEST.cpp:172-212, parameters LArParameters.hh:96-128
NB_D LArYields lar_alpha(double energy, double efield, const LArFieldTerms& t) {
  double Total = (t.al_Ye + t.al_Yph) * energy;
  return lar_recombination_fast(Total, t.al_Ye, t.al_Yph, energy, efield);
}

// LArNEST::GetYields LArNEST.cpp:35-50 (species NR=0, ER=1, Alpha=2); t: this thread's field-term memo
NB_D LArYields lar_yields(int species, double energy, double efield, double density, LArFieldTerms& t) {
  if (!(t.field == efield && t.density == density)) lar_field_terms(species, efield, density, t);
  if (species == 0) return lar_nr(energy, efield, t);
  if (species == 1) return lar_er(energy, efield, t);
  return lar_alpha(energy, efield, t);
}

// GetDefaultFluctuations LArNEST.cpp:363-406.  out = {Nph, Ne, Nex, Nion fluctuated}.
// Nex / Nion is 0.21 for every event (GetRecombinationYields: Nion = Nq / 1.21, Nex = Nq - Nion) up to the rounding
// of those two operations (2e-16 relative), so alf = 1 / (1 + Nex / Nion) = 1 / 1.21: the constants stand for the
// reference's per-event quotients -- the means and widths of the draws move by rounding errors, eight orders below
// anything a sample can resolve.  Box-Muller's pair becomes two ziggurat normals (the hit loop's sampler: exact, a
// table look-up each) taken from the event's one Philox block, which also carries the wedge test's spare word.
//
// The `Lindhard == 1.` branch ends in Binomial(Nq, alf), Nq in the tens of thousands: Hormann's BTRS.  For ER the
// branch is taken by 58 % of the events, pseudo-randomly (it hangs on the last bit of E * 1e3 / 19.5 / E), and a
// rejection loop in lockstep runs as long as its unluckiest lane: measured, the draw cost twice the rest of the
// kernel.  lar_fluct_head therefore does everything but that draw and reports it as pending (Nq_pending >= 0);
// k_lar queues the pending draws of a warp in shared memory and makes them 32 at a time, one candidate per lane and
// pass (lar_binomial_pass), re-queueing the rejected ones: every lane of every pass is a live candidate.  Candidate k
// of an event's draw is Philox block k of its STREAM_BIN_LAR stream whoever makes it, so the result is that of the
// in-line sampler (binomial, nb_rng.cuh) bit for bit.
#define NB_LAR_ALF (1. / (1. + NB_LAR_NEXONI))
NB_D void lar_fluct_store(double Nex, double Nion, double p_r, double* out) {
  out[0] = Nex + p_r * Nion;
  out[1] = (1. - p_r) * Nion;
  out[2] = Nex;
  out[3] = Nion;
}
// defer: the caller can take a pending BTRS draw (else it is made here).  Returns Nq when the draw is pending
// (out not written), -1 otherwise
NB_D double lar_fluct_head(const uint32_t* rk, uint64_t seed, uint64_t ev, const LArYields& y, bool defer, double& p_r,
                           double* out) {
  const double Fano = 0.1115;
  const double NexOverNion = NB_LAR_NEXONI;
  const double alf = NB_LAR_ALF;
  p_r = (y.Nph - y.Nex) / y.Nion;
  double Nex = 0, Nion = 0;
  double Nq_mean = y.Ne + y.Nph;
  const uint32_t e0 = uint32_t(ev), e1 = uint32_t(ev >> 32);
  const u128 r = philox4x32_10_rk(e0, e1, 0u, STREAM_LAR, rk);
  uint32_t A0, L0, A1, L1;
  pe_fields(r, 0, A0, L0);
  pe_fields(r, 1, A1, L1);
  double z0, z1;
  const bool f0 = zig_fast(A0, L0, g_zig_k, g_zig_w, z0);
  if (!f0) z0 = zig_finish(A0, L0, r.w, true, e0, e1, 0u, STREAM_ZIG, uint32_t(seed), uint32_t(seed >> 32));
  if (y.Lindhard == 1.) {  // exact comparison, as the reference (:374)
    const double d = Nq_mean + sqrt(Fano * Nq_mean) * z0;
    double Nq = floor(((d < 0.) ? 0. : d) + 0.5);
    if (!(Nq >= 0.)) Nq = 0.;  // (a NaN mean: no quanta)
    if (defer && Nq * (1. - alf) >= 10.) return Nq;  // BTRS regime: the caller's queue
    Nion = double(binomial(rk, ev, STREAM_BIN_LAR, (long long)Nq, alf));
    if (Nion > Nq) Nion = Nq;
    Nex = Nq - Nion;
  } else {
    const bool f1 = zig_fast(A1, L1, g_zig_k, g_zig_w, z1);
    if (!f1) z1 = zig_finish(A1, L1, 0u, false, e0, e1, 1u, STREAM_ZIG, uint32_t(seed), uint32_t(seed >> 32));
    const double mi = Nq_mean * alf, si = sqrt(Fano * mi);
    double d = mi + si * z0;
    Nion = floor(((d < 0.) ? 0. : d) + 0.5);
    d = mi * NexOverNion + si * 0.45825756949558400066 * z1;  // sqrt(0.21)
    Nex = floor(((d < 0.) ? 0. : d) + 0.5);
    if (Nex < 0.) Nex = 0.;
    if (Nion < 0.) Nion = 0.;
  }
  lar_fluct_store(Nex, Nion, p_r, out);
  return -1.;
}
// candidate `att` of Binomial(Nq, alf) for event ev, the statements of binomial() (alf > 1/2: drawn for 1 - alf and
// flipped): true when accepted
NB_D bool lar_binomial_try(const uint32_t* rk, uint64_t ev, uint32_t att, double Nq, double& Nion) {
  const double alf = NB_LAR_ALF;
  const double pp = 1. - alf, q = 1. - pp;
  Btrs B = btrs_setup(Nq, pp, q);
  B.logr = -1.5606477482646686;  // ln(pp / q) = ln 0.21: the same for every event (the setup's own logarithm is dead code then)
  const u128 w = philox4x32_10_rk(uint32_t(ev), uint32_t(ev >> 32), att, STREAM_BIN_LAR, rk);
  double kk;
  const bool ok = btrs_try(B, w, kk);
  Nion = double((long long)Nq - (long long)kk);
  return ok;
}
struct LArPending {
  double Nq, p_r;
  uint32_t off, att;  // event = block's first + off; candidates already rejected
};

#endif
}  // namespace nb
