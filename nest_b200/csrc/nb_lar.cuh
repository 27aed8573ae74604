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

// NR: LArNEST.cpp:86-118, parameters LArParameters.hh:39-47
NB_D LArYields lar_nr(double energy, double efield, const LArFieldTerms& t) {
  const double alpha = 11.10, beta = 0.087, epsilon = 2.998, zeta = 0.3, eta = 2.94;
  const double lnE = log(energy);
  const double Eb = lar_pow(lnE, beta);
  double Total = alpha * Eb;
  const double rs = 1.0 / sqrt(energy + epsilon);
  double Ye = t.nr_fd * rs * (1.0 - 1.0 / (1.0 + lar_pow(lnE - log(zeta), eta)));
  if (Ye < 0.0) Ye = 0.0;
  double Yph = alpha * Eb - t.nr_fd * rs;
  if (Yph < 0.0) Yph = 0.0;
  return lar_recombination(Total, Ye, Yph, energy, efield);
}

// ER: LArNEST.cpp:120-170, parameters LArParameters.hh:48-94
NB_D LArYields lar_er(double energy, double efield, const LArFieldTerms& t) {
  double Total = (energy * 1.0e3 / NB_LAR_WQ);
  const double a = t.er_a, b = t.er_b, g = t.er_g, db = t.er_db;
  const double p1 = 1.0, p2 = 10.304, p3 = 13.0654, p4 = 0.10535, p5 = 0.7, dl = 15.7489, let = -2.07763;
  double Ye = a * b + (g - a * b) / lar_pow(log(p1 + p2 * lar_pow(log(energy + 0.5), p3)), p4) +
              dl / (p5 + db * lar_pow(log(energy), let));
  if (Ye < 0.0) Ye = 0.0;
  double Yph = Total / energy - Ye;
  if (Yph < 0.0) Yph = 0.0;
  return lar_recombination(Total, Ye, Yph, energy, efield);
}

// Alpha: LArNEST.cpp:172-212, parameters LArParameters.hh:96-128
NB_D LArYields lar_alpha(double energy, double efield, const LArFieldTerms& t) {
  double Total = (t.al_Ye + t.al_Yph) * energy;
  return lar_recombination(Total, t.al_Ye, t.al_Yph, energy, efield);
}

// LArNEST::GetYields LArNEST.cpp:35-50 (species NR=0, ER=1, Alpha=2); t: this thread's field-term memo
NB_D LArYields lar_yields(int species, double energy, double efield, double density, LArFieldTerms& t) {
  if (!(t.field == efield && t.density == density)) lar_field_terms(species, efield, density, t);
  if (species == 0) return lar_nr(energy, efield, t);
  if (species == 1) return lar_er(energy, efield, t);
  return lar_alpha(energy, efield, t);
}

// GetDefaultFluctuations LArNEST.cpp:363-406.  out = {Nph, Ne, Nex, Nion fluctuated}
NB_D void lar_fluctuations(Stream& g, uint64_t ev, const LArYields& y, double* out) {
  const double Fano = 0.1115;
  double NexOverNion = y.Nex / y.Nion;
  double p_r = (y.Nph - y.Nex) / y.Nion;
  double alf = 1. / (1. + NexOverNion);
  double Nq = 0, Nex = 0, Nion = 0;
  double Nq_mean = y.Ne + y.Nph;
  if (y.Lindhard == 1.) {  // exact comparison, as the reference (:374)
    Nq = floor(gauss(g, Nq_mean, sqrt(Fano * Nq_mean), true) + 0.5);
    if (Nq < 0.) Nq = 0.;
    Nion = double(binomial(g.rk, ev, STREAM_BIN_LAR, (long long)Nq, alf));
    if (Nion > Nq) Nion = Nq;
    Nex = Nq - Nion;
  } else {
    double z0, z1;
    normal_pair(g, z0, z1);
    double d = Nq_mean * alf + sqrt(Fano * Nq_mean * alf) * z0;
    Nion = floor(((d < 0.) ? 0. : d) + 0.5);
    d = Nq_mean * alf * NexOverNion + sqrt(Fano * Nq_mean * alf * NexOverNion) * z1;
    Nex = floor(((d < 0.) ? 0. : d) + 0.5);
    if (Nex < 0.) Nex = 0.;
    if (Nion < 0.) Nion = 0.;
    Nq = Nion + Nex;
  }
  out[0] = Nex + p_r * Nion;
  out[1] = (1. - p_r) * Nion;
  out[2] = Nex;
  out[3] = Nion;
}

#endif
}  // namespace nb
