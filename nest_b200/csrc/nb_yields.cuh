// nb_yields.cuh -- deterministic mean-yield models and the detector position maps.
//
// Device restatement of NESTcalc::GetYields and friends (reference src/NEST.cpp,
// lines cited per function).  These are the functions held to 1e-12 relative against
// the reference: the operation order of the reference expressions is kept, and the
// translation unit is compiled with -fmad=false so that no multiply-add is contracted
// (the reference is x86-64 without FMA contraction).
#pragma once
#include "nb_params.h"
#include "nb_rng.cuh"

namespace nb {

// physics constants: include/NEST/NEST.hh:81-113, src/NEST.cpp:7-19
#define NB_W_SCINT 8.5e-3
#define NB_NEST_AVO 6.0221409e+23
#define NB_ATOM_NUM 54.
#define NB_PHE_MIN 1e-6
#define NB_FIELD_MIN 1.
#define NB_DENSITY 2.90
#define NB_SPIKES_MAXM 120
#define NB_ZurichEXOW 1.1716263232
#define NB_ZurichEXOQ 1.08
#define NB_SQRT2 1.41421356237309504880
#define NB_INV_SQRT2PI 0.39894228040143267794
#define NB_PI 3.14159265358979323846

// std::max / std::min argument-order semantics (NaN handling differs from fmax/fmin)
NB_HD double smax(double a, double b) { return (a < b) ? b : a; }
NB_HD double smin(double a, double b) { return (b < a) ? b : a; }

// ValidityTests::nearlyEqual src/ValidityTests.cpp:6-21
NB_HD bool nearly_equal(double a, double b, double eps = 1e-9) {
  if (a == b) return true;
  double absA = fabs(a), absB = fabs(b), diff = fabs(a - b);
  if (a == 0 || b == 0 || (absA + absB < DBL_MIN)) return diff < (eps * DBL_MIN);
  const double sum = fmin(absA + absB, DBL_MAX);
  // diff / sum < eps, decided without the division unless the quotient is within a factor two of eps
  if (diff > 2. * eps * sum) return false;
  if (diff < 0.5 * eps * sum) return true;
  return diff / sum < eps;
}

// a / b for a numerator that is often exactly zero (an event without a signal) and a positive finite b: the
// quotient is then a itself, and the division is given a numerator of one so that it stays on its in-line
// path (a zero numerator sends the lane through the out-of-line special-case routine, ~60 instructions)
NB_HD double div_z(double a, double b) {
  const bool z = a == 0. && b > 0. && b < DBL_MAX;
  double num = z ? 1. : a;
#ifdef __CUDA_ARCH__
  asm volatile("" : "+d"(num));  // opaque: otherwise the select is folded away (the quotient of a z lane is unused) and a / b comes back
#endif
  const double q = num / b;
  return z ? a : q;
}

struct Yields {  // YieldResult NEST.hh:162-169
  double Nph, Ne, NexONi, L, field, dT;
};
struct Wvalue {
  double Wq_eV, alpha;
};

// NESTcalc::WorkFunction NEST.cpp:2829-2848 (xenon branch; ATOM_NUM is 54 at compile time)
NB_HD Wvalue work_function(double density, double molarMass, bool oldW13eV) {
  Wvalue w;
  w.alpha = 0.067366 + density * 0.039693;
  double eDensity = (density / molarMass) * NB_NEST_AVO * NB_ATOM_NUM;
  w.Wq_eV = 18.7263 - 1.01e-23 * eDensity;
  if (oldW13eV) w.Wq_eV *= NB_ZurichEXOW;
  return w;
}

// NESTcalc::YieldResultValidity NEST.cpp:1254-1271
NB_HD Yields validity(Yields r, double energy, double Wq_eV) {
  if (r.Nph > energy / NB_W_SCINT) r.Nph = energy / NB_W_SCINT;
  if (r.Ne > energy / NB_W_SCINT) r.Ne = energy / NB_W_SCINT;
  if (r.Nph < 0.) r.Nph = 0.;
  if (r.Ne < 0.) r.Ne = 0.;
  r.L = smax(0., smin(r.L, 1.));
  if (energy < 0.001 * Wq_eV / r.L) {
    r.Nph = 0.;
    r.Ne = 0.;
  }
  return r;
}

NB_HD int imin_(int a, int b) { return a < b ? a : b; }
NB_HD int imax_(int a, int b) { return a > b ? a : b; }

// integer powers: on the host exactly the reference's pow(x, n.) calls (the per-run constants are
// bit-identical to the reference's), on the device plain products (<= 1 ulp apart)
#ifdef __CUDA_ARCH__
#define NB_POW2(x) ((x) * (x))
#define NB_POW3(x) ((x) * (x) * (x))
#define NB_POW4(x) (((x) * (x)) * ((x) * (x)))
#define NB_POW5(x) (((x) * (x)) * ((x) * (x)) * (x))
#define NB_POW6(x) (((x) * (x) * (x)) * ((x) * (x) * (x)))
#define NB_POW7(x) ((((x) * (x)) * ((x) * (x))) * ((x) * (x) * (x)))
#else
#define NB_POW2(x) pow((x), 2.)
#define NB_POW3(x) pow((x), 3.)
#define NB_POW4(x) pow((x), 4.)
#define NB_POW5(x) pow((x), 5.)
#define NB_POW6(x) pow((x), 6.)
#define NB_POW7(x) pow((x), 7.)
#endif

// ---- position maps: VDetector::FitS1/FitS2/FitEF/FitTBA (VDetector.hh:131-148) -------
NB_HD double map_lut_rz(const nb200_map& m, double r, double z) {
  double fr = (m.n0 > 1) ? (r - m.c[0]) / (m.c[1] - m.c[0]) * double(m.n0 - 1) : 0.;
  double fz = (m.n1 > 1) ? (z - m.c[2]) / (m.c[3] - m.c[2]) * double(m.n1 - 1) : 0.;
  fr = fmin(fmax(fr, 0.), double(m.n0 - 1));
  fz = fmin(fmax(fz, 0.), double(m.n1 - 1));
  int i0 = imin_(int(fr), imax_(m.n0 - 2, 0)), j0 = imin_(int(fz), imax_(m.n1 - 2, 0));
  int i1 = imin_(i0 + 1, m.n0 - 1), j1 = imin_(j0 + 1, m.n1 - 1);
  double tr = fr - double(i0), tz = fz - double(j0);
  const double* t = m.lut;
  double v00 = t[size_t(i0) * m.n1 + j0], v01 = t[size_t(i0) * m.n1 + j1];
  double v10 = t[size_t(i1) * m.n1 + j0], v11 = t[size_t(i1) * m.n1 + j1];
  double a = v00 + (v01 - v00) * tz, b = v10 + (v11 - v10) * tz;
  return a + (b - a) * tr;
}

// FitS1: LUX_Run03.hh:90-105.  c = {307.9, 0.3071, 0.0002257, 1.1525e-7, 318.84, 307.9, radmax}
NB_HD double fit_s1(const nb200_map& m, double x, double y, double z) {
  if (m.kind == NB200_MAP_CONST) return m.c[0];
  double radius = sqrt(NB_POW2(x) + NB_POW2(y));
  if (m.kind == NB200_MAP_LUX_RUN03) {
    double amplitude = m.c[0] - m.c[1] * z + m.c[2] * NB_POW2(z);
    double shape = m.c[3] * sqrt(fabs(z - m.c[4]));
    double corr = -shape * NB_POW3(radius) + amplitude;
    corr /= m.c[5];
    if ((corr < 0.5 || corr > 1.5 || (corr != corr)) && radius < m.c[6]) return 1.;
    return corr;
  }
  return map_lut_rz(m, radius, z);
}
// FitS2: LUX_Run03.hh:129-146.  c[0..7] polynomial in r, c[8] normalisation, c[9] radmax
NB_HD double fit_s2(const nb200_map& m, double x, double y) {
  if (m.kind == NB200_MAP_CONST) return m.c[0];
  double r = sqrt(NB_POW2(x) + NB_POW2(y));
  if (m.kind == NB200_MAP_LUX_RUN03) {
    double corr = m.c[0] + m.c[1] * r + m.c[2] * NB_POW2(r) - m.c[3] * NB_POW3(r) + m.c[4] * NB_POW4(r) -
                  m.c[5] * NB_POW5(r) + m.c[6] * NB_POW6(r) - m.c[7] * NB_POW7(r);
    corr /= m.c[8];
    if ((corr < 0.5 || corr > 1.5 || (corr != corr)) && r < m.c[9]) return 1.;
    return corr;
  }
  return map_lut_rz(m, r, 0.);
}
// FitEF: LUX_Run03.hh:109-125
NB_HD double fit_ef(const nb200_map& m, double x, double y, double z) {
  if (m.kind == NB200_MAP_CONST) return m.c[0];
  if (m.kind == NB200_MAP_LUX_RUN03) {
    double ef = m.c[0] - m.c[1] * z + m.c[2] * NB_POW2(z) - m.c[3] * NB_POW3(z) + m.c[4] * NB_POW4(z) -
                m.c[5] * NB_POW5(z);
    if (ef <= NB_FIELD_MIN || ef > 1e5 || (ef != ef)) return NB_FIELD_MIN;
    return ef;
  }
  return map_lut_rz(m, sqrt(x * x + y * y), z);
}
// FitTBA(...)[1]: LUX_Run03.hh:148-167 (constant 0.449 for LUX)
NB_HD double fit_tba(const nb200_map& m, double x, double y, double z) {
  if (m.kind == NB200_MAP_CONST) return m.c[0];
  return map_lut_rz(m, sqrt(x * x + y * y), z);
}

#ifdef __CUDACC__
// ---- yield models ---------------------------------------------------------------------

// NESTcalc::NexONi NEST.cpp:2850-2857
NB_D double nex_o_ni(double energy, double alpha) { return alpha * erf(0.05 * energy); }

// pow with the exponents 1, 2 and 1/2 taken exactly: x, x*x and sqrt(x) are the correctly rounded
// results (what glibc's pow returns to within its last bit) and cost a few instructions instead
// of ~150.  The default model vectors use exactly these (NR: [6] = [8] = 2, [9] = 0.5,
// [10] = [11] = 1; gamma: m4 = m8 = 2).
NB_HD double pow_small(double x, double y) {
  if (y == 1.) return x;
  if (y == 2.) return x * x;
  if (y == 0.5) return sqrt(x);
  return pow(x, y);
}

// Factors of the mean-yield models that depend on the drift field and the density only
// (GetYieldNR NEST.cpp:778-780, GetYieldBetaGR :1135-1193, GetYieldGamma :441-447).  With a
// uniform field they are the same for every event of a run: the host evaluates them once
// (build_params) and the kernels reuse them; an event whose field or density differs (FitEF
// maps, nb200_yields on arbitrary lists) evaluates them itself with the same functions.
// (struct YieldHoist: nb_params.h)
NB_HD double nr_thomas_imel(double dfield, double density, const double* p) {
  return p[2] * pow(dfield, p[3]) * pow(density / NB_DENSITY, 0.3);
}
NB_HD void beta_field_terms(double dfield, double density, const double* ep, double* t) {
  t[0] = (ep[0] < 0.) ? 30.66 + (6.1978 - 30.66) / pow(1. + pow(dfield / 73.855, 2.0318), 0.41883) : ep[0];
  t[1] = (ep[2] < 0.) ? log10(dfield) * 0.13946236 + 0.52561312 : ep[2];
  t[2] = (ep[3] < 0.) ? 1.82217496 + (2.82528809 - 1.82217496) / (1. + pow(dfield / 144.65029656, -2.80532006))
                      : ep[3];
  t[3] = (ep[6] < 0.) ? 7.02921301 + (98.27936794 - 7.02921301) / (1. + pow(dfield / 256.48156448, 1.29119251))
                      : ep[6];
  t[4] = (ep[9] < 0.)
             ? 0.0508273937 + (0.1166087199 - 0.0508273937) / (1. + pow(dfield / 1.39260460e+02, -0.65763592))
             : ep[9];
  t[5] = 1.;
  if (ep[5] < 0.) {
    double coeff_TI = pow(1. / NB_DENSITY, 0.3);
    double coeff_Ni = pow(1. / NB_DENSITY, 1.4);
    double coeff_OL = pow(1. / NB_DENSITY, -1.7) / log(1. + coeff_TI * coeff_Ni * pow(NB_DENSITY, 1.7));
    t[5] = coeff_OL * log(1. + coeff_TI * coeff_Ni * pow(density, 1.7)) * pow(density, -1.7);
  }
}
NB_HD void gamma_field_terms(double dfield, double density, double* t) {
  t[0] = 33.951 + (3.3284 - 33.951) / (1. + pow(dfield / 165.34, .72665));
  t[1] = 23.156 + (10.737 - 23.156) / (1. + pow(dfield / 34.195, .87459));
  double densCorr = 240720. / pow(density, 8.2076);
  t[2] = 66.825 + (829.25 - 66.825) / (1. + pow(dfield / densCorr, .83344));
}
NB_HD bool hoisted(const YieldHoist* h, double dfield, double density) {
  return h && h->field == dfield && h->density == density;
}

// NESTcalc::GetYieldNR NEST.cpp:773-854.  massNumber: int(massNum) or the random isotope.
// err is raised where the reference throws "Quanta not conserved" (:837-840).
NB_D Yields yield_nr(double energy, double density, double dfield, int massNumber,
                     const nb200_detector& det, const double* p, int* err, const YieldHoist* h = nullptr) {
  double scale = sqrt(det.molarMass / double(massNumber));
  // E^beta as exp(beta ln E): CUDA's exp and log are good to ~1 ulp, so this is within |beta ln E| 2^-52 (< 2e-15
  // up to 1 MeV) of libm's pow -- inside the 1e-12 the yields are held to -- at a third of pow's instructions
  double Nq = p[0] * exp(p[1] * log(energy)) + 0.0;  // McMConst = 0 (NEST.cpp:19)
  if (!det.OldW13eV) Nq *= NB_ZurichEXOW;
  double ThomasImel = hoisted(h, dfield, density) ? h->nr_ti : nr_thomas_imel(dfield, density, p);
  double Qy = 1. / (ThomasImel * pow_small(energy + p[4], p[9]));
  Qy *= 1. - 1. / pow_small(1. + pow_small((energy / p[5]), p[6]), p[10]);
  if (!det.OldW13eV) Qy *= NB_ZurichEXOQ;
  double Ly = Nq / energy - Qy;
  if (Qy < 0.0) Qy = 0.0;
  if (Ly < 0.0) Ly = 0.0;
  double Ne = Qy * energy * scale;
  double Nph = Ly * energy * scale;
  Nph *= (1. - 1. / pow_small(1. + pow_small((energy / p[7]), p[8]), p[11]));
  Nq = Nph + Ne;
  double ex = exp(Ne * ThomasImel / 4.);
  double Ni = (4. / ThomasImel) * (ex - 1.);
  double Nex = (-1. / ThomasImel) * (4. * ex - (Ne + Nph) * ThomasImel - 4.);
  double NexONi = Nex / Ni;
  Wvalue w = work_function(density, det.molarMass, det.OldW13eV != 0);
  if (NexONi < w.alpha && energy > 1e2) {
    NexONi = w.alpha;
    Ni = Nq / (1. + NexONi);
    Nex = Nq - Ni;
  }
  if (NexONi > 1.0 && energy < 1.) {
    NexONi = 1.00;
    Ni = Nq / (1. + NexONi);
    Nex = Nq - Ni;
  }
  if (fabs(Nex + Ni - Nq) > 2. * NB_PHE_MIN) *err = 1;
  double L = (Nq / energy) * w.Wq_eV * 1e-3;
  Yields r{Nph, Ne, NexONi, L, dfield, -999.};
  return validity(r, energy, w.Wq_eV);
}

// NESTcalc::GetYieldBetaGR NEST.cpp:1089-1209 (xenon branch)
NB_D Yields yield_beta(double energy, double density, double dfield, const nb200_detector& det,
                       const double* ep, double multFact, const YieldHoist* h = nullptr) {
  Wvalue w = work_function(density, det.molarMass, det.OldW13eV != 0);
  double Wq_eV = w.Wq_eV, alpha = w.alpha;
  double Nq = energy * 1e3 / Wq_eV, m01, m02, m03, m04, m05, m07, m08, m09, m10;
  double tl[6];
  const double* t = tl;
  if (hoisted(h, dfield, density))
    t = h->b;
  else
    beta_field_terms(dfield, density, ep, tl);
  m01 = t[0];
  m02 = (ep[1] < 0.) ? 77.2931084 : ep[1];
  m03 = t[1];
  m04 = t[2];
  if (ep[4] < 0.)
    m05 = Nq / energy / (1. + alpha * erf(0.05 * energy)) - m01;
  else
    m05 = ep[4];
  m07 = t[3];
  m08 = (ep[7] < 0.) ? 4.285781736 : ep[7];
  m09 = (ep[8] < 0.) ? 0.3344049589 : ep[8];
  m10 = t[4];
  double Qy = m01 + (m02 - m01) / pow((1. + pow(energy / m03, m04)), m09) + m05 +
              (0.0 - m05) / pow((1. + pow(energy / m07, m08)), m10);
  if (ep[5] < 0.) {
    Qy *= t[5];
    if (!det.OldW13eV) Qy *= NB_ZurichEXOQ;
  }
  Qy *= multFact;
  double Ly = Nq / energy - Qy;
  double Ne = Qy * energy;
  double Nph = Ly * energy;
  Yields r{Nph, Ne, alpha * erf(0.05 * energy), 1., dfield, -999.};
  return validity(r, energy, Wq_eV);
}

// NESTcalc::GetYieldGamma NEST.cpp:434-467
NB_D Yields yield_gamma(double energy, double density, double dfield, const nb200_detector& det,
                        double multFact, const YieldHoist* h = nullptr) {
  Wvalue w = work_function(density, det.molarMass, det.OldW13eV != 0);
  double Wq_eV = w.Wq_eV;
  const double m3 = 2., m4 = 2., m6 = 0.;
  double tl[3];
  const double* t = tl;
  if (hoisted(h, dfield, density))
    t = h->g;
  else
    gamma_field_terms(dfield, density, tl);
  const double m1 = t[0];
  double m2 = 1000 / Wq_eV;
  double m5 = t[1];
  double m7 = t[2];
  double Nq = energy * 1000. / Wq_eV;
  double m8 = 2.;
  if (det.inGas) m8 = -2.;
  double Qy = m1 + (m2 - m1) / (1. + pow_small(energy / m3, m4)) + m5 +
              (m6 - m5) / (1. + (det.inGas ? pow(energy / m7, m8) : pow_small(energy / m7, m8)));
  if (!det.OldW13eV) Qy *= NB_ZurichEXOQ;
  Qy *= multFact;
  double Ly = Nq / energy - Qy;
  Yields r{Ly * energy, Qy * energy, nex_o_ni(energy, w.alpha), 1., dfield, -999.};
  return validity(r, energy, Wq_eV);
}

// NESTcalc::GetYieldERWeighted NEST.cpp:469-501 with default_EnergyParams /
// default_FieldParams (NEST.hh:128-129), as execNEST.cpp:1047-1048 calls it
NB_D Yields yield_er_weighted(double energy, double density, double dfield,
                              const nb200_detector& det, const double* ep, const YieldHoist* h = nullptr) {
  Wvalue w = work_function(density, det.molarMass, det.OldW13eV != 0);
  Yields yB = yield_beta(energy, density, dfield, det, ep, 1., h);
  Yields yG = yield_gamma(energy, density, dfield, det, 1., h);
  const double E0 = 0.23, E1 = 0.77, E2 = 2.95, E3 = -1.44, F0 = 421.15, F1 = 3.27;
  double weightG =
      E0 + E1 * erf(E2 * (log(energy) + E3)) * (1. - (1. / (1. + pow(dfield / F0, F1))));
  double weightB = 1. - weightG;
  Yields r;
  r.Nph = weightG * yG.Nph + weightB * yB.Nph;
  r.Ne = weightG * yG.Ne + weightB * yB.Ne;
  r.NexONi = weightG * yG.NexONi + weightB * yB.NexONi;
  r.L = weightG * yG.L + weightB * yB.L;
  r.field = weightG * yG.field + weightB * yB.field;
  r.dT = weightG * yG.dT + weightB * yB.dT;
  return validity(r, energy, w.Wq_eV);
}

// NESTcalc::GetYieldIon NEST.cpp:856-952 (xenon target: Z2 = 54; A2 = random isotope)
NB_D Yields yield_ion(double energy, double density, double dfield, double massNum,
                      double atomNum, double A2, const nb200_detector& det) {
  double A1 = massNum, Z1 = atomNum, Z2 = NB_ATOM_NUM;
  double Z_mean = pow(pow(Z1, (2. / 3.)) + pow(Z2, (2. / 3.)), 1.5);
  double E1c = pow(A1, 3.) * pow(A1 + A2, -2.) * pow(Z_mean, (4. / 3.)) * pow(Z1, (-1. / 3.)) * 500.;
  double E2c = pow(A1 + A2, 2.) * pow(A1, -1.) * Z2 * 125.;
  double gamma = 4. * A1 * A2 / pow(A1 + A2, 2.);
  double Ec_eV = gamma * E2c;
  double Constant = (2. / 3.) * (1. / sqrt(E1c) + 0.5 * sqrt(gamma / Ec_eV));
  double L = Constant * sqrt(energy * 1e3);
  double L_max = 0.96446 / (1. + pow(massNum * massNum / 19227., 0.99199));
  bool isAlpha = nearly_equal(atomNum, 2.) && nearly_equal(massNum, 4.);
  bool isPb206 = nearly_equal(A1, 206.) && nearly_equal(Z1, 82.);
  if (isAlpha) L = 0.56136 * pow(energy, 0.056972);
  if (L > L_max) L = L_max;
  double densDep = pow(density / 0.2679, -2.3245);
  double massDep = 0.02966094 * exp(0.17687876 * (massNum / 4. - 1.)) + 1. - 0.02966094;
  double fieldDep = 0.095 * pow(dfield, 0.515) - 0.025;
  if (fieldDep < 0.) fieldDep = 0.;
  if (det.inGas) fieldDep = sqrt(dfield);
  double eneDep = 0.13 - 0.0013 * energy;
  if (eneDep < 0.) eneDep = 0.;
  double ThomasImel = 0.00625 * massDep / (1. + densDep) / fieldDep + eneDep;
  double alpha = 0.64 / pow(1. + pow(density / 10., 2.), 449.61);
  double NexONi = alpha + 0.00178 * pow(atomNum, 1.587);
  if (isPb206) {
    ThomasImel = exp(-1.28683 - 1.42053e-03 * dfield);
    NexONi = 1.1;
  }
  const double logden = log10(density);
  double Wq_eV = 28.259 + 25.667 * logden - 33.611 * pow(logden, 2.) - 123.73 * pow(logden, 3.) -
                 136.47 * pow(logden, 4.) - 74.194 * pow(logden, 5.) - 20.276 * pow(logden, 6.) -
                 2.2352 * pow(logden, 7.);
  double Nq = 1e3 * L * energy / Wq_eV;
  if (!det.OldW13eV) Nq *= NB_ZurichEXOW;
  double Ni = Nq / (1. + NexONi);
  double recombProb;
  if (Ni > 0. && ThomasImel > 0.)
    recombProb = 1. - log(1. + (ThomasImel / 4.) * Ni) / ((ThomasImel / 4.) * Ni);
  else
    recombProb = 0.0;
  double Nph = Nq * NexONi / (1. + NexONi) + recombProb * Ni;
  double Ne = Nq - Nph;
  // binom_draw(Ne, ChargeLoss = 1.000) returns int64_t(Ne)  (NEST.cpp:13, 912-914)
  if (isPb206) Ne = double((long long)Ne);
  Yields r{Nph, Ne, NexONi, L, dfield, -999.};
  return validity(r, energy, Wq_eV);
}

// NESTcalc::GetYieldKr83m NEST.cpp:954-1044: the 9.4 keV, 32.1 keV or merged 41.5 keV line of
// 83mKr as a function of the time between the two decays.  That separation is
// rand_exponential(154.4 ns, minTimeSeparation, maxTimeSeparation) (RandomGen.cpp:72-85) from
// the uniform u01, or maxTimeSeparation itself when the two are nearly equal.  The 32.1 keV
// yields are the weighted sum of five GetYieldGamma calls (2, 10, 12, 18, 30 keV).  A rare
// species: out of line and looped, so that it costs the chain kernels no code they execute.
__device__ __noinline__ Yields yield_kr83m(double energy, double density, double dfield, double maxT,
                                           double minT, double u01, const nb200_detector& det, int* err) {
  if (minT > maxT && (energy < 32.0 || energy >= 32.2)) {  // throws "min greater than max" (:957-959)
    *err = 4;
    return Yields{0., 0., 0., 0., 0., 0.};
  }
  Wvalue w = work_function(density, det.molarMass, det.OldW13eV != 0);
  const double Wq_eV = w.Wq_eV;
  const double halflife = 154.4, ln2 = 0.69314718055994530942;
  double deltaT_ns;
  if (nearly_equal(minT, maxT))
    deltaT_ns = maxT;
  else {
    double r_min = 0., r_max = 1.;
    if (minT >= 0.) r_min = 1. - exp(-minT * ln2 / halflife);
    if (maxT >= 0.) r_max = 1. - exp(-maxT * ln2 / halflife);
    double r = (r_max - r_min) * u01 + r_min;
    deltaT_ns = -log(1. - r) * halflife / ln2;
  }
  double Nq_9 = 9.4e3 / Wq_eV;
  double a1 = -19.575;
  if (a1 > 0.) a1 = 0.;
  double a2 = 5.823e-4;
  double a3 = 69.986 + 3e5 / pow(deltaT_ns, 2.);
  if (a3 > 87.) a3 = 87.;
  double Nph_9 = 9.4 * (a1 * a2 * dfield * log(1. + 1. / (a2 * dfield)) + a3);
  double Ne_9 = Nq_9 - Nph_9;
  if (Ne_9 < 0.) Ne_9 = 0.;
  double NphG[5];
  const double eG[5] = {2., 10., 12., 18., 30.};
#pragma unroll 1
  for (int k = 0; k < 5; ++k) NphG[k] = yield_gamma(eG[k], density, dfield, det, 1.).Nph;
  const double Nph_2 = NphG[0], Nph_10 = NphG[1], Nph_12 = NphG[2], Nph_18 = NphG[3], Nph_30 = NphG[4];
  double Nph_76 = Nph_30 + Nph_2;
  double Nph_09 = Nph_18 + Nph_10 + Nph_2 + Nph_2;
  double Nph_15 = Nph_18 + Nph_12 + Nph_2;
  double Nph_32 = 0.76 * Nph_76 + 0.15 * Nph_15 + 0.09 * Nph_09;
  double Ne_32 = 32e3 / Wq_eV - Nph_32;
  double Nph, Ne;
  if (energy > 9.35 && energy < 9.45) {
    Nph = Nph_9;
    Ne = Ne_9;
  } else if (energy >= 32.0 && energy < 32.2) {
    Nph = Nph_32;
    Ne = Ne_32;
  } else {  // merged 41.5 keV decay
    Ne = Ne_9 + Ne_32;
    Nph = Nph_9 + Nph_32;
  }
  Yields r{Nph, Ne, 0., 1., dfield, deltaT_ns};
  if (!det.OldW13eV) {
    Ne *= NB_ZurichEXOQ;
    Nph = (r.Nph + r.Ne) - Ne;
    r.Nph = Nph;
    r.Ne = Ne;
  }
  r.NexONi = nex_o_ni(energy, w.alpha);
  return validity(r, energy, Wq_eV);
}

// NESTcalc::GetYields dispatch NEST.cpp:1211-1252.  A2iso: random Xe isotope, used by NR when
// massNum ~ 0 (:789-790) and by ion (:864).  u01: uniform for Kr83m's decay-time separation
// (massNum / atomNum carry max / min separation there, :1238-1239).
NB_D Yields get_yields(int species, double energy, double density, double dfield, double massNum,
                       double atomNum, int A2iso, const nb200_detector& det, const nb200_model& mo,
                       int* err, double u01 = 0.5, const YieldHoist* h = nullptr) {
  switch (species) {
    case NB200_NR:
    case NB200_WIMP:
    case NB200_B8:
    case NB200_atmNu:
    case NB200_DD:
    case NB200_AmBe:
    case NB200_Cf: {
      int massNumber = nearly_equal(massNum, 0.) ? A2iso : int(massNum);
      return yield_nr(energy, density, dfield, massNumber, det, mo.NRYieldsParam, err, h);
    }
    case NB200_ion:
      return yield_ion(energy, density, dfield, massNum, atomNum, double(A2iso), det);
    case NB200_gammaRay:
      return yield_gamma(energy, density, dfield, det, 1., h);
    case NB200_Kr83m:
      return yield_kr83m(energy, density, dfield, massNum, atomNum, u01, det, err);
    default:
      return yield_beta(energy, density, dfield, det, mo.ERYieldsParam, 1., h);
  }
}

#endif  // __CUDACC__
}  // namespace nb
