// TEST INFRASTRUCTURE ONLY -- see nest_oracle.h.  Never linked into or called by the product.
//
// Sequential CPU restatement of the reference's hot path (NESTCollaboration/nest), written
// as free functions over the flattened detector POD of include/nest_b200.h.  It keeps the
// reference's generator (xoroshiro128+ seeded through splitmix64), its draw ORDER and its
// libm / libstdc++ calls, so that with equal seeds it reproduces the unmodified reference
// (oracle/_ref) bit for bit; tests/test_oracle_pin.py checks exactly that.  Each function
// cites the reference lines it follows (paths relative to the reference root).
#include "nest_oracle.h"

#include <cfloat>
#include <cmath>
#include <cstring>
#include <random>
#include <stdexcept>
#include <vector>

namespace {

// ---- constants: include/NEST/NEST.hh:81-113, src/NEST.cpp:7-19 ------------------------------
const double kW_SCINT = 8.5e-3, kAVO = 6.0221409e+23, kATOM = 54., kPHE_MIN = 1e-6, kFIELD_MIN = 1.,
             kDENSITY = 2.90, kEPS_GAS = 1.00126, kEPS_LIQ = 1.85, kZurichEXOW = 1.1716263232,
             kZurichEXOQ = 1.08, kMcM = 0.0;
const int kSPIKES_MAXM = 120;
const double kRideal = 8.31446261815324, kRealGasA = 0.4250, kRealGasB = 5.105e-5;
const double kSqrt2 = __builtin_sqrt(2.), kSqrt2PI = __builtin_sqrt(2. * M_PI),
             kInvSqrt2PI = 1. / __builtin_sqrt(2. * M_PI), kTwoPI = 2. * M_PI,
             kSqrt2DivPI = __builtin_sqrt(2. / M_PI), kLog2 = __builtin_log(2.);

// ValidityTests::nearlyEqual src/ValidityTests.cpp:6-21
bool nearlyEqual(double a, double b, double epsilon = 1e-9) {
  if (a == b) return true;
  double absA = fabs(a), absB = fabs(b), diff = fabs(a - b);
  if (a == 0 || b == 0 || (absA + absB < DBL_MIN)) return diff < (epsilon * DBL_MIN);
  return diff / fmin((absA + absB), DBL_MAX) < epsilon;
}

// ---- generator: include/NEST/xoroshiro.hh:54-116,538-541; src/RandomGen.cpp:16-42 -----------
struct Xoro {
  typedef uint64_t result_type;
  uint64_t s0, s1;
  static constexpr result_type min() { return 0; }
  static constexpr result_type max() { return ~result_type(0); }
  static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
  static uint64_t splitmix(uint64_t z) {
    z += 0x9e3779b97f4a7c15ULL;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
  }
  void seed(uint64_t s) {
    uint64_t a = splitmix(s);
    uint64_t b = splitmix(a);
    s0 = a;
    s1 = (a || b) ? b : 1;
  }
  result_type operator()() {
    const uint64_t r = s0 + s1;
    s1 ^= s0;
    s0 = rotl(s0, 24) ^ s1 ^ (s1 << 16);
    s1 = rotl(s1, 37);
    return r;
  }
};
Xoro G;  // the reference's RandomGen is a process-wide singleton (RandomGen.cpp:8-14)

// ---- samplers: src/RandomGen.cpp:39-182 ---------------------------------------------------------
double rand_uniform() {
  return (static_cast<double>(G()) - 0.) / static_cast<double>(Xoro::max() - Xoro::min());
}
double rand_gauss(double mean, double sigma, bool zero_min = false) {
  double u = rand_uniform(), v = rand_uniform();
  double draw = mean + sigma * kSqrt2 * sqrt(-log(u)) * cos(kTwoPI * v);
  if (zero_min) return std::max(draw, 0.);
  return draw;
}
double rand_zero_trunc_gauss(double mean, double sigma) {
  double r = rand_gauss(mean, sigma, false);
  while (r <= 0) r = rand_gauss(mean, sigma, false);
  return r;
}
double FindNewMean(double sigma) {
  double a = -1. / sigma;
  double phi = exp(-0.5 * a * a) / __builtin_sqrt(2. * M_PI);
  double Phi = 0.5 * (1. + erf(a / kSqrt2));
  return 1. + (phi / (1. - Phi)) * sigma;
}
double rand_exponential(double half_life, double t_min, double t_max) {
  double r = rand_uniform();
  double r_min = 0, r_max = 1;
  if (t_min >= 0) r_min = 1 - exp(-t_min * kLog2 / half_life);
  if (t_max >= 0) r_max = 1 - exp(-t_max * kLog2 / half_life);
  r = (r_max - r_min) * r + r_min;
  return -log(1 - r) * half_life / kLog2;
}
double rand_skewGauss(double xi, double omega, double alpha) {
  double delta = alpha / sqrt(1 + alpha * alpha);
  double gamma1 = (0.5 * (4. - M_PI)) * (pow(delta * kSqrt2DivPI, 3.) / pow(1 - 2. * delta * delta / M_PI, 1.5));
  double muz = delta * kSqrt2DivPI;
  double sigz = sqrt(1. - muz * muz);
  double m_o = 0.;
  if (alpha > 0.) m_o = muz - 0.5 * gamma1 * sigz - 0.5 * exp(-kTwoPI / alpha);
  if (alpha < 0.) m_o = muz - 0.5 * gamma1 * sigz + 0.5 * exp(kTwoPI / alpha);
  double mode = xi + omega * m_o;
  double height = exp(-0.5 * (pow((mode - xi) / omega, 2.))) / (kSqrt2PI * omega) *
                  erfc(-1. * alpha * (mode - xi) / omega / kSqrt2);
  double minX = xi - 6. * omega, maxX = xi + 6. * omega;
  for (;;) {
    double testX = minX + (maxX - minX) * rand_uniform();
    double testY = height * rand_uniform();
    double testProb = exp(-0.5 * (pow((testX - xi) / omega, 2.))) / (kSqrt2PI * omega) *
                      erfc(-1. * alpha * (testX - xi) / omega / kSqrt2);
    if (testProb >= testY) return testX;
  }
}
uint64_t poisson_draw(double mean) { return std::poisson_distribution<uint64_t>(mean)(G); }
int64_t binom_draw(int64_t N0, double prob) { return std::binomial_distribution<int64_t>(N0, prob)(G); }
int SelectRanXeAtom() {
  double isotope = rand_uniform() * 100.;
  if (isotope > 0.000 && isotope <= 0.090) return 124;
  if (isotope > 0.090 && isotope <= 0.180) return 126;
  if (isotope > 0.180 && isotope <= 2.100) return 128;
  if (isotope > 2.100 && isotope <= 28.54) return 129;
  if (isotope > 28.54 && isotope <= 32.62) return 130;
  if (isotope > 32.62 && isotope <= 53.80) return 131;
  if (isotope > 53.80 && isotope <= 80.69) return 132;
  if (isotope > 80.69 && isotope <= 91.13) return 134;
  return 136;
}

// ---- detector maps: include/Detectors/LUX_Run03.hh:90-167, VDetector.hh:131-148 ------------
double lut_rz(const nb200_map& m, double r, double z) {
  double fr = (m.n0 > 1) ? (r - m.c[0]) / (m.c[1] - m.c[0]) * double(m.n0 - 1) : 0.;
  double fz = (m.n1 > 1) ? (z - m.c[2]) / (m.c[3] - m.c[2]) * double(m.n1 - 1) : 0.;
  fr = fmin(fmax(fr, 0.), double(m.n0 - 1));
  fz = fmin(fmax(fz, 0.), double(m.n1 - 1));
  int i0 = std::min(int(fr), std::max(m.n0 - 2, 0)), j0 = std::min(int(fz), std::max(m.n1 - 2, 0));
  int i1 = std::min(i0 + 1, m.n0 - 1), j1 = std::min(j0 + 1, m.n1 - 1);
  double tr = fr - i0, tz = fz - j0;
  const double* t = m.lut;
  double a = t[size_t(i0) * m.n1 + j0] + (t[size_t(i0) * m.n1 + j1] - t[size_t(i0) * m.n1 + j0]) * tz;
  double b = t[size_t(i1) * m.n1 + j0] + (t[size_t(i1) * m.n1 + j1] - t[size_t(i1) * m.n1 + j0]) * tz;
  return a + (b - a) * tr;
}
double FitS1(const nb200_detector& d, double x, double y, double z) {
  const nb200_map& m = d.FitS1;
  if (m.kind == NB200_MAP_CONST) return m.c[0];
  double radius = sqrt(pow(x, 2.) + pow(y, 2.));
  if (m.kind == NB200_MAP_LUT_RZ) return lut_rz(m, radius, z);
  double amplitude = 307.9 - 0.3071 * z + 0.0002257 * pow(z, 2.);
  double shape = 1.1525e-7 * sqrt(std::abs(z - 318.84));
  double finalCorr = -shape * pow(radius, 3.) + amplitude;
  finalCorr /= 307.9;
  if ((finalCorr < 0.5 || finalCorr > 1.5 || std::isnan(finalCorr)) && radius < d.radmax) return 1.;
  return finalCorr;
}
double FitEF(const nb200_detector& d, double x, double y, double z) {
  const nb200_map& m = d.FitEF;
  if (m.kind == NB200_MAP_CONST) return m.c[0];
  if (m.kind == NB200_MAP_LUT_RZ) return lut_rz(m, sqrt(x * x + y * y), z);
  double finalEF = 158.92 - 0.2209000 * pow(z, 1.) + 0.0024485 * pow(z, 2.) - 8.7098e-6 * pow(z, 3.) +
                   1.5049e-8 * pow(z, 4.) - 1.0110e-11 * pow(z, 5.);
  if (finalEF <= kFIELD_MIN || finalEF > 1e5 || std::isnan(finalEF)) return kFIELD_MIN;
  return finalEF;
}
double FitS2(const nb200_detector& d, double x, double y) {
  const nb200_map& m = d.FitS2;
  if (m.kind == NB200_MAP_CONST) return m.c[0];
  double radius = sqrt(pow(x, 2.) + pow(y, 2.));
  if (m.kind == NB200_MAP_LUT_RZ) return lut_rz(m, radius, 0.);
  double finalCorr = 9156.3 + 6.22750 * pow(radius, 1.) + 0.38126 * pow(radius, 2.) -
                     0.017144 * pow(radius, 3.) + 0.0002474 * pow(radius, 4.) - 1.6953e-6 * pow(radius, 5.) +
                     5.6513e-9 * pow(radius, 6.) - 7.3989e-12 * pow(radius, 7.);
  finalCorr /= 9156.3;
  if ((finalCorr < 0.5 || finalCorr > 1.5 || std::isnan(finalCorr)) && radius < d.radmax) return 1.;
  return finalCorr;
}
double FitTBA1(const nb200_detector& d, double x, double y, double z) {
  const nb200_map& m = d.FitTBA;
  if (m.kind == NB200_MAP_CONST) return m.c[0];
  return lut_rz(m, sqrt(x * x + y * y), z);
}

// ---- thermodynamics / transport -------------------------------------------------------------------
// NESTcalc::GetDensity src/NEST.cpp:2278-2350 (xenon)
double GetDensity(double Kelvin, double bara, bool& inGas, double molarMass) {
  if (Kelvin < 161.40) return 3.41;
  double VaporP_bar;
  if (Kelvin < 289.7)
    VaporP_bar = pow(10., 4.0519 - 667.16 / Kelvin);
  else
    VaporP_bar = DBL_MAX;
  if (bara < VaporP_bar || inGas) {
    double p_Pa = bara * 1e5;
    double density = 1. / (pow(kRideal * Kelvin, 3.) / (p_Pa * pow(kRideal * Kelvin, 2.) + kRealGasA * p_Pa * p_Pa) +
                           kRealGasB);
    density *= molarMass * 1e-6;
    inGas = true;
    return density;
  }
  inGas = false;
  return 2.9970938084691329E+02 * exp(-8.2598864714323525E-02 * Kelvin) -
         1.8801286589442915E+06 * exp(-pow((Kelvin - 4.0820251276172212E+02) / 2.7863170223154846E+01, 2.)) -
         5.4964506351743057E+03 * exp(-pow((Kelvin - 6.3688597345042672E+02) / 1.1225818853661815E+02, 2.)) +
         8.3450538370682614E+02 * exp(-pow((Kelvin + 4.8840568924597342E+01) / 7.3804147172071107E+03, 2.)) -
         8.3086310405942265E+02;
}
// NESTcalc::WorkFunction src/NEST.cpp:2829-2848
void WorkFunction(double density, double MolarMass, bool OldW13eV, double& Wq_eV, double& alpha) {
  alpha = 0.067366 + density * 0.039693;
  double eDensity = (density / MolarMass) * kAVO * kATOM;
  Wq_eV = 18.7263 - 1.01e-23 * eDensity;
  if (OldW13eV) Wq_eV *= kZurichEXOW;
}
// NESTcalc::GetDriftVelocity_Liquid src/NEST.cpp:2366-2521 (xenon, StdDev = -1)
double DriftVelocityLiquid(double Kelvin, double eField) {
  static const double polyExp[11][7] = {
      {-3.1046, 27.037, -2.1668, 193.27, -4.8024, 646.04, 9.2471}, {-2.7394, 22.760, -1.7775, 222.72, -5.0836, 724.98, 8.7189},
      {-2.3646, 164.91, -1.6984, 21.473, -4.4752, 1202.2, 7.9744}, {-1.8097, 235.65, -1.7621, 36.855, -3.5925, 1356.2, 6.7865},
      {-1.5000, 37.021, -1.1430, 6.4590, -4.0337, 855.43, 5.4238}, {-1.4939, 47.879, 0.12608, 8.9095, -1.3480, 1310.9, 2.7598},
      {-1.5389, 26.602, -.44589, 196.08, -1.1516, 1810.8, 2.8912}, {-1.5000, 28.510, -.21948, 183.49, -1.4320, 1652.9, 2.884},
      {-1.1781, 49.072, -1.3008, 3438.4, -.14817, 312.12, 2.8049}, {1.2466, 85.975, -.88005, 918.57, -3.0085, 27.568, 2.3823},
      {334.60, 37.556, 0.92211, 345.27, -338.00, 37.346, 1.9834}};
  static const double T[11] = {100., 120., 140., 155., 157., 163., 165., 167., 184., 200., 230.};
  if (Kelvin < 100.) Kelvin = 100.;
  if (Kelvin > 230.) Kelvin = 230.;
  int i = -1;
  for (int k = 0; k < 9 && i < 0; ++k)
    if (Kelvin >= T[k] && Kelvin < T[k + 1]) i = k;
  if (i < 0) {
    if (Kelvin >= T[9] && Kelvin <= T[10])
      i = 9;
    else
      throw std::runtime_error("ERROR: TEMPERATURE OUT OF RANGE (100-230 K)");
  }
  int j = i + 1;
  double Ti = T[i], Tf = T[j];
  double vi = polyExp[i][0] * exp(-eField / polyExp[i][1]) + polyExp[i][2] * exp(-eField / polyExp[i][3]) +
              polyExp[i][4] * exp(-eField / polyExp[i][5]) + polyExp[i][6];
  double vf = polyExp[j][0] * exp(-eField / polyExp[j][1]) + polyExp[j][2] * exp(-eField / polyExp[j][3]) +
              polyExp[j][4] * exp(-eField / polyExp[j][5]) + polyExp[j][6];
  if (nearlyEqual(Kelvin, Ti)) return vi;
  if (nearlyEqual(Kelvin, Tf)) return vf;
  double speed, slope, offset;
  if (vf < vi) {
    offset = (sqrt((Tf * (vf - vi) - Ti * (vf - vi) - 4.) * (vf - vi)) + sqrt(Tf - Ti) * (vf + vi)) / (2. * sqrt(Tf - Ti));
    slope = -(sqrt(Tf - Ti) * sqrt((Tf * (vf - vi) - Ti * (vf - vi) - 4.) * (vf - vi)) - (Tf + Ti) * (vf - vi)) /
            (2. * (vf - vi));
    speed = 1. / (Kelvin - slope) + offset;
  } else {
    slope = (vf - vi) / (Tf - Ti);
    speed = slope * (Kelvin - Ti) + vi;
  }
  if (speed <= 0.) speed = 0.1;
  return speed;
}
// NESTcalc::SetDriftVelocity_NonUniform src/NEST.cpp:2664-2697 (liquid)
std::vector<double> DriftTableNonUniform(const nb200_detector& d, double zStep, double dx, double dy) {
  std::vector<double> speedTable;
  double driftTime, zz;
  for (double pos_z = 0.0; pos_z < d.TopDrift; pos_z += zStep) {
    driftTime = 0.0;
    for (zz = pos_z; zz < d.TopDrift; zz += zStep) {
      if (pos_z > d.gate)
        driftTime += zStep / DriftVelocityLiquid(d.T_Kelvin, d.E_gas / (kEPS_LIQ / std::abs(kEPS_GAS)) * 1e3);
      else
        driftTime += zStep / DriftVelocityLiquid(d.T_Kelvin, FitEF(d, dx, dy, zz));
    }
    speedTable.emplace_back((zz - pos_z) / driftTime);
  }
  return speedTable;
}

// ---- yields: src/NEST.cpp:434-501, 773-952, 1089-1271 -----------------------------------------------
struct Y {
  double Nph, Ne, NexONi, L, F, dT;
};
Y YieldResultValidity(Y res, double energy, double Wq_eV) {  // :1254-1271
  if (res.Nph > energy / kW_SCINT) res.Nph = energy / kW_SCINT;
  if (res.Ne > energy / kW_SCINT) res.Ne = energy / kW_SCINT;
  if (res.Nph < 0.) res.Nph = 0.;
  if (res.Ne < 0.) res.Ne = 0.;
  res.L = std::max(0., std::min(res.L, 1.));
  if (energy < 0.001 * Wq_eV / res.L) {
    res.Nph = 0.;
    res.Ne = 0.;
  }
  return res;
}
double NexONi(const nb200_detector& d, double energy, double density) {  // :2850-2857
  double Wq, alpha;
  WorkFunction(density, d.molarMass, d.OldW13eV, Wq, alpha);
  return alpha * erf(0.05 * energy);
}
Y GetYieldGamma(const nb200_detector& d, double energy, double density, double dfield, double multFact = 1.) {  // :434-467
  double Wq_eV, alpha;
  WorkFunction(density, d.molarMass, d.OldW13eV, Wq_eV, alpha);
  const double m3 = 2., m4 = 2., m6 = 0.;
  const double m1 = 33.951 + (3.3284 - 33.951) / (1. + pow(dfield / 165.34, .72665));
  double m2 = 1000 / Wq_eV;
  double m5 = 23.156 + (10.737 - 23.156) / (1. + pow(dfield / 34.195, .87459));
  double densCorr = 240720. / pow(density, 8.2076);
  double m7 = 66.825 + (829.25 - 66.825) / (1. + pow(dfield / densCorr, .83344));
  double Nq = energy * 1000. / Wq_eV;
  double m8 = 2.;
  if (d.inGas) m8 = -2.;
  double Qy = m1 + (m2 - m1) / (1. + pow(energy / m3, m4)) + m5 + (m6 - m5) / (1. + pow(energy / m7, m8));
  if (!d.OldW13eV) Qy *= kZurichEXOQ;
  Qy *= multFact;
  double Ly = Nq / energy - Qy;
  Y r{Ly * energy, Qy * energy, NexONi(d, energy, density), 1, dfield, -999};
  return YieldResultValidity(r, energy, Wq_eV);
}
Y GetYieldBetaGR(const nb200_detector& d, double energy, double density, double dfield, const double* ER,
                 double multFact = 1.) {  // :1089-1209
  double Wq_eV, alpha;
  WorkFunction(density, d.molarMass, d.OldW13eV, Wq_eV, alpha);
  double Nq = energy * 1e3 / Wq_eV, m01, m02, m03, m04, m05, m07, m08, m09, m10;
  if (ER[0] < 0.)
    m01 = 30.66 + (6.1978 - 30.66) / pow(1. + pow(dfield / 73.855, 2.0318), 0.41883);
  else
    m01 = ER[0];
  if (ER[1] < 0.) m02 = 77.2931084; else m02 = ER[1];
  if (ER[2] < 0.) m03 = log10(dfield) * 0.13946236 + 0.52561312; else m03 = ER[2];
  if (ER[3] < 0.)
    m04 = 1.82217496 + (2.82528809 - 1.82217496) / (1. + pow(dfield / 144.65029656, -2.80532006));
  else
    m04 = ER[3];
  if (ER[4] < 0.) m05 = Nq / energy / (1. + alpha * erf(0.05 * energy)) - m01; else m05 = ER[4];
  if (ER[6] < 0.)
    m07 = 7.02921301 + (98.27936794 - 7.02921301) / (1. + pow(dfield / 256.48156448, 1.29119251));
  else
    m07 = ER[6];
  if (ER[7] < 0.) m08 = 4.285781736; else m08 = ER[7];
  if (ER[8] < 0.) m09 = 0.3344049589; else m09 = ER[8];
  if (ER[9] < 0.)
    m10 = 0.0508273937 + (0.1166087199 - 0.0508273937) / (1. + pow(dfield / 1.39260460e+02, -0.65763592));
  else
    m10 = ER[9];
  double Qy = m01 + (m02 - m01) / pow((1. + pow(energy / m03, m04)), m09) + m05 +
              (0.0 - m05) / pow((1. + pow(energy / m07, m08)), m10);
  if (ER[5] < 0.) {
    double coeff_TI = pow(1. / kDENSITY, 0.3);
    double coeff_Ni = pow(1. / kDENSITY, 1.4);
    double coeff_OL = pow(1. / kDENSITY, -1.7) / log(1. + coeff_TI * coeff_Ni * pow(kDENSITY, 1.7));
    Qy *= coeff_OL * log(1. + coeff_TI * coeff_Ni * pow(density, 1.7)) * pow(density, -1.7);
    if (!d.OldW13eV) Qy *= kZurichEXOQ;
  }
  Qy *= multFact;
  double Ly = Nq / energy - Qy;
  double Ne = Qy * energy, Nph = Ly * energy;
  Y r{Nph, Ne, alpha * erf(0.05 * energy), 1., dfield, -999};
  return YieldResultValidity(r, energy, Wq_eV);
}
Y GetYieldERWeighted(const nb200_detector& d, double energy, double density, double dfield, const double* ER) {  // :469-501
  static const double EP[4] = {0.23, 0.77, 2.95, -1.44}, FP[2] = {421.15, 3.27};  // NEST.hh:128-129
  double Wq_eV, alpha;
  WorkFunction(density, d.molarMass, d.OldW13eV, Wq_eV, alpha);
  Y yB = GetYieldBetaGR(d, energy, density, dfield, ER);
  Y yG = GetYieldGamma(d, energy, density, dfield);
  double weightG = EP[0] + EP[1] * erf(EP[2] * (log(energy) + EP[3])) * (1. - (1. / (1. + pow(dfield / FP[0], FP[1]))));
  double weightB = 1. - weightG;
  Y r;
  r.Nph = weightG * yG.Nph + weightB * yB.Nph;
  r.Ne = weightG * yG.Ne + weightB * yB.Ne;
  r.NexONi = weightG * yG.NexONi + weightB * yB.NexONi;
  r.L = weightG * yG.L + weightB * yB.L;
  r.F = weightG * yG.F + weightB * yB.F;
  r.dT = weightG * yG.dT + weightB * yB.dT;
  return YieldResultValidity(r, energy, Wq_eV);
}
Y GetYieldNR(const nb200_detector& d, double energy, double density, double dfield, double massNum,
             const double* NR) {  // :773-854
  int massNumber;
  if (nearlyEqual(massNum, 0.))
    massNumber = SelectRanXeAtom();
  else
    massNumber = int(massNum);
  double ScaleFactor[2];
  ScaleFactor[0] = sqrt(d.molarMass / (double)massNumber);
  ScaleFactor[1] = ScaleFactor[0];
  double Nq = NR[0] * pow(energy, NR[1]) + kMcM;
  if (!d.OldW13eV) Nq *= kZurichEXOW;
  double ThomasImel = NR[2] * pow(dfield, NR[3]) * pow(density / kDENSITY, 0.3);
  double Qy = 1. / (ThomasImel * pow(energy + NR[4], NR[9]));
  Qy *= 1. - 1. / pow(1. + pow((energy / NR[5]), NR[6]), NR[10]);
  if (!d.OldW13eV) Qy *= kZurichEXOQ;
  double Ly = Nq / energy - Qy;
  if (Qy < 0.0) Qy = 0.0;
  if (Ly < 0.0) Ly = 0.0;
  double Ne = Qy * energy * ScaleFactor[1];
  double Nph = Ly * energy * ScaleFactor[0];
  Nph *= (1. - 1. / pow(1. + pow((energy / NR[7]), NR[8]), NR[11]));
  Nq = Nph + Ne;
  double Ni = (4. / ThomasImel) * (exp(Ne * ThomasImel / 4.) - 1.);
  double Nex = (-1. / ThomasImel) * (4. * exp(Ne * ThomasImel / 4.) - (Ne + Nph) * ThomasImel - 4.);
  double NexONi_ = Nex / Ni;
  double Wq_eV, alpha;
  WorkFunction(density, d.molarMass, d.OldW13eV, Wq_eV, alpha);
  if (NexONi_ < alpha && energy > 1e2) {
    NexONi_ = alpha;
    Ni = Nq / (1. + NexONi_);
    Nex = Nq - Ni;
  }
  if (NexONi_ > 1.0 && energy < 1.) {
    NexONi_ = 1.00;
    Ni = Nq / (1. + NexONi_);
    Nex = Nq - Ni;
  }
  if (std::abs(Nex + Ni - Nq) > 2. * kPHE_MIN)
    throw std::runtime_error("ERROR: Quanta not conserved. Tell Matthew Immediately!");
  double L = (Nq / energy) * Wq_eV * 1e-3;
  Y r{Nph, Ne, NexONi_, L, dfield, -999};
  return YieldResultValidity(r, energy, Wq_eV);
}
Y GetYieldIon(const nb200_detector& d, double energy, double density, double dfield, double massNum,
              double atomNum) {  // :856-952
  double A1 = massNum, A2 = SelectRanXeAtom();
  double Z1 = atomNum, Z2 = kATOM;
  double Z_mean = pow(pow(Z1, (2. / 3.)) + pow(Z2, (2. / 3.)), 1.5);
  double E1c = pow(A1, 3.) * pow(A1 + A2, -2.) * pow(Z_mean, (4. / 3.)) * pow(Z1, (-1. / 3.)) * 500.;
  double E2c = pow(A1 + A2, 2.) * pow(A1, -1.) * Z2 * 125.;
  double gamma = 4. * A1 * A2 / pow(A1 + A2, 2.);
  double Ec_eV = gamma * E2c;
  double Constant = (2. / 3.) * (1. / sqrt(E1c) + 0.5 * sqrt(gamma / Ec_eV));
  double L = Constant * sqrt(energy * 1e3);
  double L_max = 0.96446 / (1. + pow(massNum * massNum / 19227., 0.99199));
  if (nearlyEqual(atomNum, 2.) && nearlyEqual(massNum, 4.)) L = 0.56136 * pow(energy, 0.056972);
  if (L > L_max) L = L_max;
  double densDep = pow(density / 0.2679, -2.3245);
  double massDep = 0.02966094 * exp(0.17687876 * (massNum / 4. - 1.)) + 1. - 0.02966094;
  double fieldDep = 0.095 * pow(dfield, 0.515) - 0.025;
  if (fieldDep < 0.) fieldDep = 0.;
  if (d.inGas) fieldDep = sqrt(dfield);
  double eneDep = 0.13 - 0.0013 * energy;
  if (eneDep < 0.) eneDep = 0.;
  double ThomasImel = 0.00625 * massDep / (1. + densDep) / fieldDep + eneDep;
  double alpha = 0.64 / pow(1. + pow(density / 10., 2.), 449.61);
  double NexONi_ = alpha + 0.00178 * pow(atomNum, 1.587);
  if (nearlyEqual(A1, 206.) && nearlyEqual(Z1, 82.)) {
    ThomasImel = exp(-1.28683 - 1.42053e-03 * dfield);
    NexONi_ = 1.1;
  }
  const double logden = log10(density);
  double Wq_eV = 28.259 + 25.667 * logden - 33.611 * pow(logden, 2.) - 123.73 * pow(logden, 3.) -
                 136.47 * pow(logden, 4.) - 74.194 * pow(logden, 5.) - 20.276 * pow(logden, 6.) -
                 2.2352 * pow(logden, 7.);
  double Nq = 1e3 * L * energy / Wq_eV;
  if (!d.OldW13eV) Nq *= kZurichEXOW;
  double Ni = Nq / (1. + NexONi_);
  double recombProb;
  if (Ni > 0. && ThomasImel > 0.)
    recombProb = 1. - log(1. + (ThomasImel / 4.) * Ni) / ((ThomasImel / 4.) * Ni);
  else
    recombProb = 0.0;
  double Nph = Nq * NexONi_ / (1. + NexONi_) + recombProb * Ni;
  double Ne = Nq - Nph;
  if (nearlyEqual(A1, 206.) && nearlyEqual(Z1, 82.)) Ne = binom_draw(Ne, 1.000);
  Y r{Nph, Ne, NexONi_, L, dfield, -999};
  return YieldResultValidity(r, energy, Wq_eV);
}
Y GetYieldKr83m(const nb200_detector& d, double energy, double density, double dfield, double maxTimeSeparation,
                double minTimeSeparation) {  // :954-1044
  if (minTimeSeparation > maxTimeSeparation && (energy < 32.0 || energy >= 32.2))
    throw std::runtime_error("ERROR: min greater than max");
  double Wq_eV, alpha;
  WorkFunction(density, d.molarMass, d.OldW13eV, Wq_eV, alpha);
  const double deltaT_ns_halflife = 154.4;
  double deltaT_ns = -999.;
  if (nearlyEqual(minTimeSeparation, maxTimeSeparation)) deltaT_ns = maxTimeSeparation;
  if (!nearlyEqual(minTimeSeparation, maxTimeSeparation))
    deltaT_ns = rand_exponential(deltaT_ns_halflife, minTimeSeparation, maxTimeSeparation);
  double Nq_9 = 9.4e3 / Wq_eV;
  double a1 = -19.575;
  if (a1 > 0.) a1 = 0.;
  double a2 = 5.823e-4;
  double a3 = 69.986 + 3e5 / pow(deltaT_ns, 2.);
  if (a3 > 87.) a3 = 87.;
  double Nph_9 = 9.4 * (a1 * a2 * dfield * log(1. + 1. / (a2 * dfield)) + a3);
  double Ne_9 = Nq_9 - Nph_9;
  if (Ne_9 < 0.) Ne_9 = 0.;
  double Nph_2 = GetYieldGamma(d, 2., density, dfield).Nph;
  double Nph_10 = GetYieldGamma(d, 10., density, dfield).Nph;
  double Nph_12 = GetYieldGamma(d, 12., density, dfield).Nph;
  double Nph_18 = GetYieldGamma(d, 18., density, dfield).Nph;
  double Nph_30 = GetYieldGamma(d, 30., density, dfield).Nph;
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
  } else {
    Ne = Ne_9 + Ne_32;
    Nph = Nph_9 + Nph_32;
  }
  Y r{Nph, Ne, 0., 1, dfield, deltaT_ns};
  if (!d.OldW13eV) {
    Ne *= kZurichEXOQ;
    Nph = (r.Nph + r.Ne) - Ne;
    r.Nph = Nph;
    r.Ne = Ne;
  }
  r.NexONi = NexONi(d, energy, density);
  return YieldResultValidity(r, energy, Wq_eV);
}
Y GetYields(const nb200_detector& d, int species, double energy, double density, double dfield, double massNum,
            double atomNum, const double* NR, const double* ER) {  // :1211-1252
  switch (species) {
    case NB200_NR: case NB200_WIMP: case NB200_B8: case NB200_atmNu: case NB200_DD: case NB200_AmBe: case NB200_Cf:
      return GetYieldNR(d, energy, density, dfield, massNum, NR);
    case NB200_ion:
      return GetYieldIon(d, energy, density, dfield, massNum, atomNum);
    case NB200_gammaRay:
      return GetYieldGamma(d, energy, density, dfield);
    case NB200_Kr83m:
      return GetYieldKr83m(d, energy, density, dfield, massNum, atomNum);  // :1238-1239
    default:
      return GetYieldBetaGR(d, energy, density, dfield, ER);
  }
}

// ---- quanta: src/NEST.cpp:164-432 ---------------------------------------------------------------------
double RecombOmegaNR(double elecFrac, const double* W) {
  double omega = W[2] * exp(-0.5 * pow(elecFrac - W[3], 2.) / (W[4] * W[4]));
  if (omega < 0.) omega = 0;
  return omega;
}
double RecombOmegaER(double efield, double elecFrac, double numQuanta, const double* W, int nw) {
  double ampl = 0.07 + (W[7] - 0.07) / pow(1. + pow(efield / 295.2, 251.6), 0.0069114);
  if (ampl < 0.) ampl = 0.;
  double cntr = W[8], wide = W[9];
  if (wide > cntr) throw std::runtime_error("Low-E r-fluct Gauss peak's width greater than centroid.");
  double skew = W[10];
  double mode = cntr + 2. * kInvSqrt2PI * skew * wide / sqrt(1. + skew * skew);
  double norm = 1. / (exp(-0.5 * pow(mode - cntr, 2.) / (wide * wide)) * (1. + erf(skew * (mode - cntr) / (wide * kSqrt2))));
  double omega;
  if (cntr < 1.)
    omega = norm * ampl * exp(-0.5 * pow(elecFrac - cntr, 2.) / (wide * wide)) *
            (1. + erf(skew * (elecFrac - cntr) / (wide * kSqrt2)));
  else {
    double ampl2 = 0.06103, cntr2 = 5.207, wide2 = 1.361, skew2 = -2.495;
    if (nw > 15) {
      ampl2 = W[13];
      cntr2 = W[14];
      wide2 = W[15];
    }
    if (nw > 16) skew2 = W[16];
    double mode2 = cntr2 + 2. * kInvSqrt2PI * skew2 * wide2 / sqrt(1. + skew2 * skew2);
    double logNQ = log10(numQuanta);
    double norm2 = 1. / (exp(-0.5 * pow(mode2 - cntr2, 2.) / (wide2 * wide2)) *
                         (1. + erf(skew2 * (mode2 - cntr2) / (wide2 * kSqrt2))));
    omega = 0.00 +
            norm * ampl * exp(-0.5 * pow(logNQ - cntr, 2.) / (wide * wide)) * (1. + erf(skew * (logNQ - cntr) / (wide * kSqrt2))) +
            norm2 * ampl2 * exp(-0.5 * pow(logNQ - cntr2, 2.) / (wide2 * wide2)) *
                (1. + erf(skew2 * (logNQ - cntr2) / (wide2 * kSqrt2)));
  }
  if (omega < 0.) omega = 0;
  return omega;
}
double FanoER(const nb200_detector& d, double density, double Nq_mean, double efield, const double* W) {
  double Fano = 0.12707 - 0.029623 * density - 0.0057042 * pow(density, 2.) + 0.0015957 * pow(density, 3.);
  if (!d.inGas) {
    if (W[6] < 0.)
      Fano += fabs(W[6]) * sqrt(Nq_mean) * pow(efield, 0.5);
    else
      Fano = W[6];
  }
  return Fano;
}
struct Q {
  int photons, electrons, ions, excitons;
  double recombProb, Variance;
};
Q GetQuanta(const nb200_detector& d, const Y& y, double density, const double* W, int nw, double SkewnessER) {
  Q result{};
  bool HighE;
  int Nq_actual, Ne, Nph, Ni, Nex;
  if (nw < 13) throw std::runtime_error("ERROR: You need a minimum of 13 free parameters for the resolution model.");
  double excitonRatio = y.NexONi;
  double Nq_mean = y.Nph + y.Ne;
  double elecFrac = std::max(0., std::min(y.Ne / Nq_mean, 1.));
  if (excitonRatio < 0.) {
    excitonRatio = 0.;
    HighE = true;
  } else
    HighE = false;
  double alf = 1. / (1. + excitonRatio);
  double recombProb = 1. - (excitonRatio + 1.) * elecFrac;
  if (recombProb < 0.) excitonRatio = 1. / elecFrac - 1.;
  if (nearlyEqual(y.L, 1.)) {
    if (nearlyEqual(Nq_mean, 0.))
      Nq_actual = 0;
    else {
      double Fano = FanoER(d, density, Nq_mean, y.F, W);
      Nq_actual = int(floor(rand_gauss(Nq_mean, sqrt(Fano * Nq_mean), true) + 0.5));
    }
    Ni = binom_draw(Nq_actual, alf);
    Nex = Nq_actual - Ni;
  } else {
    double Ni_mean = Nq_mean * alf;
    double Fano = W[0] + W[11] * Ni_mean;
    if (Fano < 0.) Fano = 0.;
    Ni = int(floor(rand_gauss(Ni_mean, sqrt(Fano * Ni_mean), true) + 0.5));
    double Nex_mean = Nq_mean * excitonRatio * alf;
    Fano = W[1] + W[12] * Nex_mean;
    if (Fano < 0.) Fano = 0.;
    Nex = int(floor(rand_gauss(Nex_mean, sqrt(Fano * Nex_mean), true) + 0.5));
    Nq_actual = Nex + Ni;
  }
  if (Nq_actual == 0) return result;
  if (Nex < 0) Nex = 0;
  if (Ni < 0) Ni = 0;
  if (Nex > Nq_actual) Nex = Nq_actual;
  if (Ni > Nq_actual) Ni = Nq_actual;
  result.ions = Ni;
  result.excitons = Nex;
  if (Nex <= 0 && HighE) recombProb = y.Nph / double(Ni);
  recombProb = std::max(0., std::min(recombProb, 1.));
  if (std::isnan(recombProb) || std::isnan(elecFrac) || Ni == 0 || nearlyEqual(recombProb, 0.0)) {
    result.photons = Nex;
    result.electrons = Ni;
    return result;
  }
  double omega = y.L < 1 ? RecombOmegaNR(elecFrac, W) : RecombOmegaER(y.F, elecFrac, Nq_mean, W, nw);
  double lambdaChen;
  if (W[2] < 0. && y.L < 1) lambdaChen = 1.0 + W[2]; else lambdaChen = 1.0;
  if (lambdaChen <= 0.) lambdaChen = 1E-6;
  double Variance = lambdaChen * recombProb * (1. - recombProb) * Ni + omega * omega * Ni * Ni;
  double skewness;
  if ((y.Nph + y.Ne) > 1e4 || y.F > 4e3 || y.F < 50.)
    skewness = 0.00;
  else {
    double Wq_eV, a_;
    WorkFunction(density, d.molarMass, d.OldW13eV, Wq_eV, a_);
    double engy = 1e-3 * Wq_eV * (y.Nph + y.Ne), fld = y.F;
    double alpha0 = 1.39, cc0 = 4.0, cc1 = 22.1, E0 = 7.7, E1 = 54., E2 = 26.7, E3 = 6.4, F0 = 225., F1 = 71.;
    skewness = 0.;
    if (nearlyEqual(y.L, 1.)) {
      if (SkewnessER != -999.)
        skewness = SkewnessER;
      else
        skewness = 1. / (1. + exp((engy - E2) / E3)) * (alpha0 + cc0 * exp(-1. * fld / F0) * (1. - exp(-1. * engy / E0))) +
                   1. / (1. + exp(-1. * (engy - E2) / E3)) * cc1 * exp(-1. * engy / E1) * exp(-1. * sqrt(fld) / sqrt(F1));
    } else
      skewness = W[5];
  }
  double widthCorrection = sqrt(1. - (2. / M_PI) * skewness * skewness / (1. + skewness * skewness));
  double muCorrection = (sqrt(Variance) / widthCorrection) * (skewness / sqrt(1. + skewness * skewness)) * 2. * kInvSqrt2PI;
  if (std::abs(skewness) > DBL_MIN)
    Ne = int(floor(rand_skewGauss((1. - recombProb) * Ni - muCorrection, sqrt(Variance) / widthCorrection, skewness) + 0.5));
  else
    Ne = int(floor(rand_gauss((1. - recombProb) * Ni, sqrt(Variance), true) + 0.5));
  Ne = std::min(std::max(Ne, 0), Ni);
  Nph = Nq_actual - Ne;
  if (Nph > Nq_actual) Nph = Nq_actual;
  if (Nph < Nex) Nph = Nex;
  if ((Nph + Ne) != (Nex + Ni)) throw std::runtime_error("ERROR: Quanta not conserved. Tell Matthew Immediately!");
  result.Variance = Variance;
  result.recombProb = recombProb;
  result.photons = Nph;
  result.electrons = Ne;
  return result;
}

// ---- xyResolution src/NEST.cpp:2699-2736 ------------------------------------------------------------
void xyResolution(const nb200_detector& d, double x, double y, double A_top, double* out) {
  A_top *= 1. - FitTBA1(d, x, y, d.TopDrift / 2.);
  double kappa = d.PosResBase;
  double sigmaR = kappa / sqrt(A_top);
  sigmaR = sqrt(sigmaR * sigmaR + d.PosResFlat * d.PosResFlat);
  double phi = kTwoPI * rand_uniform();
  sigmaR = std::abs(rand_gauss(0.0, sigmaR, false));
  double sigmaX = sigmaR * cos(phi), sigmaY = sigmaR * sin(phi);
  if (sigmaR > 1e2 || std::isnan(sigmaR) || sigmaR < 0. || std::abs(sigmaX) > 1e2 || std::abs(sigmaY) > 1e2) {
    if (A_top > 40.) {
      sigmaX = 0.;
      sigmaY = 0.;
    }
  }
  out[0] = x + sigmaX;
  out[1] = y + sigmaY;
}

// ---- S1: src/NEST.cpp:1288-1722, 2237-2268 -------------------------------------------------------------
double nCr(double n, double r) { return tgammal(n + 1.) / (tgammal(r + 1.) * tgammal(n - r + 1.)); }

void GetS1(const nb200_detector& d, const Q& quanta, double tx, double ty, double tz, double sx, double sy, double sz,
           double dV_mid, double energy, int mode, double* scint) {
  int Nph = quanta.photons;
  for (int i = 0; i < 9; ++i) scint[i] = 0;
  double posDep = FitS1(d, tx, ty, tz);
  double posDepSm = FitS1(d, sx, sy, sz);
  double dz_center = d.TopDrift - dV_mid * d.dtCntr;
  posDep /= FitS1(d, 0., 0., dz_center);
  posDepSm /= FitS1(d, 0., 0., dz_center);
  double g1_XYZ = std::max(0., std::min(d.g1 * posDep, 1.));
  if (nearlyEqual(d.g1, 1.)) {
    g1_XYZ = 1.;
    posDep = 1.;
    posDepSm = 1.;
  }
  if (nearlyEqual(d.g1, 0.)) {
    g1_XYZ = 0.;
    posDep = 0.;
    posDepSm = 0.;
  }
  int nHits = static_cast<int>(binom_draw(Nph, g1_XYZ));
  int Nphe = 0;
  double eff = d.sPEeff;
  if (eff < 1.) eff += ((1. - eff) / (2. * static_cast<double>(d.numPMTs))) * static_cast<double>(nHits);
  eff = std::max(0., std::min(eff, 1.));
  double pulseArea = 0., spike = 0., prob;
  int current = mode;
  if (current == NB200_S1_HYBRID) current = energy < 100. ? NB200_S1_FULL : NB200_S1_PARAMETRIC;
  double NewMean = FindNewMean(d.sPEres);
  if (current == NB200_S1_FULL) {
    for (int i = 0; i < nHits; ++i) {
      double TruncGauss = rand_zero_trunc_gauss(1. / NewMean, d.sPEres);
      double phe1 = TruncGauss + rand_gauss(d.noiseBaseline[0], d.noiseBaseline[1], false);
      ++Nphe;
      if (phe1 > DBL_MAX) phe1 = 1.;
      prob = rand_uniform();
      if (prob > eff) phe1 = 0.;
      double phe2 = 0.;
      if (rand_uniform() < d.P_dphe) {
        TruncGauss = rand_zero_trunc_gauss(1. / NewMean, d.sPEres);
        phe2 = TruncGauss + rand_gauss(d.noiseBaseline[0], d.noiseBaseline[1], false);
        ++Nphe;
        if (phe2 > DBL_MAX) phe2 = 1.;
        if (rand_uniform() > eff && prob > eff) phe2 = 0.;
      }
      if ((phe1 + phe2) > d.sPEthr && (-20. * log(rand_uniform()) < d.coinWind || nHits > d.coinLevel)) {
        ++spike;
        pulseArea += phe1 + phe2;
      }
    }
  } else {
    double a_spe = 0.5 * (1. + erf(-1. / d.sPEres / kSqrt2));
    double x_spe = 0.5 * (1. + erf((d.sPEthr - 1.) / d.sPEres / kSqrt2));
    double spe_pct = (x_spe - a_spe) / (1. - a_spe);
    double a_dpe = 0.5 * (1. + erf(-2. / (kSqrt2 * d.sPEres) / kSqrt2));
    double x_dpe = 0.5 * (1. + erf((d.sPEthr - 2.) / (kSqrt2 * d.sPEres) / kSqrt2));
    double dpe_pct = (x_dpe - a_dpe) / (1. - a_dpe);
    double below = spe_pct * (1. - d.P_dphe) + dpe_pct * d.P_dphe;
    Nphe = nHits + static_cast<int>(binom_draw(nHits, d.P_dphe));
    eff = d.sPEeff;
    if (eff < 1.) eff += ((1. - eff) / (2. * static_cast<double>(d.numPMTs))) * static_cast<double>(nHits);
    eff = std::max(0., std::min(eff, 1.));
    double Nphe_det = static_cast<double>(binom_draw(Nphe, 1. - (1. - eff) / (1. + d.P_dphe)));
    pulseArea = rand_gauss(Nphe_det * (1. / NewMean), d.sPEres * sqrt(Nphe_det / NewMean), true);
    spike = static_cast<double>(binom_draw(nHits, eff * (1. - below)));
  }
  if (d.noiseQuadratic[0] != 0)
    pulseArea = rand_gauss(pulseArea, sqrt(pow(d.noiseQuadratic[0] * pulseArea * pulseArea, 2) + pow(d.noiseLinear[0] * pulseArea, 2)), true);
  else
    pulseArea = rand_gauss(pulseArea, d.noiseLinear[0] * pulseArea, true);
  if (pulseArea < d.sPEthr) pulseArea = 0.;
  if (spike < 0) spike = 0;
  double pulseAreaC = pulseArea / posDepSm;
  double Nphd = pulseArea / (1. + d.P_dphe);
  double NphdC = pulseAreaC / (1. + d.P_dphe);
  double spikeC = spike / posDepSm;
  scint[0] = nHits; scint[1] = Nphe; scint[2] = pulseArea; scint[3] = pulseAreaC; scint[4] = Nphd;
  scint[5] = NphdC; scint[6] = spike; scint[7] = spikeC; scint[8] = spike;
  if (spike < d.coinLevel)
    prob = 0.;
  else {
    if (spike > 10.)
      prob = 1.;
    else {
      if (d.coinLevel == 0)
        prob = 1.;
      else if (d.coinLevel == 1)
        prob = (spike >= 1.) ? 1. : 0.;
      else if (d.coinLevel == 2)
        prob = 1. - pow((double)d.numPMTs, 1. - spike);
      else {
        double numer = 0., denom = 0.;
        for (int i = spike; i > 0; --i) {
          denom += nCr(d.numPMTs, i);
          if (i >= d.coinLevel) numer += nCr(d.numPMTs, i);
        }
        prob = numer / denom;
      }
    }
  }
  if (spike >= 1. && spikeC > 0.) {  // GetSpike :2243-2268
    double ns0, ns1;
    if (scint[7] > kSPIKES_MAXM) {
      ns0 = scint[4];
      ns1 = scint[5];
    } else {
      ns0 = rand_zero_trunc_gauss(scint[6], (d.sPEres / 4.) * sqrt(std::abs(scint[6])));
      ns1 = ns0 / FitS1(d, sx, sy, sz) * FitS1(d, 0., 0., d.TopDrift - dV_mid * d.dtCntr);
    }
    scint[6] = ns0;
    scint[7] = ns1;
  }
  if (rand_uniform() > prob && prob < 1.) {
    for (int i = 0; i < 8; ++i) {
      if (nearlyEqual(scint[i], 0.)) scint[i] = kPHE_MIN;
      scint[i] = -1. * std::abs(scint[i]);
    }
  }
}

// ---- S2 (Full): src/NEST.cpp:1724-1776, 1975-2063 -----------------------------------------------------
void GetS2(const nb200_detector& d, int Ne, double tx, double ty, double tz, double sx, double sy, double dt,
           double dfield, const double* g2p, double* ion) {
  double elYield = g2p[0], ExtEff = g2p[1], SE = g2p[2], g2 = g2p[3], gasGap = g2p[4];
  for (int i = 0; i < 9; ++i) ion[i] = 0;
  if (dfield < kFIELD_MIN || elYield <= 0. || ExtEff <= 0. || SE <= 0. || g2 <= 0. || gasGap <= 0.) return;
  double posDep = FitS2(d, tx, ty);
  double posDepSm = FitS2(d, sx, sy);
  posDep /= FitS2(d, 0., 0.);
  posDepSm /= FitS2(d, 0., 0.);
  if (nearlyEqual(d.g1_gas, 1.)) {
    posDep = 1.;
    posDepSm = 1.;
  }
  if (nearlyEqual(d.g1_gas, 0.)) {
    posDep = 0.;
    posDepSm = 0.;
  }
  int Nee = binom_draw(Ne, ExtEff * exp(-dt / d.eLife_us));
  uint64_t Nph = 0, nHits = 0, Nphe = 0;
  double pulseArea = 0.;
  Nph = uint64_t(floor(rand_gauss(elYield * double(Nee), sqrt(d.s2Fano * elYield * double(Nee)), true) + 0.5));
  nHits = binom_draw(Nph, d.g1_gas * posDep);
  Nphe = nHits + binom_draw(nHits, d.P_dphe);
  double NewMean = FindNewMean(d.sPEres);
  pulseArea = rand_gauss(Nphe / NewMean, d.sPEres * sqrt(Nphe / NewMean), true);
  if (d.noiseQuadratic[1] != 0)
    pulseArea = rand_gauss(pulseArea, sqrt(pow(d.noiseQuadratic[1] * pow(pulseArea, 2), 2) + pow(d.noiseLinear[1] * pulseArea, 2)), true);
  else
    pulseArea = rand_gauss(pulseArea, d.noiseLinear[1] * pulseArea, true);
  double pulseAreaC = pulseArea / exp(-dt / d.eLife_us) / posDepSm;
  double Nphd = pulseArea / (1. + d.P_dphe);
  double NphdC = pulseAreaC / (1. + d.P_dphe);
  double tba = FitTBA1(d, tx, ty, tz);
  double S2b = rand_gauss(tba * pulseArea, sqrt(tba * pulseArea * (1. - tba)), true);
  double S2bc = S2b / exp(-dt / d.eLife_us) / posDepSm;
  ion[0] = Nee; ion[1] = Nph; ion[2] = nHits; ion[3] = Nphe;
  if (d.s2_thr >= 0) {
    ion[4] = pulseArea; ion[5] = pulseAreaC; ion[6] = Nphd; ion[7] = NphdC;
  } else {
    ion[4] = S2b; ion[5] = S2bc; ion[6] = S2b / (1. + d.P_dphe); ion[7] = S2bc / (1. + d.P_dphe);
  }
  if (pulseArea < std::abs(d.s2_thr)) {
    for (int i = 0; i < 8; ++i) {
      if (nearlyEqual(ion[i], 0.)) ion[i] = kPHE_MIN;
      ion[i] = -1. * std::abs(ion[i]);
    }
  }
  ion[8] = g2;
}

// ---- CalculateG2 src/NEST.cpp:2065-2235 (the verbosity>0 branch consumes 10 000 S2 draws) ------
bool CalculateG2(const nb200_detector& d, int verbosity, double* g2p) {
  double alpha = 0.137, beta = 4.70e-18;
  double epsilonRatio = kEPS_LIQ / std::abs(kEPS_GAS);
  if (d.inGas) epsilonRatio = 1.;
  double E_liq = d.E_gas / epsilonRatio;
  double T_Kelvin = d.T_Kelvin;
  double em1 = 8.807528626640e4 - 2.026247730928e3 * T_Kelvin + 1.747197309338e1 * pow(T_Kelvin, 2.) -
               6.692362929271e-2 * pow(T_Kelvin, 3.) + 9.607626262594e-5 * pow(T_Kelvin, 4.);
  double em2 = 5.074800229635e5 - 1.460168019275e4 * T_Kelvin + 1.680089978382e2 * pow(T_Kelvin, 2.) -
               9.663183204468e-1 * pow(T_Kelvin, 3.) + 2.778229721617e-3 * pow(T_Kelvin, 4.) -
               3.194249083426e-6 * pow(T_Kelvin, 5.);
  double em3 = -4.659269964120e6 + 1.366555237249e5 * T_Kelvin - 1.602830617076e3 * pow(T_Kelvin, 2.) +
               9.397480411915e-0 * pow(T_Kelvin, 3.) - 2.754232523872e-2 * pow(T_Kelvin, 4.) +
               3.228101180588e-5 * pow(T_Kelvin, 5.);
  double ExtEff = 1. - em1 * exp(-em2 * pow(E_liq, em3));
  if (ExtEff > 1. || d.inGas) ExtEff = 1.;
  if (ExtEff < 0. || E_liq <= 0.) ExtEff = 0.;
  double gasGap = d.anode - d.TopDrift;
  if (gasGap <= 0. && E_liq > 0.) return false;
  bool YesGas = true;
  double rho = GetDensity(d.T_Kelvin, d.p_bar, YesGas, d.molarMass);
  double elYield = (alpha * d.E_gas * 1e3 - beta * (kAVO * rho / d.molarMass)) * gasGap * 0.1;
  double SE = elYield * d.g1_gas;
  if (d.s2_thr < 0) SE *= FitTBA1(d, 0., 0., d.TopDrift / 2.);
  double g2 = ExtEff * SE;
  if (verbosity > 0) {
    double StdDev = 0., muSE = 0.;
    const int numSE = 10000;
    for (int i = 0; i < numSE; ++i) {
      int Nph = int(floor(rand_gauss(elYield, sqrt(d.s2Fano * elYield), true) + 0.5));
      double phi = kTwoPI * rand_uniform();
      double r = d.radius * sqrt(rand_uniform());
      double x = r * cos(phi), y = r * sin(phi);
      double posDep = FitS2(d, x, y) / FitS2(d, 0., 0.);
      int nHits = binom_draw(Nph, d.g1_gas * posDep);
      double Nphe = nHits + binom_draw(nHits, d.P_dphe);
      double pulseArea = rand_gauss(Nphe, d.sPEres * sqrt(Nphe), true);
      if (d.noiseQuadratic[1] != 0)
        pulseArea = rand_gauss(pulseArea, sqrt(pow(d.noiseQuadratic[1] * pow(pulseArea, 2), 2) + pow(d.noiseLinear[1] * pulseArea, 2)), true);
      else
        pulseArea = rand_gauss(pulseArea, d.noiseLinear[1] * pulseArea, true);
      if (d.s2_thr < 0.) {
        double t = FitTBA1(d, 0.0, 0.0, d.TopDrift / 2.);
        pulseArea = rand_gauss(t * pulseArea, sqrt(t * pulseArea * (1. - t)), true);
      }
      double pulseAreaC = pulseArea / posDep;
      double NphdC = pulseAreaC / (1. + d.P_dphe);
      StdDev += (SE - NphdC) * (SE - NphdC);
      muSE += NphdC;
    }
    SE = muSE / double(numSE);
  }
  g2p[0] = elYield; g2p[1] = ExtEff; g2p[2] = SE; g2p[3] = g2; g2p[4] = gasGap;
  return true;
}

// ---- TestSpectra: src/TestSpectra.cpp:31-58, 110-128, 151-167, 192-211 -------------------------------
// RandomGen::VonNeumann src/RandomGen.cpp:140-157
bool VonNeumann(double xMin, double xMax, double yMin, double yMax, double& x, double& y, double f) {
  if (y > f) {
    x = xMin + (xMax - xMin) * rand_uniform();
    y = yMin + (yMax - yMin) * rand_uniform();
    return true;
  }
  return false;
}
double CH3T_spectrum(double xMin, double xMax) {
  double m_e = 510.9989461, aa = 0.0072973525664, ZZ = 2., qValue = 18.5898;
  if (xMax > qValue) xMax = qValue;
  if (xMin < 0.) xMin = 0.;
  double yMax = 1.1e7;
  double x0 = xMin + (xMax - xMin) * rand_uniform();
  double y0 = yMax * rand_uniform();
  bool again = true;
  while (again) {
    double B = sqrt(x0 * x0 + 2. * x0 * m_e) / (x0 + m_e);
    double x = (2. * M_PI * ZZ * aa) * (x0 + m_e) / sqrt(x0 * x0 + 2. * x0 * m_e);
    double FuncValue = (sqrt(2. * x0 * m_e) * (x0 + m_e) * (qValue - x0) * (qValue - x0) * x * (1. / (1. - exp(-x))) *
                        (1.002037 - 0.001427 * (B)));
    again = VonNeumann(xMin, xMax, 0., yMax, x0, y0, FuncValue);
  }
  return x0;
}
double B8_spectrum(double xMin, double xMax) {
  const double m1 = -2.6455e-10, m2 = 5.0843;
  if (xMax != 5.) xMax = 5.;
  if (xMin != 0.) xMin = 0.;
  double yMax = pow(10., 3.4155);
  double x0 = xMin + (xMax - xMin) * rand_uniform();
  double y0 = yMax * rand_uniform();
  bool again = true;
  while (again) {
    double FuncValue = 3.4155 - 1.2184 * x0 + 0.32849 * pow(x0, 2.) - 0.12441 * pow(x0, 3.) + m1 * exp(m2 * x0);
    FuncValue = pow(10., FuncValue);
    again = VonNeumann(xMin, xMax, 0., yMax, x0, y0, FuncValue);
  }
  return x0;
}
double DD_spectrum(double xMin, double xMax, double expFall, double peakFrac, double peakMu, double peakSig, double peakSkew) {
  if (xMax > 80.) xMax = 80.;
  if (xMin < 0.000) xMin = 0.000;
  double yMin = 0.0, yMax = 1.;
  if (peakMu < xMax) yMax += peakFrac;
  double x0 = xMin + (xMax - xMin) * rand_uniform();
  double y0 = yMin + (yMax - yMin) * rand_uniform();
  bool again = true;
  while (again) {
    double FuncValue = exp(-x0 / expFall) + peakFrac * exp(-pow((x0 - peakMu) / peakSig, 2.)) * (erf(peakSkew * (x0 - peakMu) / peakSig) + 1);
    again = VonNeumann(xMin, xMax, yMin, yMax, x0, y0, FuncValue);
  }
  return x0;
}
double C14_spectrum(double xMin, double xMax) {  // TestSpectra.cpp:60-108
  double m_e = 510.9989461, aa = 0.0072973525664, ZZ = 7., V0 = 0.495, qValue = 156.;
  if (xMax > qValue) xMax = qValue;
  if (xMin < 0.) xMin = 0.;
  double yMax = 2.5e9;
  double x0 = xMin + (xMax - xMin) * rand_uniform();
  double y0 = yMax * rand_uniform();
  bool again = true;
  while (again) {
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
    double x_screen = (2 * M_PI * ZZ * aa) / B_screen;
    double F_nr_screen = W_screen * p_screen / (WW * pp) * x_screen * (1 / (1 - exp(-x_screen)));
    double F_bb_screen =
        F_nr_screen * pow(W_screen * W_screen * (1 + 4 * (aa * ZZ) * (aa * ZZ)) - 1, sqrt(1 - aa * aa * ZZ * ZZ) - 1);
    again = VonNeumann(xMin, xMax, 0., yMax, x0, y0, dNdE_phasespace * F_bb_screen);
  }
  return x0;
}
double Cf_spectrum(double xMin, double xMax) {  // TestSpectra.cpp:169-190, `power` :20-21
  const double power = 3.7488;
  if (xMax > 200.) xMax = 200.;
  if (xMin < DBL_MIN) xMin = DBL_MIN;
  double yMax = 2. * pow(10., power), yMin = 0.0;
  double x0 = xMin + (xMax - xMin) * rand_uniform();
  double y0 = yMax * rand_uniform();
  bool again = true;
  while (again) {
    double l = log10(x0);
    double FuncValue = power * pow(l, 0.) - 0.77942 * pow(l, 1.) + 1.30300 * pow(l, 2.) - 2.75280 * pow(l, 3.) +
                       1.57310 * pow(l, 4.) - 0.30072 * pow(l, 5.);
    FuncValue = pow(10., FuncValue);
    FuncValue *= 1.9929 - .033214 * pow(x0, 1.) + .00032857 * pow(x0, 2.) - 1.000e-6 * pow(x0, 3.);
    again = VonNeumann(xMin, xMax, yMin, yMax, x0, y0, FuncValue);
  }
  return x0;
}
double ppSolar_spectrum(double xMin, double xMax) {  // TestSpectra.cpp:213-227
  if (xMax > 250.) xMax = 250.;
  if (xMin < 0.00) xMin = 0.00;
  double yMax = 0.000594;
  double x0 = xMin + (xMax - xMin) * rand_uniform();
  double y0 = yMax * rand_uniform();
  bool again = true;
  while (again) {
    double FuncValue = 9.2759e-6 - 3.5556e-8 * x0 - 5.4608e-12 * x0 * x0;
    again = VonNeumann(xMin, xMax, 0., yMax, x0, y0, FuncValue);
  }
  return x0;
}
double atmNu_spectrum(double xMin, double xMax) {  // TestSpectra.cpp:229-245
  if (xMax > 85.) xMax = 85.;
  if (xMin < 0.0) xMin = 0.0;
  double x0 = xMin + (xMax - xMin) * rand_uniform();
  double y0 = rand_uniform();
  bool again = true;
  while (again) {
    double FuncValue = (1. + 0.0041482 * x0 + 0.00079972 * x0 * x0 - 1.0201e-5 * x0 * x0 * x0) * exp(-x0 / 12.355);
    again = VonNeumann(xMin, xMax, 0., 1., x0, y0, FuncValue);
  }
  return x0;
}
double PowLawFit_spectrum(double xMin, double xMax, double expo) {
  double yMax, yMin;
  ++expo;
  if (expo < 0.) {
    if (xMin <= 0.) xMin = 0.1;
    yMax = pow(xMin, expo);
    yMin = pow(xMax, expo);
  } else {
    yMax = pow(xMax, expo);
    yMin = pow(xMin, expo);
  }
  double xTry = pow(yMin + (yMax - yMin) * rand_uniform(), 1. / expo);
  if (xTry < xMin) xTry = xMin;
  if (xTry > xMax) xTry = xMax;
  return xTry;
}

// execNEST.cpp:1121-1158 signal selection
void SelectSignals(const nb200_analysis& A, const double* scint, const double* scint2, double& signal1, double& signal2) {
  if (A.usePD == 0 && std::abs(scint[3]) > A.minS1 && scint[3] < A.maxS1) signal1 = scint[3];
  else if (A.usePD == 1 && std::abs(scint[5]) > A.minS1 && scint[5] < A.maxS1) signal1 = scint[5];
  else if ((A.usePD >= 2 && std::abs(scint[7]) > A.minS1 && scint[7] < A.maxS1) || A.maxS1 >= 998.5) signal1 = scint[7];
  else signal1 = -999.;
  if (A.usePD == 0 && std::abs(scint2[5]) > A.minS2 && scint2[5] < A.maxS2) signal2 = scint2[5];
  else if (A.usePD >= 1 && std::abs(scint2[7]) > A.minS2 && scint2[7] < A.maxS2) signal2 = scint2[7];
  else signal2 = -999.;
}

const nb200_detector& D(const nb200_detector* d) { return *d; }

}  // namespace

extern "C" {

int orc_chain_ncol(void) { return ORC_NCOL; }

double orc_density(double T, double P, int* inGas) {
  bool g = false;
  double r = GetDensity(T, P, g, 131.293);
  if (inGas) *inGas = g;
  return r;
}
void orc_work_function(double rho, double molarMass, int OldW13eV, double* Wq_eV, double* alpha) {
  WorkFunction(rho, molarMass, OldW13eV != 0, *Wq_eV, *alpha);
}
double orc_drift_velocity_liquid(double T, double F) { return DriftVelocityLiquid(T, F); }
double orc_find_new_mean(double sigma) { return FindNewMean(sigma); }

// what execNEST evaluates before its loop (execNEST.cpp:612-631, 878-890), CalculateG2(0)
int orc_prepare(const nb200_detector* det, double inField, nb200_run_consts* out) {
  std::memset(out, 0, sizeof *out);
  bool g = false;
  out->rho = GetDensity(det->T_Kelvin, det->p_bar, g, det->molarMass);
  WorkFunction(out->rho, det->molarMass, det->OldW13eV, out->Wq_eV, out->alpha);
  if (!det->OldW13eV) out->Wq_eV = 11.5;
  nb200_detector d = *det;
  d.inGas = g;
  if (!CalculateG2(d, 0, out->g2_params)) return 1;
  out->NewMean = FindNewMean(det->sPEres);
  out->field = inField;
  if (inField == -1.) {
    std::vector<double> t = DriftTableNonUniform(d, 0.1, 0., 0.);
    double centralZ = (d.gate * 0.8 + d.cathode * 1.03) / 2.;
    out->vD_middle = t[int(floor(centralZ / 0.1 + 0.5))];
    out->vD = out->vD_middle;
  } else {
    out->vD_middle = DriftVelocityLiquid(det->T_Kelvin, inField);
    out->vD = out->vD_middle;
  }
  return 0;
}

void orc_maps(const nb200_detector* det, int n, const double* x, const double* y, const double* z, double* s1,
              double* s2, double* ef, double* tba) {
  for (int i = 0; i < n; ++i) {
    s1[i] = FitS1(D(det), x[i], y[i], z[i]);
    s2[i] = FitS2(D(det), x[i], y[i]);
    ef[i] = FitEF(D(det), x[i], y[i], z[i]);
    tba[i] = FitTBA1(D(det), x[i], y[i], z[i]);
  }
}

int orc_yields(const nb200_detector* det, int which, int species, int n, const double* E, const double* field,
               double rho, double massNum, double atomNum, const double* nrpar, const double* erpar,
               double multFact, double* out) {
  try {
    for (int i = 0; i < n; ++i) {
      Y y;
      if (which == 1) y = GetYieldBetaGR(D(det), E[i], rho, field[i], erpar, multFact);
      else if (which == 2) y = GetYieldGamma(D(det), E[i], rho, field[i], multFact);
      else if (which == 3) y = GetYieldERWeighted(D(det), E[i], rho, field[i], erpar);
      else y = GetYields(D(det), species, E[i], rho, field[i], massNum, atomNum, nrpar, erpar);
      double* o = out + 6 * size_t(i);
      o[0] = y.Nph; o[1] = y.Ne; o[2] = y.NexONi; o[3] = y.L; o[4] = y.F; o[5] = y.dT;
    }
  } catch (std::exception&) {
    return 1;
  }
  return 0;
}

void orc_widths(const nb200_detector* det, int n, const double* efield, const double* elecFrac, const double* nq,
                double rho, const double* widths, int nw, double* out) {
  for (int i = 0; i < n; ++i) {
    out[3 * i + 0] = RecombOmegaNR(elecFrac[i], widths);
    out[3 * i + 1] = RecombOmegaER(efield[i], elecFrac[i], nq[i], widths, nw);
    out[3 * i + 2] = FanoER(D(det), rho, nq[i], efield[i], widths);
  }
}

int orc_quanta(const nb200_detector* det, uint64_t seed, int n, const double* y6, double rho, const double* widths,
               int nw, double skewER, int* iout, double* dout) {
  G.seed(seed);
  Y y{y6[0], y6[1], y6[2], y6[3], y6[4], y6[5]};
  try {
    for (int i = 0; i < n; ++i) {
      Q q = GetQuanta(D(det), y, rho, widths, nw, skewER);
      iout[4 * i] = q.photons; iout[4 * i + 1] = q.electrons; iout[4 * i + 2] = q.ions; iout[4 * i + 3] = q.excitons;
      dout[2 * i] = q.recombProb; dout[2 * i + 1] = q.Variance;
    }
  } catch (std::exception&) {
    return 1;
  }
  return 0;
}

void orc_sampler(int kind, uint64_t seed, int n, double p0, double p1, double p2, double* out) {
  G.seed(seed);
  for (int i = 0; i < n; ++i) {
    switch (kind) {
      case 0: out[i] = rand_uniform(); break;
      case 1: out[i] = rand_gauss(p0, p1, p2 != 0.); break;
      case 2: out[i] = rand_zero_trunc_gauss(p0, p1); break;
      case 3: out[i] = rand_skewGauss(p0, p1, p2); break;
      case 4: out[i] = double(binom_draw(int64_t(p0), p1)); break;
      case 5: out[i] = rand_exponential(p0, p1, p2); break;
      case 6: out[i] = SelectRanXeAtom(); break;
      default: out[i] = double(poisson_draw(p0)); break;
    }
  }
}

void orc_spectrum(int kind, uint64_t seed, int n, double eMin, double eMax, double* out) {
  G.seed(seed);
  for (int i = 0; i < n; ++i) {
    if (kind == NB200_SRC_CH3T) out[i] = CH3T_spectrum(eMin, eMax);
    else if (kind == NB200_SRC_DD) out[i] = DD_spectrum(eMin, eMax, 13., 0.12, 71.2, 20., -20.5);
    else if (kind == NB200_SRC_AMBE) out[i] = PowLawFit_spectrum(eMin, eMax, -0.95351);
    else if (kind == NB200_SRC_C14) out[i] = C14_spectrum(eMin, eMax);
    else if (kind == NB200_SRC_CF) out[i] = Cf_spectrum(eMin, eMax);
    else if (kind == NB200_SRC_PPSOLAR) out[i] = ppSolar_spectrum(eMin, eMax);
    else if (kind == NB200_SRC_ATMNU) out[i] = atmNu_spectrum(eMin, eMax);
    else out[i] = B8_spectrum(eMin, eMax);
  }
}

void orc_s1(const nb200_detector* det, uint64_t seed, int n, int Nph, const double* pos, const double* spos, double vD,
            double vDmid, int species, double field, double energy, int mode, double* out) {
  (void)vD; (void)species; (void)field;
  G.seed(seed);
  Q q{};
  q.photons = Nph;
  for (int i = 0; i < n; ++i)
    GetS1(D(det), q, pos[0], pos[1], pos[2], spos[0], spos[1], spos[2], vDmid, energy, mode, out + 9 * size_t(i));
}

void orc_s2(const nb200_detector* det, uint64_t seed, int n, int Ne, const double* pos, const double* spos, double dt,
            double vD, double field, double* out) {
  (void)vD;
  G.seed(seed);
  double g2p[5];
  CalculateG2(D(det), 0, g2p);
  for (int i = 0; i < n; ++i)
    GetS2(D(det), Ne, pos[0], pos[1], pos[2], spos[0], spos[1], dt, field, g2p, out + 9 * size_t(i));
}

void orc_xyres(const nb200_detector* det, uint64_t seed, int n, double x, double y, double A_top, double* out) {
  G.seed(seed);
  for (int i = 0; i < n; ++i) xyResolution(D(det), x, y, A_top, out + 2 * size_t(i));
}

// The loop body of execNEST() src/execNEST.cpp:676-1250 (energy-based species)
int orc_chain(const nb200_detector* det, const nb200_analysis* ana, uint64_t seed, int n, int species, int src_kind,
              double eMin, double eMax, double inField, int fixed_pos, const double* xyz, int er_model,
              double multFact, const double* nrpar, const double* erpar, const double* widths, int nw, double skewER,
              int g2_verbosity, double* out) {
  nb200_detector d = *det;
  const nb200_analysis& A = *ana;
  G.seed(seed);
  bool inGas = false;
  double rho = GetDensity(d.T_Kelvin, d.p_bar, inGas, d.molarMass);  // :612
  d.inGas = inGas;
  if (rho < 1.75) d.inGas = 1;
  double Wq_eV, alpha_;
  WorkFunction(rho, d.molarMass, d.OldW13eV, Wq_eV, alpha_);  // :621-627
  if (!d.OldW13eV) Wq_eV = 11.5;
  double g2p[5];
  if (!CalculateG2(d, g2_verbosity, g2p)) return 1;  // :631
  double g2 = std::abs(g2p[3]), g1 = d.g1;
  double centralZ = (d.gate * 0.8 + d.cathode * 1.03) / 2.;
  double massNum = d.molarMass, atomNum = 0;
  if (species == NB200_ion) { massNum = 4; atomNum = 2; }  // "alpha" :498-500
  if (species < 6) massNum = 0;  // :670
  if (species == NB200_Kr83m) { massNum = eMax; atomNum = 0.; }  // maxTimeSep :583-588,671; minTimeSeparation :38,1052
  std::vector<double> W(widths, widths + nw);
  if (species == NB200_ion) W = {1.00, 1.00, 0., 0.50, 0.19, 0., 1., 0.046452, 0.45, 0.205, -0.2, 0., 0.};  // :1056-1063
  double vD = 0, vD_middle = 0, pos_x = 0, pos_y = 0, pos_z = 0, field, driftTime;
  std::vector<double> vTable;
  try {
    for (int j = 0; j < n; ++j) {
      double keV;
      if (src_kind == NB200_SRC_MONO || species == NB200_Kr83m) keV = eMin;  // :681-684
      else if (src_kind == NB200_SRC_CH3T) keV = CH3T_spectrum(eMin, eMax);
      else if (src_kind == NB200_SRC_B8) keV = B8_spectrum(eMin, eMax);
      else if (src_kind == NB200_SRC_DD) keV = DD_spectrum(eMin, eMax, 13., 0.12, 71.2, 20., -20.5);
      else if (src_kind == NB200_SRC_AMBE) keV = PowLawFit_spectrum(eMin, eMax, -0.95351);
      else if (src_kind == NB200_SRC_C14) keV = C14_spectrum(eMin, eMax);
      else if (src_kind == NB200_SRC_CF) keV = Cf_spectrum(eMin, eMax);
      else if (src_kind == NB200_SRC_PPSOLAR) keV = ppSolar_spectrum(eMin, eMax);
      else if (src_kind == NB200_SRC_ATMNU) keV = atmNu_spectrum(eMin, eMax);
      else if (src_kind == NB200_SRC_GAUSS) keV = rand_zero_trunc_gauss(eMin, eMax);
      else if (src_kind == NB200_SRC_EXP) keV = rand_exponential(-eMax * log(2.), 0., eMin);
      else keV = eMin + (eMax - eMin) * rand_uniform();
      if (species != NB200_B8 && species != NB200_Kr83m && species != NB200_ppSolar && species != NB200_atmNu &&
          eMax > 0. && eMin < eMax) {  // :786-790
        if (keV > eMax) keV = eMax;
        if (keV < eMin) keV = eMin;
      }
      double keV_true = keV;
      for (;;) {  // Z_NEW :806-967
        if (fixed_pos == 0 || fixed_pos == 2) pos_z = 0. + (d.TopDrift - 0.) * rand_uniform();
        else pos_z = xyz[2];
        if (fixed_pos == 0 || fixed_pos == 3) {
          double r = d.radius * sqrt(rand_uniform());
          double phi = 2. * M_PI * rand_uniform();
          pos_x = r * cos(phi);
          pos_y = r * sin(phi);
        } else {
          pos_x = xyz[0];
          pos_y = xyz[1];
        }
        if (inField == -1.) field = FitEF(d, pos_x, pos_y, pos_z);
        else field = inField;
        if (j == 0 && vD_middle == 0.) {  // :878-890
          if (inField == -1.) {
            vTable = DriftTableNonUniform(d, A.z_step, pos_x, pos_y);
            vD_middle = vTable[int(floor(centralZ / A.z_step + 0.5))];
          } else {
            vD_middle = DriftVelocityLiquid(d.T_Kelvin, inField);
            vD = DriftVelocityLiquid(d.T_Kelvin, field);
          }
        }
        if (inField == -1.) vD = vTable[int(floor(pos_z / A.z_step + 0.5))];
        driftTime = (d.TopDrift - pos_z) / vD;
        if ((driftTime > d.dt_max || driftTime < d.dt_min) && (fixed_pos == 0 || fixed_pos == 2) && field >= kFIELD_MIN)
          continue;
        break;
      }
      if (pos_z <= 0) pos_z = A.z_step;  // :975-992
      if ((pos_z > (d.TopDrift + A.z_step) || driftTime < 0.0) && field >= kFIELD_MIN) {
        driftTime = 0.0;
        pos_z = d.TopDrift - A.z_step;
      }
      Y yields{0., 0., 0., 0., 0., 0.};
      Q quanta{};
      if (keV > 0.001 * Wq_eV) {  // :1031-1077
        if (er_model == 0) yields = GetYieldBetaGR(d, keV, rho, field, erpar, multFact);
        else if (er_model == 1) yields = GetYieldGamma(d, keV, rho, field, -multFact);
        else if (er_model == 2) yields = GetYieldERWeighted(d, keV, rho, field, erpar);
        else yields = GetYields(d, species, keV, rho, field, massNum, atomNum, nrpar, erpar);
        quanta = GetQuanta(d, yields, rho, W.data(), int(W.size()), skewER);
      }
      if (d.noiseBaseline[2] != 0. || d.noiseBaseline[3] != 0.)  // :1080-1086
        quanta.electrons += int(floor(rand_gauss(d.noiseBaseline[2], d.noiseBaseline[3], false) + 0.5));
      double smear[2] = {pos_x, pos_y};
      double Nphd_S2 = g2 * quanta.electrons * exp(-driftTime / d.eLife_us);
      if (!A.MCtruthPos && Nphd_S2 > kPHE_MIN) xyResolution(d, pos_x, pos_y, Nphd_S2, smear);  // :1090-1099
      double scint[9], scint2[9];
      GetS1(d, quanta, pos_x, pos_y, pos_z, smear[0], smear[1], pos_z, vD_middle, keV, A.s1mode, scint);
      if (pos_z < d.cathode) quanta.electrons = 0;  // :1107
      GetS2(d, quanta.electrons, pos_x, pos_y, pos_z, smear[0], smear[1], driftTime, field, g2p, scint2);
      double signal1, signal2;
      SelectSignals(A, scint, scint2, signal1, signal2);
      double Nph = 0.0, Ne = 0.0;
      if (!A.MCtruthE) {  // :1170-1244
        double MultFact = 1., eff = d.sPEeff;
        if (A.s1mode != NB200_S1_PARAMETRIC) {
          if (d.sPEthr >= 0. && d.sPEres > 0. && eff > 0.) {
            MultFact = 0.5 * (1. + erf((d.sPEthr - 1.) / (d.sPEres * sqrt(2.)))) / (d.sPEres * sqrt(2. * M_PI));
            if (A.usePD == 2 && scint[7] < kSPIKES_MAXM) MultFact = 1.04;
            else MultFact = 1. + MultFact * d.P_dphe;
            if (eff < 1.) eff += ((1. - eff) / (2. * double(d.numPMTs))) * scint[0];
            eff = std::max(0., std::min(eff, 1.));
            MultFact /= eff;
          } else
            MultFact = 1.;
        }
        if (A.usePD == 0) Nph = std::abs(scint[3]) / (g1 * (1. + d.P_dphe));
        else if (A.usePD == 1) Nph = std::abs(scint[5]) / g1;
        else Nph = std::abs(scint[7]) / g1;
        if (A.usePD == 0) Ne = std::abs(scint2[5]) / (g2 * (1. + d.P_dphe));
        else Ne = std::abs(scint2[7]) / g2;
        Nph *= MultFact;
        if (signal1 <= 0.) Nph = 0.;
        if (signal2 <= 0.) Ne = 0.;
        if (d.coinLevel <= 0 && Nph <= kPHE_MIN) Nph = DBL_MIN;
        if (yields.L > DBL_MIN && Nph > 0. && Ne > 0.) {
          if (nearlyEqual(yields.L, 1.))
            keV = (Nph / 1. + Ne / 1.) * Wq_eV * 1e-3;
          else if (species <= NB200_Cf) {
            keV = pow((Ne + Nph) / nrpar[0], 1. / nrpar[1]);
            Ne *= 1. - 1. / pow(1. + pow((keV / nrpar[5]), nrpar[6]), nrpar[10]);
            Nph *= 1. - 1. / pow(1. + pow((keV / nrpar[7]), nrpar[8]), nrpar[11]);
            keV = pow((Ne + Nph) / nrpar[0], 1. / nrpar[1]);
          } else {
            const double logden = log10(rho);
            const double Wq_eV_alpha = 28.259 + 25.667 * logden - 33.611 * pow(logden, 2.) - 123.73 * pow(logden, 3.) -
                                       136.47 * pow(logden, 4.) - 74.194 * pow(logden, 5.) - 20.276 * pow(logden, 6.) -
                                       2.2352 * pow(logden, 7.);
            keV = (Nph + Ne) * Wq_eV_alpha * 1e-3 / yields.L;
          }
        } else
          keV = 0.;
      }
      double signalE;
      if ((signal1 <= 0. || signal2 <= 0.) && field >= kFIELD_MIN && d.coinLevel != 0) signalE = 0.;  // :1245-1250
      else signalE = keV;
      double* o = out + size_t(j) * ORC_NCOL;
      o[ORC_KEV_TRUE] = keV_true; o[ORC_KEV_OUT] = keV; o[ORC_FIELD] = field; o[ORC_DT] = driftTime;
      o[ORC_X] = pos_x; o[ORC_Y] = pos_y; o[ORC_Z] = pos_z; o[ORC_SX] = smear[0]; o[ORC_SY] = smear[1];
      o[ORC_NPH] = quanta.photons; o[ORC_NE] = quanta.electrons; o[ORC_NI] = quanta.ions; o[ORC_NEX] = quanta.excitons;
      o[ORC_YPH] = yields.Nph; o[ORC_YNE] = yields.Ne;
      for (int k = 0; k < 9; ++k) {
        o[ORC_S1_0 + k] = scint[k];
        o[ORC_S2_0 + k] = scint2[k];
      }
      o[ORC_SIG1] = signal1; o[ORC_SIG2] = signal2; o[ORC_SIGE] = signalE;
    }
  } catch (std::exception&) {
    return 1;
  }
  return 0;
}

// GetBand src/execNEST.cpp:1508-1621 with resol == false; out [numBins][7]
int orc_getband(const nb200_analysis* ana, int nFold, int n, const double* S1s, const double* S2s, double* out) {
  const nb200_analysis& A = *ana;
  const int numBins = A.numBins;
  if (numBins > 1000) return -1;
  std::vector<std::vector<double>> signals(numBins, std::vector<double>(1, -999.));
  std::vector<double> band(size_t(numBins) * 7, 0.);
  std::vector<uint64_t> reject(numBins, 0);
  double binWidth, border;
  if (A.useS2 == 2) {
    binWidth = (A.maxS2 - A.minS2) / double(numBins);
    border = A.minS2;
  } else {
    binWidth = (A.maxS1 - A.minS1) / double(numBins);
    border = A.minS1;
  }
  double ReqS1 = 0.;
  if (nFold == 0) ReqS1 = -DBL_MAX;
  for (int i = 0; i < n; ++i) {
    for (int j = 0; j < numBins; ++j) {
      double s1c = border + binWidth / 2. + double(j) * binWidth;
      if (i == 0) band[7 * j + 0] = s1c;
      if ((std::abs(S1s[i]) > (s1c - binWidth / 2.) && std::abs(S1s[i]) <= (s1c + binWidth / 2.)) || !nFold) {
        if (S1s[i] >= ReqS1 && S2s[i] >= 0.) {
          double v;
          if (A.useS2 == 0) v = (S1s[i] && S2s[i] && log10(S2s[i] / S1s[i]) > A.logMin && log10(S2s[i] / S1s[i]) < A.logMax) ? log10(S2s[i] / S1s[i]) : 0.;
          else if (A.useS2 == 1) v = (S1s[i] && S2s[i] && log10(S2s[i]) > A.logMin && log10(S2s[i]) < A.logMax) ? log10(S2s[i]) : 0.;
          else v = (S1s[i] && S2s[i] && log10(S1s[i] / S2s[i]) > A.logMin && log10(S1s[i] / S2s[i]) < A.logMax) ? log10(S1s[i] / S2s[i]) : 0.;
          signals[j].push_back(v);
          band[7 * j + 2] += signals[j].back();
          band[7 * j + 1] += S1s[i];
        } else
          ++reject[j];
        break;
      }
    }
  }
  for (int j = 0; j < numBins; ++j) {
    double* b = &band[7 * size_t(j)];
    if (b[0] <= 0.) b[0] = border + binWidth / 2. + double(j) * binWidth;
    signals[j].erase(signals[j].begin());
    double numPts = (double)signals[j].size();
    b[1] /= numPts;
    b[2] /= numPts;
    for (int i = 0; i < (int)numPts; ++i)
      if (signals[j][i] != -999.) b[3] += pow(signals[j][i] - b[2], 2.);
    b[3] /= numPts - 1.;
    b[3] = sqrt(b[3]);
    b[4] = b[3] / sqrt(numPts);
    b[5] = numPts / (numPts + double(reject[j]));
    for (int i = 0; i < (int)numPts; ++i)
      if (signals[j][i] != -999.) b[6] += pow((signals[j][i] - b[2]) / b[3], 3.);
    b[6] /= (numPts - 2.);
  }
  std::memcpy(out, band.data(), sizeof(double) * band.size());
  return numBins;
}

// runNESTvec src/execNEST.cpp:329-409 (FullCalculation NEST.cpp:22-42 with SkewnessER = -999);
// out [n][26] as oracle/ref_probe.cpp::ref_runvec
int orc_runvec(const nb200_detector* det, int species, int n, const double* E, const double* xyz, double inField,
               int seed, const double* erpar, const double* nrpar, const double* widths, int nw, int s1mode,
               double* out) {
  nb200_detector d = *det;
  G.seed(seed);
  double g2p[5];
  if (!CalculateG2(d, 0, g2p)) return 1;
  bool inGas = false;
  double rho = GetDensity(d.T_Kelvin, d.p_bar, inGas, d.molarMass);
  d.inGas = inGas;
  try {
    for (int i = 0; i < n; ++i) {
      const double* p = xyz + 3 * size_t(i);
      double useField = inField < 0.0 ? FitEF(d, p[0], p[1], p[2]) : inField;
      if (rho < 1.) d.inGas = 1;
      Y y = GetYields(d, species, E[i], rho, useField, d.molarMass, kATOM, nrpar, erpar);
      Q q = GetQuanta(d, y, rho, widths, nw, -999.);
      double vD = DriftVelocityLiquid(d.T_Kelvin, useField);
      double driftTime = (d.TopDrift - p[2]) / vD;
      double s1[9], s2[9];
      GetS1(d, q, p[0], p[1], p[2], p[0], p[1], p[2], vD, E[i], s1mode, s1);
      GetS2(d, q.electrons, p[0], p[1], p[2], p[0], p[1], driftTime, useField, g2p, s2);
      double* o = out + 26 * size_t(i);
      o[0] = E[i]; o[1] = p[0]; o[2] = p[1]; o[3] = p[2]; o[4] = q.electrons; o[5] = q.photons;
      o[6] = std::abs(int(s1[0])); o[7] = std::abs(int(s1[1])); o[8] = std::abs(int(s1[8]));
      o[9] = std::abs(s1[2]); o[10] = std::abs(s1[3]); o[11] = std::abs(s1[4]); o[12] = std::abs(s1[5]);
      o[13] = std::abs(s1[6]); o[14] = std::abs(s1[7]);
      o[15] = std::abs(int(s2[0])); o[16] = std::abs(int(s2[1])); o[17] = std::abs(int(s2[2])); o[18] = std::abs(int(s2[3]));
      o[19] = std::abs(s2[4]); o[20] = std::abs(s2[5]); o[21] = std::abs(s2[6]); o[22] = std::abs(s2[7]);
      o[23] = 0; o[24] = 0; o[25] = 0;
    }
  } catch (std::exception&) {
    return 1;
  }
  return 0;
}

// LArNEST::FullCalculation src/LArNEST.cpp:17-212, 363-406; parameters include/NEST/LArParameters.hh:39-128
void orc_lar(uint64_t seed, int species, int n, const double* E, const double* field, double density, double* yields,
             double* fluct) {
  G.seed(seed);
  const double WQ = 19.5, NexOverNion0 = 0.21;
  for (int i = 0; i < n; ++i) {
    double energy = E[i], efield = field[i];
    double Total, Ye, Yph;
    if (species == 0) {
      const double alpha = 11.10, beta = 0.087, gamma = 0.1, delta = -0.0932, epsilon = 2.998, zeta = 0.3, eta = 2.94;
      Total = alpha * pow(energy, beta);
      Ye = (1.0 / (gamma * pow(efield, delta))) * (1.0 / sqrt(energy + epsilon)) * (1.0 - 1.0 / (1.0 + pow(energy / zeta, eta)));
      if (Ye < 0.0) Ye = 0.0;
      Yph = alpha * pow(energy, beta) - (1.0 / (gamma * pow(efield, delta))) * (1.0 / sqrt(energy + epsilon));
      if (Yph < 0.0) Yph = 0.0;
    } else if (species == 1) {
      Total = (energy * 1.0e3 / WQ);
      double a = 32.988 + -552.988 / (17.2346 + pow(efield / (-4.7 + 0.025115 * exp(density / 0.265360653)), 0.242671));
      double b = 0.778482 + 25.9 * pow(1.105 + pow(efield / 0.4, 4.55), -7.502);
      double g = 0.659509 * (1000 / WQ + 6.5 * (5.0 + -0.5 / pow(efield / 1047.408, 0.01851)));
      double db = 1052.264 + (14159350000 - 1652.264) / (-5.0 + pow(efield / 0.157933, 1.83894));
      Ye = a * b + (g - a * b) / pow(1.0 + 10.304 * pow(energy + 0.5, 13.0654), 0.10535) + 15.7489 / (0.7 + db * pow(energy, -2.07763));
      if (Ye < 0.0) Ye = 0.0;
      Yph = Total / energy - Ye;
      if (Yph < 0.0) Yph = 0.0;
    } else {
      const double eA = 1.0 / 6200.0, eB = 64478398.7663, eC = 0.173553719, eD = 1.21, eE = 0.02852, eF = 0.01, eG = 4.71598,
                   eH = 7.72848, eI = -0.109802, eJ = 3.0;
      double fieldTerm = eF * pow((eG + pow(efield, eH)), eI);
      Ye = eA * (eB - (eB * eC + (eB / eD) * (1 - (eE * log(1 + (eB / eD) * (fieldTerm / eJ)) / ((eB / eD) * fieldTerm)))));
      if (Ye < 0.0) Ye = 0.0;
      const double pA = 1.16, pB = -0.012, pC = 1.0 / 6500.0, pD = 278037.250283, pE = 0.173553719, pF = 1.21, pG = 2,
                   pH = 0.653503, pI = 4.98483, pJ = 10.0822, pK = 1.2076, pL = -0.97977, pM = 3.0;
      double quench = 1.0 / (pA * pow(efield, pB));
      double ft = pH * pow((pI + pow((efield / pJ), pK)), pL);
      Yph = quench * pC * (pD * pE + (pD / pF) * (1 - (pG * log(1 + (pD / pF) * (ft / pM)) / ((pD / pF) * ft))));
      if (Yph < 0.0) Yph = 0.0;
      Total = (Ye + Yph) * energy;
    }
    double Ne = Ye * energy, Nph = Yph * energy, Nq = Ne + Nph;
    double Nion = Nq / (1 + NexOverNion0);
    double Nex = Nq - Nion;
    double Lindhard = (Total / energy) * WQ * 1e-3;
    if (yields) {
      double* y = yields + 9 * size_t(i);
      y[0] = Total; y[1] = Ye; y[2] = Yph; y[3] = Nph; y[4] = Ne; y[5] = Nex; y[6] = Nion; y[7] = Lindhard; y[8] = efield;
    }
    if (fluct) {  // GetDefaultFluctuations :363-406
      double Fano = 0.1115;
      double NexOverNion = Nex / Nion;
      double p_r = (Nph - Nex) / Nion;
      double alf = 1. / (1. + NexOverNion);
      double fNq = 0, fNex = 0, fNion = 0;
      double Nq_mean = Ne + Nph;
      if (Lindhard == 1.) {
        fNq = floor(rand_gauss(Nq_mean, sqrt(Fano * Nq_mean), true) + 0.5);
        if (fNq < 0.) fNq = 0.;
        fNion = double(binom_draw(fNq, alf));
        if (fNion > fNq) fNion = fNq;
        fNex = fNq - fNion;
      } else {
        fNion = floor(rand_gauss(Nq_mean * alf, sqrt(Fano * Nq_mean * alf), true) + 0.5);
        fNex = floor(rand_gauss(Nq_mean * alf * NexOverNion, sqrt(Fano * Nq_mean * alf * NexOverNion), true) + 0.5);
        if (fNex < 0.) fNex = 0.;
        if (fNion < 0.) fNion = 0.;
        fNq = fNion + fNex;
      }
      double* f = fluct + 4 * size_t(i);
      f[0] = fNex + p_r * fNion; f[1] = (1. - p_r) * fNion; f[2] = fNex; f[3] = fNion;
    }
  }
}

}  // extern "C"
