// nb_host.h -- per-run constants computed once on the host (plain C++, glibc libm).
//
// Restates NESTcalc::GetDensity (src/NEST.cpp:2278-2350), WorkFunction (:2829-2848),
// GetDriftVelocity_Liquid (:2366-2521), SetDriftVelocity_NonUniform (:2664-2697),
// CalculateG2 with verbosity 0 (:2065-2235), RandomGen::FindNewMean (RandomGen.cpp:63-70)
// and TestSpectra::WIMP_dRate / WIMP_prep_spectrum (TestSpectra.cpp:251-454), i.e. everything
// execNEST() evaluates before its event loop (execNEST.cpp:612-631, 878-890).
#pragma once
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "nb_yields.cuh"

namespace nb {
namespace host {

static const double kRidealGas = 8.31446261815324, kRealGasA = 0.4250, kRealGasB = 5.105e-5;
static const double kEpsGas = 1.00126, kEpsLiq = 1.85;

// GetDensity, xenon branch. err: warning text (solid / gas phase) as the reference prints to cerr
inline double density(double Kelvin, double bara, bool& inGas, double molarMass, std::string* warn) {
  if (Kelvin < 161.40) {
    if (warn) *warn = "WARNING: SOLID PHASE. IS THAT WHAT YOU WANTED?";
    return 3.41;
  }
  double VaporP_bar;
  if (Kelvin < 289.7)
    VaporP_bar = pow(10., 4.0519 - 667.16 / Kelvin);
  else
    VaporP_bar = DBL_MAX;
  if (bara < VaporP_bar || inGas) {
    double p_Pa = bara * 1e5;
    double d = 1. / (pow(kRidealGas * Kelvin, 3.) /
                         (p_Pa * pow(kRidealGas * Kelvin, 2.) + kRealGasA * p_Pa * p_Pa) +
                     kRealGasB);
    d *= molarMass * 1e-6;
    if (bara < VaporP_bar && warn) *warn = "WARNING: GAS PHASE. IS THAT WHAT YOU WANTED?";
    inGas = true;
    return d;
  }
  inGas = false;
  return 2.9970938084691329E+02 * exp(-8.2598864714323525E-02 * Kelvin) -
         1.8801286589442915E+06 *
             exp(-pow((Kelvin - 4.0820251276172212E+02) / 2.7863170223154846E+01, 2.)) -
         5.4964506351743057E+03 *
             exp(-pow((Kelvin - 6.3688597345042672E+02) / 1.1225818853661815E+02, 2.)) +
         8.3450538370682614E+02 *
             exp(-pow((Kelvin + 4.8840568924597342E+01) / 7.3804147172071107E+03, 2.)) -
         8.3086310405942265E+02;
}

// polyExp / Temperatures tables of GetDriftVelocity_Liquid NEST.cpp:2429-2445
static const double kPolyExp[11][7] = {
    {-3.1046, 27.037, -2.1668, 193.27, -4.8024, 646.04, 9.2471},
    {-2.7394, 22.760, -1.7775, 222.72, -5.0836, 724.98, 8.7189},
    {-2.3646, 164.91, -1.6984, 21.473, -4.4752, 1202.2, 7.9744},
    {-1.8097, 235.65, -1.7621, 36.855, -3.5925, 1356.2, 6.7865},
    {-1.5000, 37.021, -1.1430, 6.4590, -4.0337, 855.43, 5.4238},
    {-1.4939, 47.879, 0.12608, 8.9095, -1.3480, 1310.9, 2.7598},
    {-1.5389, 26.602, -.44589, 196.08, -1.1516, 1810.8, 2.8912},
    {-1.5000, 28.510, -.21948, 183.49, -1.4320, 1652.9, 2.884},
    {-1.1781, 49.072, -1.3008, 3438.4, -.14817, 312.12, 2.8049},
    {1.2466, 85.975, -.88005, 918.57, -3.0085, 27.568, 2.3823},
    {334.60, 37.556, 0.92211, 345.27, -338.00, 37.346, 1.9834}};
static const double kDriftT[11] = {100., 120., 140., 155., 157., 163., 165., 167., 184., 200., 230.};

// temperature bracket of GetDriftVelocity_Liquid (:2447-2480). Returns the lower row or -1
// where the reference throws; Kelvin is clamped to [100,230] as the reference does.
inline int drift_bracket(double& Kelvin) {
  if (Kelvin < 100.) Kelvin = 100.;
  if (Kelvin > 230.) Kelvin = 230.;
  for (int k = 0; k < 9; ++k)
    if (Kelvin >= kDriftT[k] && Kelvin < kDriftT[k + 1]) return k;
  if (Kelvin >= kDriftT[9] && Kelvin <= kDriftT[10]) return 9;
  return -1;
}

// GetDriftVelocity_Liquid (xenon, StdDev = -1). ok=false where the reference throws.
inline double drift_velocity_liquid(double Kelvin, double eField, bool* ok) {
  if (ok) *ok = true;
  int i = drift_bracket(Kelvin);
  if (i < 0) {
    if (ok) *ok = false;
    return 0.;
  }
  const double(*polyExp)[7] = kPolyExp;
  int j = i + 1;
  double Ti = kDriftT[i], Tf = kDriftT[j];
  double vi = polyExp[i][0] * exp(-eField / polyExp[i][1]) + polyExp[i][2] * exp(-eField / polyExp[i][3]) +
              polyExp[i][4] * exp(-eField / polyExp[i][5]) + polyExp[i][6];
  double vf = polyExp[j][0] * exp(-eField / polyExp[j][1]) + polyExp[j][2] * exp(-eField / polyExp[j][3]) +
              polyExp[j][4] * exp(-eField / polyExp[j][5]) + polyExp[j][6];
  if (nearly_equal(Kelvin, Ti)) return vi;
  if (nearly_equal(Kelvin, Tf)) return vf;
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

// FindNewMean RandomGen.cpp:63-70
inline double find_new_mean(double sigma) {
  double a = -1. / sigma;
  double phi = exp(-0.5 * a * a) / sqrt(2. * M_PI);
  double Phi = 0.5 * (1. + erf(a / sqrt(2.)));
  return 1. + (phi / (1. - Phi)) * sigma;
}

// CalculateG2(verbosity = 0) NEST.cpp:2065-2235.  Returns false where the reference throws.
inline bool calculate_g2(const nb200_detector& d, double g2p[5]) {
  double alpha = 0.137, beta = 4.70e-18;
  double epsilonRatio = kEpsLiq / std::fabs(kEpsGas);
  if (d.inGas) epsilonRatio = 1.;
  double E_liq = d.E_gas / epsilonRatio;
  double T = d.T_Kelvin;
  double em1 = 8.807528626640e4 - 2.026247730928e3 * T + 1.747197309338e1 * pow(T, 2.) -
               6.692362929271e-2 * pow(T, 3.) + 9.607626262594e-5 * pow(T, 4.);
  double em2 = 5.074800229635e5 - 1.460168019275e4 * T + 1.680089978382e2 * pow(T, 2.) -
               9.663183204468e-1 * pow(T, 3.) + 2.778229721617e-3 * pow(T, 4.) -
               3.194249083426e-6 * pow(T, 5.);
  double em3 = -4.659269964120e6 + 1.366555237249e5 * T - 1.602830617076e3 * pow(T, 2.) +
               9.397480411915e-0 * pow(T, 3.) - 2.754232523872e-2 * pow(T, 4.) +
               3.228101180588e-5 * pow(T, 5.);
  double ExtEff = 1. - em1 * exp(-em2 * pow(E_liq, em3));
  if (ExtEff > 1. || d.inGas) ExtEff = 1.;
  if (ExtEff < 0. || E_liq <= 0.) ExtEff = 0.;
  double gasGap = d.anode - d.TopDrift;
  if (gasGap <= 0. && E_liq > 0.) return false;
  bool yesGas = true;
  double rho = density(d.T_Kelvin, d.p_bar, yesGas, d.molarMass, nullptr);
  double elYield = (alpha * d.E_gas * 1e3 - beta * (NB_NEST_AVO * rho / d.molarMass)) * gasGap * 0.1;
  double SE = elYield * d.g1_gas;
  if (d.s2_thr < 0) SE *= fit_tba(d.FitTBA, 0., 0., d.TopDrift / 2.);
  g2p[0] = elYield;
  g2p[1] = ExtEff;
  g2p[2] = SE;
  g2p[3] = ExtEff * SE;
  g2p[4] = gasGap;
  return true;
}

// SetDriftVelocity_NonUniform NEST.cpp:2664-2697 (liquid).  The reference recomputes the
// tail sum for every start position (O(n^2) drift-velocity evaluations); here the inverse
// velocities are evaluated once per grid node and suffix-summed in the SAME order the
// reference adds them (from pos_z upward) so that each table entry is the identical sum.
inline std::vector<double> drift_table_nonuniform(const nb200_detector& d, double zStep, double x,
                                                  double y) {
  std::vector<double> zs, inv_above, inv_below;
  for (double z = 0.0; z < d.TopDrift; z += zStep) zs.push_back(z);
  const size_t n = zs.size();
  // the reference's inner loop restarts zz at pos_z and keeps adding zStep, i.e. it walks
  // the same chain of additions from 0 as the outer loop: zz after k steps IS zs[i+k]
  bool ok;
  double v_gate = drift_velocity_liquid(d.T_Kelvin, d.E_gas / (kEpsLiq / std::fabs(kEpsGas)) * 1e3, &ok);
  std::vector<double> inv(n);
  for (size_t i = 0; i < n; ++i)
    inv[i] = zStep / drift_velocity_liquid(d.T_Kelvin, fit_ef(d.FitEF, x, y, zs[i]), &ok);
  std::vector<double> table(n);
  for (size_t i = 0; i < n; ++i) {
    double driftTime = 0.0, zz = zs[i];
    size_t k = i;
    for (; zz < d.TopDrift; zz += zStep, ++k) {
      if (zs[i] > d.gate)
        driftTime += zStep / v_gate;
      else
        driftTime += (k < n) ? inv[k] : zStep / drift_velocity_liquid(d.T_Kelvin, fit_ef(d.FitEF, x, y, zz), &ok);
    }
    table[i] = (zz - zs[i]) / driftTime;
  }
  return table;
}

// TestSpectra::WIMP_dRate TestSpectra.cpp:251-390 with the isotope A made explicit.
// ok=false where the reference throws ("velocity integral broke").
inline double wimp_drate(double ER, double mWimp, double dayNum, double A, bool* ok) {
  const double V_SUN = 250.2, V_WIMP = 238., V_ESCAPE = 544., RHO_NAUGHT = 0.3;
  double M_N = 0.9395654, N_A = NB_NEST_AVO, c = 2.99792458e10, GeVperAMU = 0.9315;
  double SecondsPerDay = 60. * 60. * 24., KiloGramsPerGram = 0.001, keVperGeV = 1.e6, cmPerkm = 1.e5;
  double SqrtPi = pow(M_PI, 0.5), root2 = sqrt(2.);
  double v_0 = V_WIMP * cmPerkm, v_esc = V_ESCAPE * cmPerkm;
  double v_e = (V_SUN + (0.49 * 29.8 * cos((dayNum * 2. * M_PI / 365.24) - (0.415 * 2. * M_PI)))) * cmPerkm;
  double Z = NB_ATOM_NUM;
  double M_T = A * GeVperAMU;
  double N_T = N_A / (A * KiloGramsPerGram);
  ER /= keVperGeV;
  double delta = 0. / keVperGeV;
  double m_d = mWimp, sigma_n = 1.e-36;
  double mu_ND = mWimp * M_N / (mWimp + M_N), mu_TD = mWimp * M_T / (mWimp + M_T);
  double fp = 1., fn = 1.;
  double v_min = 0.;
  if (ER != 0.) v_min = c * (((M_T * ER) / mu_TD) + delta) / (root2 * sqrt(M_T * ER));
  double bet = 0.;
  double x_min = v_min / v_0, x_e = v_e / v_0, x_esc = v_esc / v_0;
  double N = SqrtPi * SqrtPi * SqrtPi * v_0 * v_0 * v_0 *
             (erf(x_esc) - (4. / SqrtPi) * exp(-x_esc * x_esc) * (x_esc / 2. + bet * x_esc * x_esc * x_esc / 3.));
  double zeta = 0.;
  int thisCase = -1;
  if ((x_e + x_min) < x_esc) thisCase = 1;
  if ((x_min > std::fabs(x_esc - x_e)) && ((x_e + x_esc) > x_min)) thisCase = 2;
  if (x_e > (x_min + x_esc)) thisCase = 3;
  if ((x_e + x_esc) < x_min) thisCase = 4;
  if (ok) *ok = true;
  switch (thisCase) {
    case 1:
      zeta = ((SqrtPi * SqrtPi * SqrtPi * v_0 * v_0) / (2. * N * x_e)) *
             (erf(x_min + x_e) - erf(x_min - x_e) -
              ((4. * x_e) / SqrtPi) * exp(-x_esc * x_esc) *
                  (1 + bet * (x_esc * x_esc - x_e * x_e / 3. - x_min * x_min)));
      break;
    case 2:
      zeta = ((SqrtPi * SqrtPi * SqrtPi * v_0 * v_0) / (2. * N * x_e)) *
             (erf(x_esc) + erf(x_e - x_min) -
              (2. / SqrtPi) * exp(-x_esc * x_esc) *
                  (x_esc + x_e - x_min -
                   (bet / 3.) * (x_e - 2. * x_esc - x_min) * (x_esc + x_e - x_min) * (x_esc + x_e - x_min)));
      break;
    case 3: zeta = 1. / (x_e * v_0); break;
    case 4: zeta = 0.; break;
    default:
      if (ok) *ok = false;
      return 0.;
  }
  double a = 0.52, C = 1.23 * pow(A, 1. / 3.) - 0.60, s = 0.9;
  double rn = sqrt(C * C + (7. / 3.) * M_PI * M_PI * a * a - 5. * s * s);
  double q = 6.92 * sqrt(A * ER);
  double FormFactor;
  if (q * rn > 0.)
    FormFactor = 3. * exp(-0.5 * q * q * s * s) * (sin(q * rn) - q * rn * cos(q * rn)) /
                 (q * rn * q * rn * q * rn);
  else
    FormFactor = 1.;
  double dSpec = 0.5 * (c * c) * N_T * (RHO_NAUGHT / m_d) * (M_T * sigma_n / (mu_ND * mu_ND));
  dSpec *= (((Z * fp) + ((A - Z) * fn)) / fn) * (((Z * fp) + ((A - Z) * fn)) / fn) * zeta * FormFactor *
           FormFactor * SecondsPerDay / keVperGeV;
  return dSpec;
}

static const int kXeIsotopes[9] = {124, 126, 128, 129, 130, 131, 132, 134, 136};

// host SelectRanXeAtom over a Philox stream (RandomGen.cpp:159-182)
inline int select_xe_atom(Stream& g) {
  double iso = g.uniform() * 100.;
  static const double edge[8] = {0.090, 0.180, 2.100, 28.54, 32.62, 53.80, 80.69, 91.13};
  for (int i = 0; i < 8; ++i)
    if (iso <= edge[i]) return kXeIsotopes[i];
  return 136;
}

}  // namespace host
}  // namespace nb
