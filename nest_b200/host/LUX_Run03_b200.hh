// LUX_Run03_b200.hh -- the LUX Run03 template detector written against the mirrored VDetector API
// (parameter values and position maps of the reference's include/Detectors/LUX_Run03.hh:30-167).
// A user's own detector header compiles against NEST_b200.hh the same way.
#pragma once
#include <cmath>

#include "NEST_b200.hh"

class DetectorExample_LUX_RUN03 : public VDetector {
 public:
  DetectorExample_LUX_RUN03() { Initialization(); }
  void Initialization() override {
    name = "LUX Run03";
    g1 = 0.1170; sPEres = 0.37; sPEthr = (0.3 * 1.173) / 0.915; sPEeff = 1.00;
    noiseBaseline[0] = 0.00; noiseBaseline[1] = 0.08; noiseBaseline[2] = 0.; noiseBaseline[3] = 0.;
    P_dphe = 0.173; coinWind = 100; coinLevel = 2; numPMTs = 119; OldW13eV = true;
    noiseLinear[0] = 0.0e-2; noiseLinear[1] = 9.0e-2;
    g1_gas = 0.1; s2Fano = 3.6; s2_thr = (150. * 1.173) / 0.915; E_gas = 6.25; eLife_us = 800.;
    T_Kelvin = 173.; p_bar = 1.57; dtCntr = 160.; dt_min = 38.; dt_max = 305.;
    radius = 200.; radmax = 235.; TopDrift = 544.95; anode = 549.2; gate = 539.2; cathode = 55.90;
    PosResFlat = 3.00; PosResBase = 70.8364;
  }
  double FitS1(double x, double y, double z, LCE) override {
    double r = std::sqrt(x * x + y * y);
    double amplitude = 307.9 - 0.3071 * z + 0.0002257 * z * z;
    double shape = 1.1525e-7 * std::sqrt(std::fabs(z - 318.84));
    double corr = (-shape * r * r * r + amplitude) / 307.9;
    if ((corr < 0.5 || corr > 1.5 || std::isnan(corr)) && r < radmax) return 1.;
    return corr;
  }
  double FitEF(double, double, double z) override {
    double ef = 158.92 - 0.2209000 * z + 0.0024485 * z * z - 8.7098e-6 * z * z * z + 1.5049e-8 * std::pow(z, 4.) -
                1.0110e-11 * std::pow(z, 5.);
    if (ef <= 1. || ef > 1e5 || std::isnan(ef)) return 1.;
    return ef;
  }
  double FitS2(double x, double y, LCE) override {
    double r = std::sqrt(x * x + y * y);
    double corr = (9156.3 + 6.22750 * r + 0.38126 * r * r - 0.017144 * std::pow(r, 3.) + 0.0002474 * std::pow(r, 4.) -
                   1.6953e-6 * std::pow(r, 5.) + 5.6513e-9 * std::pow(r, 6.) - 7.3989e-12 * std::pow(r, 7.)) / 9156.3;
    if ((corr < 0.5 || corr > 1.5 || std::isnan(corr)) && r < radmax) return 1.;
    return corr;
  }
  std::vector<double> FitTBA(double, double, double z) override {
    double t = -0.853 + 0.00925 * (z / 10.);
    t = t < -1. ? -1. : (t > 1. ? 1. : t);
    return {(1. - t) / 2., 0.449};
  }
};
