// TEST INFRASTRUCTURE ONLY -- never linked into or called by the product path.
//
// Thin extern "C" probes over the UNMODIFIED reference objects (built by
// oracle/Makefile from /root/reference/src/*.cpp into oracle/_ref/).  Every probe
// calls the reference's own classes; the only logic restated here is the event
// loop body of execNEST() (src/execNEST.cpp:676-1250), which cannot be called
// per event because it printf()s; tests/test_oracle_pin.py checks that this loop
// reproduces the reference binary's stdout bit for bit at equal seed.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "NEST.hh"
#include "LArNEST.hh"
#include "LArDetector.hh"
#include "LUX_Run03.hh"
#include "TestSpectra.hh"
#include "execNEST.hh"
#include "nest_b200.h"
#include "nest_b200_flatten.hh"

using namespace NEST;

// compile-time "flags" of analysis.hh, defined in the execNEST.cpp object
extern int verbosity, usePD, useS2, numBins;
extern bool MCtruthE, MCtruthPos;
extern double minS1, maxS1, minS2, maxS2, logMin, logMax, z_step, E_step;
extern double minTimeSeparation;  // execNEST.cpp:38
extern double band[NUMBINS_MAX][7];
extern NEST::S1CalculationMode s1CalculationMode;
extern NEST::S2CalculationMode s2CalculationMode;

namespace {
DetectorExample_LUX_RUN03* g_lux = nullptr;
nb200_flatten_storage g_store;
DetectorExample_LUX_RUN03* lux() {
  if (!g_lux) g_lux = new DetectorExample_LUX_RUN03();
  return g_lux;
}
std::vector<double> vec(const double* p, int n) { return std::vector<double>(p, p + n); }
}  // namespace

extern "C" {

int ref_set_verbosity(int v) { int o = verbosity; verbosity = v; return o; }

// detector parameters of a sweep grid point on the probe's LUX_Run03 instance (VDetector.hh:73-103 setters, as the
// reference's loopNEST mode uses them, execNEST.cpp:134-144); a negative value restores the header's default
void ref_lux_set(double g1, double g1_gas, double E_gas) {
  DetectorExample_LUX_RUN03 def;
  lux()->set_g1(g1 >= 0. ? g1 : def.get_g1());
  lux()->set_g1_gas(g1_gas >= 0. ? g1_gas : def.get_g1_gas());
  lux()->set_E_gas(E_gas >= 0. ? E_gas : def.get_E_gas());
}

void ref_analysis(double* out) {
  out[0] = usePD; out[1] = useS2; out[2] = numBins; out[3] = minS1; out[4] = maxS1;
  out[5] = minS2; out[6] = maxS2; out[7] = logMin; out[8] = logMax;
  out[9] = MCtruthE; out[10] = MCtruthPos; out[11] = z_step;
}

// the header-only flattener compiled against the real VDetector
int ref_flatten_lux(nb200_detector* out, int force_lut) {
  nb200_flatten(*lux(), out, &g_store, 257, 513, force_lut != 0);
  return 0;
}

void ref_maps(int n, const double* x, const double* y, const double* z, double* s1,
              double* s2, double* ef, double* tba) {
  for (int i = 0; i < n; ++i) {
    s1[i] = lux()->FitS1(x[i], y[i], z[i], VDetector::fold);
    s2[i] = lux()->FitS2(x[i], y[i], VDetector::fold);
    ef[i] = lux()->FitEF(x[i], y[i], z[i]);
    tba[i] = lux()->FitTBA(x[i], y[i], z[i])[1];
  }
}

// out: rho, Wq_eV, alpha, vD(field), g2_params[5], NewMean(sPEres)
void ref_consts(double field, double* out) {
  NESTcalc n(lux());
  double rho = n.SetDensity(lux()->get_T_Kelvin(), lux()->get_p_bar());
  auto w = NESTcalc::WorkFunction(rho, lux()->get_molarMass(), lux()->get_OldW13eV());
  out[0] = rho; out[1] = w.Wq_eV; out[2] = w.alpha;
  out[3] = n.SetDriftVelocity(lux()->get_T_Kelvin(), rho, field, lux()->get_p_bar());
  auto g2 = n.CalculateG2(0);
  for (int i = 0; i < 5; ++i) out[4 + i] = g2[i];
  out[9] = RandomGen::rndm()->FindNewMean(lux()->get_sPEres());
}

double ref_density(double T, double P) {
  bool inGas = false;
  return NESTcalc::GetDensity(T, P, inGas, 1);
}
double ref_drift_velocity_liquid(double T, double F, double rho, double P) {
  return NESTcalc::GetDriftVelocity_Liquid(T, F, rho, P);
}

// which: 0 GetYields(species) dispatch, 1 GetYieldBetaGR(multFact), 2 GetYieldGamma(multFact),
// 3 GetYieldERWeighted.  out AoS [n][6].
void ref_yields(int which, int species, int n, const double* E, const double* field,
                double rho, double massNum, double atomNum, const double* nrpar,
                const double* erpar, double multFact, double* out) {
  NESTcalc calc(lux());
  calc.SetDensity(lux()->get_T_Kelvin(), lux()->get_p_bar());
  auto nr = vec(nrpar, 12), er = vec(erpar, 10);
  for (int i = 0; i < n; ++i) {
    YieldResult y;
    if (which == 1) y = calc.GetYieldBetaGR(E[i], rho, field[i], er, multFact);
    else if (which == 2) y = calc.GetYieldGamma(E[i], rho, field[i], multFact);
    else if (which == 3) y = calc.GetYieldERWeighted(E[i], rho, field[i], er);
    else y = calc.GetYields(INTERACTION_TYPE(species), E[i], rho, field[i], massNum, atomNum, nr, er);
    double* o = out + 6 * size_t(i);
    o[0] = y.PhotonYield; o[1] = y.ElectronYield; o[2] = y.ExcitonRatio;
    o[3] = y.Lindhard; o[4] = y.ElectricField; o[5] = y.DeltaT_Scint;
  }
}

// width-model helpers: RecombOmegaNR / RecombOmegaER / FanoER
void ref_widths(int n, const double* efield, const double* elecFrac, const double* nq,
                double rho, const double* widths, int nw, double* out /*[n][3]*/) {
  NESTcalc calc(lux());
  auto w = vec(widths, nw);
  for (int i = 0; i < n; ++i) {
    out[3 * i + 0] = calc.RecombOmegaNR(elecFrac[i], w);
    out[3 * i + 1] = calc.RecombOmegaER(efield[i], elecFrac[i], nq[i], w);
    out[3 * i + 2] = calc.FanoER(rho, nq[i], efield[i], w);
  }
}

// GetQuanta on fixed yields (AoS [6]); n draws.  iout [n][4] = Nph,Ne,Ni,Nex; dout [n][2]
void ref_quanta(uint64_t seed, int n, const double* y6, double rho, const double* widths,
                int nw, double skewER, int* iout, double* dout) {
  NESTcalc calc(lux());
  RandomGen::rndm()->SetSeed(seed);
  YieldResult y{y6[0], y6[1], y6[2], y6[3], y6[4], y6[5]};
  auto w = vec(widths, nw);
  for (int i = 0; i < n; ++i) {
    QuantaResult q = calc.GetQuanta(y, rho, w, skewER);
    iout[4 * i] = q.photons; iout[4 * i + 1] = q.electrons;
    iout[4 * i + 2] = q.ions; iout[4 * i + 3] = q.excitons;
    dout[2 * i] = q.recombProb; dout[2 * i + 1] = q.Variance;
  }
}

// RandomGen samplers. kind: 0 uniform, 1 gauss(p0,p1,zero_min=p2), 2 zero_trunc_gauss(p0,p1),
// 3 skewGauss(p0,p1,p2), 4 binom(N=p0,p=p1), 5 exponential(p0,p1,p2), 6 SelectRanXeAtom,
// 7 poisson(p0)
void ref_sampler(int kind, uint64_t seed, int n, double p0, double p1, double p2, double* out) {
  RandomGen* r = RandomGen::rndm();
  r->SetSeed(seed);
  for (int i = 0; i < n; ++i) {
    switch (kind) {
      case 0: out[i] = r->rand_uniform(); break;
      case 1: out[i] = r->rand_gauss(p0, p1, p2 != 0.); break;
      case 2: out[i] = r->rand_zero_trunc_gauss(p0, p1); break;
      case 3: out[i] = r->rand_skewGauss(p0, p1, p2); break;
      case 4: out[i] = double(r->binom_draw(int64_t(p0), p1)); break;
      case 5: out[i] = r->rand_exponential(p0, p1, p2); break;
      case 6: out[i] = r->SelectRanXeAtom(); break;
      default: out[i] = double(r->poisson_draw(p0)); break;
    }
  }
}

// TestSpectra samplers. kind: NB200_SRC_CH3T / DD / B8
void ref_spectrum(int kind, uint64_t seed, int n, double eMin, double eMax, double* out) {
  RandomGen::rndm()->SetSeed(seed);
  for (int i = 0; i < n; ++i) {
    if (kind == NB200_SRC_CH3T) out[i] = TestSpectra::CH3T_spectrum(eMin, eMax);
    else if (kind == NB200_SRC_DD) out[i] = TestSpectra::DD_spectrum(eMin, eMax, 13., 0.12, 71.2, 20., -20.5);
    else if (kind == NB200_SRC_AMBE) out[i] = TestSpectra::PowLawFit_spectrum(eMin, eMax, -0.95351);
    else if (kind == NB200_SRC_C14) out[i] = TestSpectra::C14_spectrum(eMin, eMax);
    else if (kind == NB200_SRC_CF) out[i] = TestSpectra::Cf_spectrum(eMin, eMax);
    else if (kind == NB200_SRC_PPSOLAR) out[i] = TestSpectra::ppSolar_spectrum(eMin, eMax);
    else if (kind == NB200_SRC_ATMNU) out[i] = TestSpectra::atmNu_spectrum(eMin, eMax);
    else out[i] = TestSpectra::B8_spectrum(eMin, eMax);
  }
}

// TestSpectra::WIMP_dRate (TestSpectra.cpp:251-390) draws its own Xe isotope (:278).  Call i runs on the generator
// seeded with seed + i, after one SelectRanXeAtom under the same seed has told us which isotope that will be:
// out[i] = the rate, A_out[i] = the isotope it was computed for.
void ref_wimp_drate(uint64_t seed, int n, const double* ER, double mass, double day, double* out, double* A_out) {
  RandomGen* r = RandomGen::rndm();
  for (int i = 0; i < n; ++i) {
    r->SetSeed(seed + uint64_t(i));
    A_out[i] = double(r->SelectRanXeAtom());
    r->SetSeed(seed + uint64_t(i));
    out[i] = TestSpectra::WIMP_dRate(ER[i], mass, day);
  }
}

// TestSpectra::WIMP_prep_spectrum (TestSpectra.cpp:392-454) on the generator seeded with `seed`: the isotopes of its
// WIMP_dRate calls are then the first draws of that stream (ref_sampler kind 6 lists them).
// base, exponent: 100 doubles each; out3 = integral, xMax, divisor.  1 where the reference throws.
int ref_wimp_prep(uint64_t seed, double mass, double eStep, double day, double* base, double* exponent, double* out3) {
  RandomGen::rndm()->SetSeed(seed);
  try {
    TestSpectra::WIMP_spectrum_prep p = TestSpectra::WIMP_prep_spectrum(mass, eStep, day);
    for (int i = 0; i < 100; ++i) { base[i] = p.base[i]; exponent[i] = p.exponent[i]; }
    out3[0] = p.integral; out3[1] = p.xMax; out3[2] = p.divisor;
  } catch (std::exception& e) {
    fprintf(stderr, "ref_wimp_prep: %s\n", e.what());
    return 1;
  }
  return 0;
}

// n draws of TestSpectra::WIMP_spectrum (TestSpectra.cpp:456-499) from the table of ref_wimp_prep(seed_prep), the
// generator re-seeded with seed_draw before the first draw
int ref_wimp_spectrum(uint64_t seed_prep, uint64_t seed_draw, int n, double mass, double eStep, double day, double* out) {
  RandomGen::rndm()->SetSeed(seed_prep);
  try {
    TestSpectra::WIMP_spectrum_prep p = TestSpectra::WIMP_prep_spectrum(mass, eStep, day);
    RandomGen::rndm()->SetSeed(seed_draw);
    for (int i = 0; i < n; ++i) out[i] = TestSpectra::WIMP_spectrum(p, mass, day);
  } catch (std::exception& e) {
    fprintf(stderr, "ref_wimp_spectrum: %s\n", e.what());
    return 1;
  }
  return 0;
}

// the Poisson-fluctuated event count of the WIMP / 8B / pp / atmNu branches (execNEST.cpp:482-488, 490, 529, 536)
void ref_poisson(uint64_t seed, int n, double mean, double* out) {
  RandomGen::rndm()->SetSeed(seed);
  for (int i = 0; i < n; ++i) out[i] = double(RandomGen::rndm()->poisson_draw(mean));
}

// GetS1 on fixed quanta/position; out [n][9]
void ref_s1(uint64_t seed, int n, int Nph, const double* pos /*xyz*/, const double* spos,
            double vD, double vDmid, int species, double field, double energy, int mode,
            double* out) {
  NESTcalc calc(lux());
  RandomGen::rndm()->SetSeed(seed);
  QuantaResult q{};
  q.photons = Nph;
  std::vector<int64_t> wt; std::vector<double> wa;
  for (int i = 0; i < n; ++i) {
    const std::vector<double>& s = calc.GetS1(q, pos[0], pos[1], pos[2], spos[0], spos[1], spos[2], vD,
                                   vDmid, INTERACTION_TYPE(species), i, field, energy,
                                   S1CalculationMode(mode), 0, wt, wa);
    for (int k = 0; k < 9; ++k) out[9 * size_t(i) + k] = s[k];
  }
}

// GetS2 on fixed Ne/position; out [n][9]
void ref_s2(uint64_t seed, int n, int Ne, const double* pos, const double* spos, double dt,
            double vD, double field, double* out) {
  NESTcalc calc(lux());
  RandomGen::rndm()->SetSeed(seed);
  auto g2 = calc.CalculateG2(0);
  std::vector<int64_t> wt; std::vector<double> wa;
  for (int i = 0; i < n; ++i) {
    const std::vector<double>& s = calc.GetS2(Ne, pos[0], pos[1], pos[2], spos[0], spos[1], spos[2], dt,
                                   vD, i, field, S2CalculationMode::Full, 0, wt, wa, g2);
    for (int k = 0; k < 9; ++k) out[9 * size_t(i) + k] = s[k];
  }
}

// xyResolution; out [n][2]
void ref_xyres(uint64_t seed, int n, double x, double y, double A_top, double* out) {
  NESTcalc calc(lux());
  RandomGen::rndm()->SetSeed(seed);
  for (int i = 0; i < n; ++i) {
    auto v = calc.xyResolution(x, y, A_top);
    out[2 * i] = v[0]; out[2 * i + 1] = v[1];
  }
}

// Columns written by ref_chain / the oracle chain, [n][NCOL] AoS
enum { C_KEV_TRUE = 0, C_KEV_OUT, C_FIELD, C_DT, C_X, C_Y, C_Z, C_SX, C_SY, C_NPH, C_NE, C_NI,
       C_NEX, C_YPH, C_YNE, C_S1_0, C_S2_0 = C_S1_0 + 9, C_SIG1 = C_S2_0 + 9, C_SIG2, C_SIGE,
       NCOL };

int ref_chain_ncol() { return NCOL; }

// The loop body of execNEST() (src/execNEST.cpp:676-1250) for the energy-based species,
// random or fixed position, uniform or FitEF field; same RNG call order as the
// reference, so equal seed => equal numbers as the reference binary prints.
//   er_model: -1 plain GetYields(species); 0 BetaGR(multFact); 1 Gamma(-multFact); 2 ERWeighted
//   src_kind: NB200_SRC_* ; g2_verbosity: what execNEST passes to CalculateG2 (=verbosity)
int ref_chain(uint64_t seed, int n, int species, int src_kind, double eMin, double eMax,
              double inField, int fixed_pos, const double* xyz, int er_model, double multFact,
              const double* nrpar, const double* erpar, const double* widths, int nw,
              double skewER, int g2_verbosity, double* out) {
  VDetector* detector = lux();
  NESTcalc n_(detector);
  RandomGen::rndm()->SetSeed(seed);
  auto NRYieldsParam = vec(nrpar, 12), ERYieldsParam = vec(erpar, 10), NRERWidthsParam = vec(widths, nw);
  INTERACTION_TYPE type_num = INTERACTION_TYPE(species);
  TestSpectra::WIMP_spectrum_prep wimp_prep;
  if (src_kind == NB200_SRC_WIMP) {  // execNEST.cpp:474-481: eMin is the WIMP mass; day 0; E_step of analysis.hh
    try {
      wimp_prep = TestSpectra::WIMP_prep_spectrum(eMin, E_step, 0.);
    } catch (std::exception& e) {
      fprintf(stderr, "ref_chain: %s\n", e.what());
      return 1;
    }
  }
  double rho = n_.SetDensity(detector->get_T_Kelvin(), detector->get_p_bar());
  double Wq_eV = NESTcalc::WorkFunction(rho, detector->get_molarMass(), detector->get_OldW13eV()).Wq_eV;
  std::vector<double> g2_params = n_.CalculateG2(g2_verbosity);
  double g2 = std::abs(g2_params[3]);
  double g1 = detector->get_g1();
  double centralZ = (detector->get_gate() * 0.8 + detector->get_cathode() * 1.03) / 2.;
  double massNum = detector->get_molarMass(), atomNum = 0;
  if (type_num == ion) { atomNum = 2; massNum = 4; }  // the CLI's "alpha" (execNEST.cpp:496-500)
  if (type_num < 6) massNum = 0;
  if (type_num == Kr83m) { massNum = eMax; atomNum = minTimeSeparation; }  // execNEST.cpp:671, 1052
  double vD = 0, vD_middle = 0, pos_x = 0, pos_y = 0, pos_z = 0, field, driftTime;
  std::vector<double> vTable;
  try {
    for (int j = 0; j < n; ++j) {
      double keV;
      if (src_kind == NB200_SRC_MONO || type_num == Kr83m) keV = eMin;
      else if (src_kind == NB200_SRC_CH3T) keV = TestSpectra::CH3T_spectrum(eMin, eMax);
      else if (src_kind == NB200_SRC_WIMP) keV = TestSpectra::WIMP_spectrum(wimp_prep, eMin, 0.);
      else if (src_kind == NB200_SRC_B8) keV = TestSpectra::B8_spectrum(eMin, eMax);
      else if (src_kind == NB200_SRC_DD) keV = TestSpectra::DD_spectrum(eMin, eMax, 13., 0.12, 71.2, 20., -20.5);
      else if (src_kind == NB200_SRC_AMBE) keV = TestSpectra::PowLawFit_spectrum(eMin, eMax, -0.95351);
      else if (src_kind == NB200_SRC_C14) keV = TestSpectra::C14_spectrum(eMin, eMax);
      else if (src_kind == NB200_SRC_CF) keV = TestSpectra::Cf_spectrum(eMin, eMax);
      else if (src_kind == NB200_SRC_PPSOLAR) keV = TestSpectra::ppSolar_spectrum(eMin, eMax);
      else if (src_kind == NB200_SRC_ATMNU) keV = TestSpectra::atmNu_spectrum(eMin, eMax);
      else if (src_kind == NB200_SRC_GAUSS) keV = RandomGen::rndm()->rand_zero_trunc_gauss(eMin, eMax);
      else if (src_kind == NB200_SRC_EXP) keV = RandomGen::rndm()->rand_exponential(-eMax * log(2.), 0., eMin);
      else keV = eMin + (eMax - eMin) * RandomGen::rndm()->rand_uniform();
      if (type_num != WIMP && type_num != B8 && type_num != Kr83m && type_num != ppSolar && type_num != atmNu && eMax > 0. &&
          eMin < eMax) {
        if (keV > eMax) keV = eMax;
        if (keV < eMin) keV = eMin;
      }
      double keV_true = keV;
    Z_NEW:
      if (!fixed_pos) {
        pos_z = 0. + (detector->get_TopDrift() - 0.) * RandomGen::rndm()->rand_uniform();
        double r = detector->get_radius() * sqrt(RandomGen::rndm()->rand_uniform());
        double phi = 2. * M_PI * RandomGen::rndm()->rand_uniform();
        pos_x = r * cos(phi);
        pos_y = r * sin(phi);
      } else {
        pos_x = xyz[0]; pos_y = xyz[1]; pos_z = xyz[2];
      }
      if (inField == -1.) field = detector->FitEF(pos_x, pos_y, pos_z);
      else field = inField;
      if (j == 0 && vD_middle == 0.) {
        if (inField == -1.) {
          vTable = n_.SetDriftVelocity_NonUniform(rho, z_step, detector->get_T_Kelvin(), detector->get_p_bar(), pos_x, pos_y);
          vD_middle = vTable[int(floor(centralZ / z_step + 0.5))];
        } else {
          vD_middle = n_.SetDriftVelocity(detector->get_T_Kelvin(), rho, inField, detector->get_p_bar());
          vD = n_.SetDriftVelocity(detector->get_T_Kelvin(), rho, field, detector->get_p_bar());
        }
      }
      if (inField == -1.) vD = vTable[int(floor(pos_z / z_step + 0.5))];
      driftTime = (detector->get_TopDrift() - pos_z) / vD;
      if ((driftTime > detector->get_dt_max() || driftTime < detector->get_dt_min()) && !fixed_pos &&
          field >= FIELD_MIN)
        goto Z_NEW;
      YieldResult yields;
      QuantaResult quanta{};
      if (keV > 0.001 * Wq_eV) {
        if (er_model == 0) yields = n_.GetYieldBetaGR(keV, rho, field, ERYieldsParam, multFact);
        else if (er_model == 1) yields = n_.GetYieldGamma(keV, rho, field, -multFact);
        else if (er_model == 2) yields = n_.GetYieldERWeighted(keV, rho, field, ERYieldsParam, default_EnergyParams, default_FieldParams);
        else yields = n_.GetYields(type_num, keV, rho, field, double(massNum), double(atomNum), NRYieldsParam, ERYieldsParam);
        if (type_num == ion)  // execNEST.cpp:1056-1063
          NRERWidthsParam = {1.00, 1.00, 0., 0.50, 0.19, 0., 1., 0.046452, 0.45, 0.205, -0.2, 0., 0.};
        quanta = n_.GetQuanta(yields, rho, NRERWidthsParam, skewER);
      } else {
        yields = YieldResult{0., 0., 0., 0., 0., 0.};
      }
      double truthPos[3] = {pos_x, pos_y, pos_z};
      double smearPos[3] = {pos_x, pos_y, pos_z};
      double Nphd_S2 = g2 * quanta.electrons * exp(-driftTime / detector->get_eLife_us());
      if (!MCtruthPos && Nphd_S2 > PHE_MIN) {
        std::vector<double> xySmeared = n_.xyResolution(pos_x, pos_y, Nphd_S2);
        smearPos[0] = xySmeared[0];
        smearPos[1] = xySmeared[1];
      }
      std::vector<int64_t> wf_time; std::vector<double> wf_amp;
      std::vector<double> scint = n_.GetS1(quanta, truthPos[0], truthPos[1], truthPos[2], smearPos[0], smearPos[1],
                                  smearPos[2], vD, vD_middle, type_num, j, field, keV, s1CalculationMode,
                                  verbosity, wf_time, wf_amp);
      if (truthPos[2] < detector->get_cathode()) quanta.electrons = 0;
      std::vector<double> scint2 = n_.GetS2(quanta.electrons, truthPos[0], truthPos[1], truthPos[2], smearPos[0],
                                   smearPos[1], smearPos[2], driftTime, vD, j, field, s2CalculationMode,
                                   verbosity, wf_time, wf_amp, g2_params);
      double signal1, signal2;
      if (usePD == 0 && std::abs(scint[3]) > minS1 && scint[3] < maxS1) signal1 = scint[3];
      else if (usePD == 1 && std::abs(scint[5]) > minS1 && scint[5] < maxS1) signal1 = scint[5];
      else if ((usePD >= 2 && std::abs(scint[7]) > minS1 && scint[7] < maxS1) || maxS1 >= 998.5) signal1 = scint[7];
      else signal1 = -999.;
      if (usePD == 0 && std::abs(scint2[5]) > minS2 && scint2[5] < maxS2) signal2 = scint2[5];
      else if (usePD >= 1 && std::abs(scint2[7]) > minS2 && scint2[7] < maxS2) signal2 = scint2[7];
      else signal2 = -999.;
      double Nph = 0.0, Ne = 0.0;
      if (!MCtruthE) {
        double MultFact = 1., eff = detector->get_sPEeff();
        if (detector->get_sPEthr() >= 0. && detector->get_sPEres() > 0. && eff > 0.) {
          MultFact = 0.5 * (1. + erf((detector->get_sPEthr() - 1.) / (detector->get_sPEres() * sqrt(2.)))) /
                     (detector->get_sPEres() * sqrt(2. * M_PI));
          if (usePD == 2 && scint[7] < SPIKES_MAXM) MultFact = 1.04;
          else MultFact = 1. + MultFact * detector->get_P_dphe();
          if (eff < 1.) eff += ((1. - eff) / (2. * double(detector->get_numPMTs()))) * scint[0];
          eff = max(0., min(eff, 1.));
          MultFact /= eff;
        } else MultFact = 1.;
        if (usePD == 0) Nph = std::abs(scint[3]) / (g1 * (1. + detector->get_P_dphe()));
        else if (usePD == 1) Nph = std::abs(scint[5]) / g1;
        else Nph = std::abs(scint[7]) / g1;
        if (usePD == 0) Ne = std::abs(scint2[5]) / (g2 * (1. + detector->get_P_dphe()));
        else Ne = std::abs(scint2[7]) / g2;
        Nph *= MultFact;
        if (signal1 <= 0.) Nph = 0.;
        if (signal2 <= 0.) Ne = 0.;
        if (detector->get_coinLevel() <= 0 && Nph <= PHE_MIN) Nph = DBL_MIN;
        if (yields.Lindhard > DBL_MIN && Nph > 0. && Ne > 0.) {
          if (ValidityTests::nearlyEqual(yields.Lindhard, 1.)) keV = (Nph + Ne) * Wq_eV * 1e-3;
          else if (type_num <= NEST::INTERACTION_TYPE::Cf) {
            keV = pow((Ne + Nph) / NRYieldsParam[0], 1. / NRYieldsParam[1]);
            Ne *= 1. - 1. / pow(1. + pow((keV / NRYieldsParam[5]), NRYieldsParam[6]), NRYieldsParam[10]);
            Nph *= 1. - 1. / pow(1. + pow((keV / NRYieldsParam[7]), NRYieldsParam[8]), NRYieldsParam[11]);
            keV = pow((Ne + Nph) / NRYieldsParam[0], 1. / NRYieldsParam[1]);
          } else {
            const double logden = log10(rho);
            const double Wq_eV_alpha = 28.259 + 25.667 * logden - 33.611 * pow(logden, 2.) -
                                       123.73 * pow(logden, 3.) - 136.47 * pow(logden, 4.) -
                                       74.194 * pow(logden, 5.) - 20.276 * pow(logden, 6.) -
                                       2.2352 * pow(logden, 7.);
            keV = (Nph + Ne) * Wq_eV_alpha * 1e-3 / yields.Lindhard;
          }
        } else keV = 0.;
      }
      double signalE;
      if ((signal1 <= 0. || signal2 <= 0.) && field >= FIELD_MIN && detector->get_coinLevel() != 0) signalE = 0.;
      else signalE = keV;
      double* o = out + size_t(j) * NCOL;
      o[C_KEV_TRUE] = keV_true; o[C_KEV_OUT] = keV; o[C_FIELD] = field; o[C_DT] = driftTime;
      o[C_X] = pos_x; o[C_Y] = pos_y; o[C_Z] = pos_z; o[C_SX] = smearPos[0]; o[C_SY] = smearPos[1];
      o[C_NPH] = quanta.photons; o[C_NE] = quanta.electrons; o[C_NI] = quanta.ions; o[C_NEX] = quanta.excitons;
      o[C_YPH] = yields.PhotonYield; o[C_YNE] = yields.ElectronYield;
      for (int k = 0; k < 9; ++k) { o[C_S1_0 + k] = scint[k]; o[C_S2_0 + k] = scint2[k]; }
      o[C_SIG1] = signal1; o[C_SIG2] = signal2; o[C_SIGE] = signalE;
    }
  } catch (std::exception& e) {
    fprintf(stderr, "ref_chain: %s\n", e.what());
    return 1;
  }
  return 0;
}

// the reference's own GetBand (execNEST.cpp:1508-1621) on caller-supplied signals.
// out [numBins][7]
int ref_getband(int n, const double* s1, const double* s2, double* out) {
  for (int j = 0; j < NUMBINS_MAX; ++j) for (int k = 0; k < 7; ++k) band[j][k] = 0.;
  GetBand(vec(s1, n), vec(s2, n), false, lux()->get_coinLevel());
  for (int j = 0; j < numBins; ++j) for (int k = 0; k < 7; ++k) out[7 * j + k] = band[j][k];
  return numBins;
}

// runNESTvec (execNEST.cpp:329-409) called directly. cols [n][26] in declaration order of
// NESTObservableArray (execNEST.hh:30-57) minus photon times / waveforms.
int ref_runvec(int species, int n, const double* E, const double* xyz /*[n][3]*/, double inField,
               int seed, const double* erpar, const double* nrpar, const double* widths, int nw,
               int s1mode, double* out) {
  std::vector<double> eList(E, E + n);
  std::vector<std::vector<double>> pos(n);
  for (int i = 0; i < n; ++i) pos[i] = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
  NESTObservableArray r = runNESTvec(lux(), INTERACTION_TYPE(species), eList, pos, inField, seed,
                                     vec(erpar, 10), vec(nrpar, 12), vec(widths, nw),
                                     S1CalculationMode(s1mode), S2CalculationMode::Full, false);
  for (int i = 0; i < n; ++i) {
    double* o = out + 26 * size_t(i);
    o[0] = r.energy_kev[i]; o[1] = r.x_mm[i]; o[2] = r.y_mm[i]; o[3] = r.z_mm[i];
    o[4] = r.n_electrons[i]; o[5] = r.n_photons[i];
    o[6] = r.s1_nhits[i]; o[7] = r.s1_nhits_dpe[i]; o[8] = r.s1_nhits_thr[i];
    o[9] = r.s1r_phe[i]; o[10] = r.s1c_phe[i]; o[11] = r.s1r_phd[i]; o[12] = r.s1c_phd[i];
    o[13] = r.s1r_spike[i]; o[14] = r.s1c_spike[i];
    o[15] = r.s2_Nee[i]; o[16] = r.s2_Nph[i]; o[17] = r.s2_nhits[i]; o[18] = r.s2_nhits_dpe[i];
    o[19] = r.s2r_phe[i]; o[20] = r.s2c_phe[i]; o[21] = r.s2r_phd[i]; o[22] = r.s2c_phd[i];
    o[23] = 0; o[24] = 0; o[25] = 0;
  }
  return 0;
}

// LArNEST (src/LArNEST.cpp:17-33): yields [n][9], fluct [n][4]
void ref_lar(uint64_t seed, int species, int n, const double* E, const double* field, double density,
             double* yields, double* fluct) {
  static LArDetector* det = new LArDetector();
  LArNEST lar(det);
  RandomGen::rndm()->SetSeed(seed);
  for (int i = 0; i < n; ++i) {
    LArNESTResult r = lar.FullCalculation(LArInteraction(species), E[i], 0., field[i], density, false);
    double* y = yields + 9 * size_t(i);
    y[0] = r.yields.TotalYield; y[1] = r.yields.QuantaYield; y[2] = r.yields.LightYield;
    y[3] = r.yields.Nph; y[4] = r.yields.Ne; y[5] = r.yields.Nex; y[6] = r.yields.Nion;
    y[7] = r.yields.Lindhard; y[8] = r.yields.ElectricField;
    double* f = fluct + 4 * size_t(i);
    f[0] = r.fluctuations.NphFluctuation; f[1] = r.fluctuations.NeFluctuation;
    f[2] = r.fluctuations.NexFluctuation; f[3] = r.fluctuations.NionFluctuation;
  }
}

}  // extern "C"
