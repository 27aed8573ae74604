// NEST_b200.cpp -- host side of the drop-in: the reference's driver logic (argument checks,
// messages, output formats) around the GPU chain.  No physics is computed here: every number comes
// from libnest_b200.so; reference lines are cited where behaviour is mirrored.
#include "NEST_b200.hh"

#include <mutex>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <stdexcept>

#include "nest_b200_flatten.hh"

using namespace std;
using namespace NEST;

namespace {
const double kHiEregime = 1E+2;  // execNEST.hh:25
const double kPHE_MIN = 1e-6;

// NEST.cpp:1154-1159: the first evaluation of the beta model with a free m5 (ERYieldsParam[4] < 0, the default) says so
// on stderr, once per process and whatever the verbosity.  The model is evaluated on the device here: the notice is
// given when a call is about to route events to it (GetYieldBetaGR, GetYieldERWeighted, the default branch of
// GetYields :1245-1249: beta, CH3T, 14C, pp solar).
void beta_model_m5_notice(const nb200_model& m, int species) {
  const bool other_model = species <= Cf || species == atmNu || species == ion || species == gammaRay ||
                           species == Kr83m || species == fullGamma_PE;
  const bool beta = m.er_model == 0 || m.er_model == 2 || (m.er_model < 0 && !other_model);
  if (!beta || !(m.ERYieldsParam[4] < 0.)) return;
  static std::once_flag said;
  std::call_once(said, []() { cerr << "Warning: m5 is slightly energy dependent in the beta model" << endl; });
}

bool nearlyEqual(double a, double b, double eps = 1e-9) {  // ValidityTests.cpp:6-21
  if (a == b) return true;
  double A = fabs(a), B = fabs(b), d = fabs(a - b);
  if (a == 0 || b == 0 || (A + B < DBL_MIN)) return d < eps * DBL_MIN;
  return d / fmin(A + B, DBL_MAX) < eps;
}
void fill_model(nb200_model* m, const vector<double>& nr, const vector<double>& er, const vector<double>& w) {
  nb200_model_default(0, m);
  if (nr.size() < 12) throw runtime_error("ERROR: You need a minimum of 12 nuisance parameters for the mean yields.");
  if (er.size() < 10) throw runtime_error("ERROR: You need a minimum of 10 nuisance parameters for the mean yields.");
  if (w.size() < 13) throw runtime_error("ERROR: You need a minimum of 13 free parameters for the resolution model.");
  for (int i = 0; i < 12; ++i) m->NRYieldsParam[i] = nr[i];
  for (int i = 0; i < 10; ++i) m->ERYieldsParam[i] = er[i];
  m->n_widths = int(min<size_t>(w.size(), 17));
  for (int i = 0; i < m->n_widths; ++i) m->NRERWidthsParam[i] = w[i];
}
}  // namespace

// ---- Analysis ---------------------------------------------------------------------------------
Analysis Analysis::G2() {  // include/NEST/analysis_G2.hh
  Analysis a;
  a.verbosity = 1; a.PrintSubThr = 0; a.MCtruthE = true; a.usePD = 1; a.useS2 = 1;
  a.minS1 = 2.5; a.maxS1 = 120.5; a.numBins = 118; a.minS2 = 200.; a.maxS2 = 9e4;
  a.logMax = 5.; a.logMin = 2.; a.z_step = 1.;
  return a;
}
nb200_analysis Analysis::flat() const {
  nb200_analysis f;
  nb200_analysis_default(0, &f);
  f.usePD = usePD; f.useS2 = useS2; f.numBins = numBins; f.logBins = logBins;
  f.minS1 = minS1; f.maxS1 = maxS1; f.minS2 = minS2; f.maxS2 = maxS2; f.logMin = logMin; f.logMax = logMax;
  f.MCtruthE = MCtruthE; f.MCtruthPos = MCtruthPos; f.z_step = z_step;
  switch (s1CalculationMode) {
    case S1CalculationMode::Full: f.s1mode = NB200_S1_FULL; break;
    case S1CalculationMode::Parametric: f.s1mode = NB200_S1_PARAMETRIC; break;
    case S1CalculationMode::Hybrid: f.s1mode = NB200_S1_HYBRID; break;
    default: f.s1mode = -1; break;  // Waveform: refused by the library
  }
  f.s2mode = (s2CalculationMode == S2CalculationMode::Full) ? NB200_S2_FULL : -1;
  return f;
}

// ---- Device -------------------------------------------------------------------------------------
nb200_ctx* Device::get() {
  static nb200_ctx* ctx = nullptr;
  if (!ctx) {
    int dev = 0;
    if (const char* e = getenv("NB200_DEVICE")) dev = atoi(e);
    int rc = nb200_create(dev, &ctx);
    if (rc) throw runtime_error(string("nest_b200: ") + nb200_last_error(nullptr) + " (there is no CPU fallback)");
  }
  return ctx;
}
void Device::check(int rc) {
  if (rc) {
    const char* m = nb200_last_error(rc == NB200_OK ? nullptr : Device::get());
    throw runtime_error(m && *m ? m : "nest_b200 error");
  }
}

// ---- NESTcalc -----------------------------------------------------------------------------------
NESTcalc::NESTcalc(VDetector* detector) : fdetector(detector) {
  if (!detector) throw runtime_error("NESTcalc needs a detector");
}
const nb200_detector& NESTcalc::flat() {
  if (!flattened) {
    static nb200_flatten_storage store;  // LUT backing store (one detector per process, as the reference)
    nb200_flatten(*fdetector, &fdet, &store);
    flattened = true;
  } else {  // scalars may have been changed through the setters since
    nb200_detector keep = fdet;
    nb200_flatten_storage dummy;
    nb200_map s1 = keep.FitS1, s2 = keep.FitS2, ef = keep.FitEF, tba = keep.FitTBA;
    nb200_flatten_scalars(*fdetector, &fdet);
    fdet.FitS1 = s1; fdet.FitS2 = s2; fdet.FitEF = ef; fdet.FitTBA = tba;
  }
  return fdet;
}
const nb200_run_consts& NESTcalc::prepared(double inField, double z_step) {
  nb200_detector d = flat();
  int rc_ = nb200_prepare(&d, inField, z_step, &rc);
  if (rc_) throw runtime_error(nb200_last_error(nullptr));
  fdet.inGas = d.inGas;
  fdetector->set_inGas(d.inGas != 0);
  rc_field = inField;
  return rc;
}
double NESTcalc::SetDensity(double Kelvin, double bara) {
  nb200_detector d = flat();
  d.T_Kelvin = Kelvin;
  d.p_bar = bara;
  nb200_run_consts c;
  int r = nb200_prepare(&d, 180., 0., &c);
  fdetector->set_inGas(d.inGas != 0);
  if (r) throw runtime_error(nb200_last_error(nullptr));
  return c.rho;
}
vector<double> NESTcalc::SetDriftVelocity_NonUniform(double, double zStep, double Kelvin, double, double dx, double dy) {
  nb200_detector d = flat();
  d.T_Kelvin = Kelvin;
  size_t n = 0;
  if (nb200_drift_table(&d, zStep, dx, dy, nullptr, &n)) throw runtime_error("SetDriftVelocity_NonUniform: bad z step or geometry");
  vector<double> t(n);
  if (nb200_drift_table(&d, zStep, dx, dy, t.data(), &n)) throw runtime_error("SetDriftVelocity_NonUniform failed");
  t.resize(n);
  return t;
}
double NESTcalc::SetDriftVelocity(double Kelvin, double, double eField, double) {
  nb200_detector d = flat();
  d.T_Kelvin = Kelvin;
  d.dt_min = 0.;
  nb200_run_consts c;
  if (nb200_prepare(&d, eField, 0., &c)) throw runtime_error(nb200_last_error(nullptr));
  return c.vD;
}
double NESTcalc::GetDriftVelocity_Liquid(double Kelvin, double eField, double, double, short) {
  const double v = nb200_drift_velocity_liquid(Kelvin, eField);
  if (v != v) throw runtime_error("ERROR: TEMPERATURE OUT OF RANGE (100-230 K)");  // NEST.cpp:2410-2416
  return v;
}
double NESTcalc::GetDriftVelocity(double Kelvin, double Density, double eField, bool inGas, double Bar) {
  if (inGas) throw runtime_error("GetDriftVelocity: gas-phase drift (MagBoltz) is outside the GPU path");
  return GetDriftVelocity_Liquid(Kelvin, eField, Density, Bar);
}
double NESTcalc::GetDensity(double Kelvin, double bara, bool& inGas, uint64_t, double molarMass) {
  int g = inGas ? 1 : 0;
  const double rho = nb200_density(Kelvin, bara, molarMass, &g);
  inGas = g != 0;
  return rho;
}
double NESTcalc::NexONi(double energy, double density) {
  Wvalue w = WorkFunction(density, fdetector->get_molarMass(), fdetector->get_OldW13eV());
  return w.alpha * erf(0.05 * energy);
}
NESTcalc::Wvalue NESTcalc::WorkFunction(double rho, double MolarMass, bool OldW13eV) {
  double alpha = 0.067366 + rho * 0.039693;
  double eDensity = (rho / MolarMass) * 6.0221409e+23 * 54.;
  double Wq = 18.7263 - 1.01e-23 * eDensity;
  if (OldW13eV) Wq *= 1.1716263232;
  return Wvalue{Wq, alpha};
}
vector<double> NESTcalc::CalculateG2(int verbosity) {
  const nb200_run_consts& c = prepared(rc_field > -12345. ? rc_field : 180.);
  vector<double> g(c.g2_params, c.g2_params + 5);
  if (verbosity > 0) {  // banner NEST.cpp:2220-2224; SE mean/width from a 10 000-event single-electron run
    double se_mean = g[2], se_width = 0.;
    nb200_se_stats(Device::get(), &fdet, &c, 10000, 1, &se_mean, &se_width);
    cout << endl << "g1 = " << fdetector->get_g1() << " phd per photon\tg2 = " << g[3] << " phd per electron (e-EE = ";
    cout << g[1] * 100. << "%, SE_mean,width = " << se_mean << "," << se_width << ")\t";
  }
  return g;
}
vector<YieldResult> NESTcalc::GetYields(INTERACTION_TYPE species, const vector<double>& energy, double density,
                                        const vector<double>& dfield, double massNum, double atomNum,
                                        const vector<double>& NRYieldsParam, const vector<double>& ERYieldsParam) {
  if (energy.size() != dfield.size()) throw runtime_error("GetYields: energy and field lists differ in length");
  nb200_model m;
  fill_model(&m, NRYieldsParam, ERYieldsParam, default_NRERWidthsParam);
  m.er_model = -1;
  const bool nrlike = species <= Cf || species == atmNu;
  if (nrlike && nearlyEqual(massNum, 0.)) massNum = double(RandomGen::rndm()->SelectRanXeAtom());  // NEST.cpp:789-795
  m.massNum = massNum;
  m.atomNum = atomNum;
  const size_t n = energy.size();
  if (n) beta_model_m5_notice(m, int(species));
  vector<double> out(6 * n);
  Device::check(nb200_yields(Device::get(), &flat(), &m, int(species), n, energy.data(), dfield.data(), density,
                             NB200_MEM_HOST, out.data()));
  vector<YieldResult> r(n);
  for (size_t i = 0; i < n; ++i)
    r[i] = YieldResult{out[i], out[n + i], out[2 * n + i], out[3 * n + i], out[4 * n + i], out[5 * n + i]};
  return r;
}
YieldResult NESTcalc::GetYields(INTERACTION_TYPE species, double energy, double density, double dfield,
                                double massNum, double atomNum, const vector<double>& NRYieldsParam,
                                const vector<double>& ERYieldsParam) {
  return GetYields(species, vector<double>{energy}, density, vector<double>{dfield}, massNum, atomNum,
                   NRYieldsParam, ERYieldsParam)[0];
}

// ---- the model methods behind GetYields: NEST.hh:256-303 ---------------------------------------------------------
vector<YieldResult> NESTcalc::yields_by(int er_model, int species, const vector<double>& energy, double density,
                                        const vector<double>& dfield, double massNum, double atomNum,
                                        const vector<double>& nr, const vector<double>& er, double multFact) {
  if (energy.size() != dfield.size()) throw runtime_error("GetYield*: energy and field lists differ in length");
  nb200_model m;
  fill_model(&m, nr, er, default_NRERWidthsParam);
  m.er_model = er_model;
  m.multFact = multFact;
  m.massNum = massNum;
  m.atomNum = atomNum;
  const size_t n = energy.size();
  if (n) beta_model_m5_notice(m, species);
  vector<double> out(6 * n);
  Device::check(nb200_yields(Device::get(), &flat(), &m, species, n, energy.data(), dfield.data(), density, NB200_MEM_HOST,
                             out.data()));
  vector<YieldResult> r(n);
  for (size_t i = 0; i < n; ++i)
    r[i] = YieldResult{out[i], out[n + i], out[2 * n + i], out[3 * n + i], out[4 * n + i], out[5 * n + i]};
  return r;
}
vector<YieldResult> NESTcalc::GetYieldGamma(const vector<double>& energy, double density, const vector<double>& dfield,
                                            double multFact) {
  return yields_by(1, gammaRay, energy, density, dfield, 0., 0., default_NRYieldsParam, default_ERYieldsParam, multFact);
}
vector<YieldResult> NESTcalc::GetYieldERWeighted(const vector<double>& energy, double density, const vector<double>& dfield,
                                                 const vector<double>& ERYieldsParam) {
  return yields_by(2, NEST::beta, energy, density, dfield, 0., 0., default_NRYieldsParam, ERYieldsParam, 1.);
}
vector<YieldResult> NESTcalc::GetYieldNR(const vector<double>& energy, double density, const vector<double>& dfield,
                                         double massNum, const vector<double>& NRYieldsParam) {
  // massNum ~ 0: the reference draws one Xe isotope per call (NEST.cpp:789-795, RandomGen::SelectRanXeAtom); so here
  if (nearlyEqual(massNum, 0.)) massNum = double(RandomGen::rndm()->SelectRanXeAtom());
  return yields_by(-1, NR, energy, density, dfield, massNum, 0., NRYieldsParam, default_ERYieldsParam, 1.);
}
vector<YieldResult> NESTcalc::GetYieldIon(const vector<double>& energy, double density, const vector<double>& dfield,
                                          double massNum, double atomNum, const vector<double>& NRYieldsParam) {
  return yields_by(-1, ion, energy, density, dfield, massNum, atomNum, NRYieldsParam, default_ERYieldsParam, 1.);
}
vector<YieldResult> NESTcalc::GetYieldBetaGR(const vector<double>& energy, double density, const vector<double>& dfield,
                                             const vector<double>& ERYieldsParam, double multFact) {
  return yields_by(0, NEST::beta, energy, density, dfield, 0., 0., default_NRYieldsParam, ERYieldsParam, multFact);
}
YieldResult NESTcalc::GetYieldGamma(double energy, double density, double dfield, double multFact) {
  return GetYieldGamma(vector<double>{energy}, density, vector<double>{dfield}, multFact)[0];
}
YieldResult NESTcalc::GetYieldERWeighted(double energy, double density, double dfield, const vector<double>& ERYieldsParam) {
  return GetYieldERWeighted(vector<double>{energy}, density, vector<double>{dfield}, ERYieldsParam)[0];
}
YieldResult NESTcalc::GetYieldNR(double energy, double density, double dfield, double massNum,
                                 const vector<double>& NRYieldsParam) {
  return GetYieldNR(vector<double>{energy}, density, vector<double>{dfield}, massNum, NRYieldsParam)[0];
}
YieldResult NESTcalc::GetYieldIon(double energy, double density, double dfield, double massNum, double atomNum,
                                  const vector<double>& NRYieldsParam) {
  return GetYieldIon(vector<double>{energy}, density, vector<double>{dfield}, massNum, atomNum, NRYieldsParam)[0];
}
YieldResult NESTcalc::GetYieldKr83m(double energy, double density, double dfield, double maxTimeSeparation,
                                    double minTimeSeparation) {
  return yields_by(-1, Kr83m, vector<double>{energy}, density, vector<double>{dfield}, maxTimeSeparation, minTimeSeparation,
                   default_NRYieldsParam, default_ERYieldsParam, 1.)[0];
}
YieldResult NESTcalc::GetYieldBetaGR(double energy, double density, double dfield, const vector<double>& ERYieldsParam,
                                     double multFact) {
  return GetYieldBetaGR(vector<double>{energy}, density, vector<double>{dfield}, ERYieldsParam, multFact)[0];
}
YieldResult NESTcalc::YieldResultValidity(YieldResult& res, const double energy, const double Wq_eV) {  // NEST.cpp:1254-1271
  if (res.PhotonYield > energy / W_SCINT) res.PhotonYield = energy / W_SCINT;
  if (res.ElectronYield > energy / W_SCINT) res.ElectronYield = energy / W_SCINT;
  if (res.PhotonYield < 0.) res.PhotonYield = 0.;
  if (res.ElectronYield < 0.) res.ElectronYield = 0.;
  res.Lindhard = std::max(0., std::min(res.Lindhard, 1.));
  if (energy < 0.001 * Wq_eV / res.Lindhard) {
    res.PhotonYield = 0.;
    res.ElectronYield = 0.;
  }
  return res;
}

// ---- the width model: NEST.cpp:164-246 -----------------------------------------------------------------------------
vector<vector<double>> NESTcalc::Widths(const vector<double>& efield, const vector<double>& elecFrac,
                                        const vector<double>& numQuanta, double density,
                                        const vector<double>& NRERWidthsParam) {
  const size_t n = efield.size();
  if (elecFrac.size() != n || numQuanta.size() != n) throw runtime_error("Widths: lists differ in length");
  nb200_model m;
  fill_model(&m, default_NRYieldsParam, default_ERYieldsParam, NRERWidthsParam);
  vector<double> out(3 * n);
  Device::check(nb200_widths(Device::get(), &flat(), &m, n, efield.data(), elecFrac.data(), numQuanta.data(), density,
                             NB200_MEM_HOST, out.data()));
  vector<vector<double>> r(n, vector<double>(3));
  for (size_t i = 0; i < n; ++i) r[i] = {out[i], out[n + i], out[2 * n + i]};
  return r;
}
double NESTcalc::RecombOmegaNR(double elecFrac, const vector<double>& w) {
  return Widths({100.}, {elecFrac}, {100.}, 2.9, w)[0][0];
}
double NESTcalc::RecombOmegaER(double efield, double elecFrac, double numQuanta, const vector<double>& w) {
  return Widths({efield}, {elecFrac}, {numQuanta}, 2.9, w)[0][1];
}
double NESTcalc::FanoER(double density, double Nq_mean, double efield, const vector<double>& w) {
  return Widths({efield}, {0.5}, {Nq_mean}, density, w)[0][2];
}

// ---- GetSpike / xyResolution: NEST.cpp:2243-2268, 2699-2736 -----------------------------------------------------------
const vector<double>& NESTcalc::GetSpike(int, double dx, double dy, double dz, double, double dS_mid,
                                         const vector<double>& origScint) {
  if (origScint.size() < 9) throw runtime_error("GetSpike: origScint must hold scintillation[0..8]");
  nb200_run_consts c = prepared(rc_field > -12345. ? rc_field : 180.);
  c.vD_middle = dS_mid;
  const double pos[3] = {dx, dy, dz};
  double out[2] = {0., 0.};
  Device::check(nb200_spike(Device::get(), &flat(), &c, 1, pos, origScint.data(), call_seed, call_index++, NB200_MEM_HOST, out));
  newSpike[0] = out[0];
  newSpike[1] = out[1];
  return newSpike;
}
vector<vector<double>> NESTcalc::xyResolution(const vector<double>& x, const vector<double>& y, const vector<double>& A_top) {
  const size_t n = x.size();
  if (y.size() != n || A_top.size() != n) throw runtime_error("xyResolution: lists differ in length");
  vector<double> out(2 * n);
  Device::check(nb200_xy_resolution(Device::get(), &flat(), n, x.data(), y.data(), A_top.data(), call_seed, call_index,
                                    NB200_MEM_HOST, out.data()));
  call_index += n;
  vector<vector<double>> r(n, vector<double>(2));
  for (size_t i = 0; i < n; ++i) r[i] = {out[i], out[n + i]};
  return r;
}
vector<double> NESTcalc::xyResolution(double xPos_mm, double yPos_mm, double A_top) {
  return xyResolution(vector<double>{xPos_mm}, vector<double>{yPos_mm}, vector<double>{A_top})[0];
}

// ---- GetQuanta / FullCalculation: NEST.cpp:248-432, 17-40 ----------------------------------------------
vector<QuantaResult> NESTcalc::GetQuanta(const vector<YieldResult>& yields, double density,
                                         const vector<double>& NRERWidthsParam, double SkewnessER) {
  nb200_model m;
  fill_model(&m, default_NRYieldsParam, default_ERYieldsParam, NRERWidthsParam);  // throws below 13 widths, as NEST.cpp:258-262
  m.er_model = -1;
  m.SkewnessER = SkewnessER;
  const size_t n = yields.size();
  vector<QuantaResult> r(n);
  if (n == 0) return r;
  vector<double> y(6 * n), qd(2 * n);
  vector<int32_t> qi(4 * n);
  for (size_t i = 0; i < n; ++i) {
    y[i] = yields[i].PhotonYield;
    y[n + i] = yields[i].ElectronYield;
    y[2 * n + i] = yields[i].ExcitonRatio;
    y[3 * n + i] = yields[i].Lindhard;
    y[4 * n + i] = yields[i].ElectricField;
    y[5 * n + i] = yields[i].DeltaT_Scint;
  }
  Device::check(nb200_quanta(Device::get(), &flat(), &m, n, y.data(), density, call_seed, call_index, NB200_MEM_HOST,
                             qi.data(), qd.data()));
  call_index += n;
  for (size_t i = 0; i < n; ++i)
    r[i] = QuantaResult{qi[i], qi[n + i], qi[2 * n + i], qi[3 * n + i], qd[i], qd[n + i]};
  return r;
}
QuantaResult NESTcalc::GetQuanta(const YieldResult& yields, double density, const vector<double>& NRERWidthsParam,
                                 double SkewnessER) {
  return GetQuanta(vector<YieldResult>{yields}, density, NRERWidthsParam, SkewnessER)[0];
}
NESTresult NESTcalc::FullCalculation(INTERACTION_TYPE species, double energy, double density, double dfield, double A,
                                     double Z, const vector<double>& NRYieldsParam, const vector<double>& NRERWidthsParam,
                                     const vector<double>& ERYieldsParam, bool do_times) {
  if (do_times)
    throw runtime_error("FullCalculation: photon times are outside the GPU path (nest-b200 DESIGN.md, scope); pass do_times = false");
  if (density < 1.) fdetector->set_inGas(true);  // NEST.cpp:25
  NESTresult result;
  result.yields = GetYields(species, energy, density, dfield, A, Z, NRYieldsParam, ERYieldsParam);
  result.quanta = GetQuanta(result.yields, density, NRERWidthsParam, -999.);
  result.photon_times = photonstream(size_t(result.quanta.photons), 0.0);
  return result;
}

// ---- GetS1: NEST.cpp:1288-1722 ------------------------------------------------------------------------------
vector<vector<double>> NESTcalc::GetS1(const vector<int>& photons, const vector<vector<double>>& truthPos,
                                       const vector<vector<double>>& smearPos, double driftSpeed, double dS_mid,
                                       uint64_t firstEvt, double dfield, const vector<double>& energy,
                                       S1CalculationMode mode) {
  const size_t n = photons.size();
  if (truthPos.size() != n || smearPos.size() != n || energy.size() != n) throw runtime_error("GetS1: lists differ in length");
  int m;
  switch (mode) {
    case S1CalculationMode::Full: m = NB200_S1_FULL; break;
    case S1CalculationMode::Parametric: m = NB200_S1_PARAMETRIC; break;
    case S1CalculationMode::Hybrid: m = NB200_S1_HYBRID; break;
    default: throw runtime_error("GetS1: the Waveform mode is outside the GPU path");
  }
  vector<vector<double>> r(n, vector<double>(9, 0.));
  if (n == 0) return r;
  nb200_run_consts c = prepared(dfield >= 1. ? dfield : 180.);
  c.vD = driftSpeed;
  c.vD_middle = dS_mid;
  vector<int32_t> nph(photons.begin(), photons.end());
  vector<double> pos(6 * n), out(9 * n);
  for (size_t i = 0; i < n; ++i) {
    if (truthPos[i].size() < 3 || smearPos[i].size() < 3) throw runtime_error("GetS1: a position needs x, y, z");
    for (int k = 0; k < 3; ++k) {
      pos[size_t(k) * n + i] = truthPos[i][k];
      pos[size_t(3 + k) * n + i] = smearPos[i][k];
    }
  }
  Device::check(nb200_s1(Device::get(), &flat(), &c, m, n, nph.data(), pos.data(), energy.data(), call_seed, firstEvt,
                         NB200_MEM_HOST, out.data()));
  for (size_t i = 0; i < n; ++i)
    for (int k = 0; k < 9; ++k) r[i][k] = out[size_t(k) * n + i];
  return r;
}
const vector<double>& NESTcalc::GetS1(const QuantaResult& quanta, double truthPosX, double truthPosY, double truthPosZ,
                                      double smearPosX, double smearPosY, double smearPosZ, double driftSpeed, double dS_mid,
                                      INTERACTION_TYPE, uint64_t evtNum, double dfield, double energy, S1CalculationMode mode,
                                      int, vector<int64_t>&, vector<double>&) {
  scintillation = GetS1(vector<int>{quanta.photons}, {{truthPosX, truthPosY, truthPosZ}}, {{smearPosX, smearPosY, smearPosZ}},
                        driftSpeed, dS_mid, evtNum, dfield, vector<double>{energy}, mode)[0];
  return scintillation;
}

// ---- GetS2: NEST.cpp:1724-2063 (Full) --------------------------------------------------------------------
vector<vector<double>> NESTcalc::GetS2(const vector<int>& Ne, const vector<vector<double>>& truthPos,
                                       const vector<vector<double>>& smearPos, const vector<double>& dt, uint64_t firstEvt,
                                       double dfield, const vector<double>& g2_params) {
  const size_t n = Ne.size();
  if (truthPos.size() != n || smearPos.size() != n || dt.size() != n) throw runtime_error("GetS2: lists differ in length");
  if (g2_params.size() < 5) throw runtime_error("GetS2: g2_params needs five values (CalculateG2)");
  vector<vector<double>> r(n, vector<double>(9, 0.));
  if (n == 0) return r;
  nb200_run_consts c = prepared(dfield >= 1. ? dfield : 180.);
  for (int k = 0; k < 5; ++k) c.g2_params[k] = g2_params[k];
  vector<int32_t> ne(Ne.begin(), Ne.end());
  vector<double> pos(5 * n), f(n, dfield), out(9 * n);
  for (size_t i = 0; i < n; ++i) {
    if (truthPos[i].size() < 3 || smearPos[i].size() < 2) throw runtime_error("GetS2: a position needs x, y, z");
    pos[i] = truthPos[i][0];
    pos[n + i] = truthPos[i][1];
    pos[2 * n + i] = truthPos[i][2];
    pos[3 * n + i] = smearPos[i][0];
    pos[4 * n + i] = smearPos[i][1];
  }
  Device::check(nb200_s2(Device::get(), &flat(), &c, n, ne.data(), pos.data(), dt.data(), f.data(), call_seed, firstEvt,
                         NB200_MEM_HOST, out.data()));
  for (size_t i = 0; i < n; ++i)
    for (int k = 0; k < 9; ++k) r[i][k] = out[size_t(k) * n + i];
  return r;
}
const vector<double>& NESTcalc::GetS2(int Ne, double truthPosX, double truthPosY, double truthPosZ, double smearPosX,
                                      double smearPosY, double smearPosZ, double dt, double, uint64_t evtNum, double dfield,
                                      S2CalculationMode mode, int, vector<int64_t>&, vector<double>&,
                                      const vector<double>& g2_params) {
  if (mode != S2CalculationMode::Full) throw runtime_error("GetS2: the Waveform modes are outside the GPU path");
  ionization = GetS2(vector<int>{Ne}, {{truthPosX, truthPosY, truthPosZ}}, {{smearPosX, smearPosY, smearPosZ}},
                     vector<double>{dt}, evtNum, dfield, g2_params)[0];
  return ionization;
}

// ---- LArNEST: src/LArNEST.cpp:17-33 ---------------------------------------------------------------------------
vector<LArNESTResult> LArNEST::FullCalculation(LArInteraction species, const vector<double>& energy,
                                               const vector<double>& efield, double density) {
  if (energy.size() != efield.size()) throw runtime_error("LArNEST::FullCalculation: energy and field lists differ in length");
  if (int(species) > int(LArInteraction::Alpha))
    throw runtime_error("LArNEST: the dE/dx and LET models are outside the GPU path");
  const size_t n = energy.size();
  vector<LArNESTResult> r(n);
  if (n == 0) return r;
  if (density < 1.) fdetector->set_inGas(true);  // LArNEST.cpp:20
  vector<double> y(9 * n), f(4 * n);
  Device::check(nb200_lar(Device::get(), int(species), n, energy.data(), efield.data(), density, call_seed, call_index,
                          NB200_MEM_HOST, y.data(), f.data()));
  call_index += n;
  for (size_t i = 0; i < n; ++i) {
    r[i].yields = LArYieldResult{y[i], y[n + i], y[2 * n + i], y[3 * n + i], y[4 * n + i], y[5 * n + i], y[6 * n + i],
                                 y[7 * n + i], y[8 * n + i]};
    r[i].fluctuations = LArYieldFluctuationResult{f[i], f[n + i], f[2 * n + i], f[3 * n + i]};
  }
  return r;
}
LArNESTResult LArNEST::FullCalculation(LArInteraction species, double energy, double, double efield, double density, bool) {
  return FullCalculation(species, vector<double>{energy}, vector<double>{efield}, density)[0];  // do_times: unused in the reference too (:25-32)
}
LArYieldResult LArNEST::GetYields(LArInteraction species, double energy, double, double efield, double density) {
  if (int(species) > int(LArInteraction::Alpha))
    throw runtime_error("LArNEST: the dE/dx and LET models are outside the GPU path");
  vector<double> y(9);
  Device::check(nb200_lar(Device::get(), int(species), 1, &energy, &efield, density, 0, 0, NB200_MEM_HOST, y.data(), nullptr));
  return LArYieldResult{y[0], y[1], y[2], y[3], y[4], y[5], y[6], y[7], y[8]};
}

// ---- runNESTvec: execNEST.cpp:329-409 ----------------------------------------------------------------
// ---- TestSpectra: include/NEST/TestSpectra.hh:38-73 over nb200_spectrum ------------------------------------------------
namespace {
struct SpectrumBlock {  // the energies drawn ahead for the scalar calls
  int kind = -1;
  double emin = 0., emax = 0., par[8] = {0., 0., 0., 0., 0., 0., 0., 0.};
  double wkey[4] = {0., 0., 0., 0.};  // WIMP: mass, day, xMax, base[0]
  vector<double> e;
  size_t next = 0;
};
SpectrumBlock g_block;
uint64_t g_spec_seed = 0, g_spec_index = 0;
}  // namespace

void TestSpectra::SetSeed(uint64_t seed) {
  g_spec_seed = seed;
  g_spec_index = 0;
  g_block.kind = -1;
}
vector<double> TestSpectra::spectrum_vec(int kind, size_t n, double emin, double emax, const vector<double>& par,
                                         const WIMP_spectrum_prep* wprep, double mass, double day) {
  nb200_source s;
  std::memset(&s, 0, sizeof s);
  s.kind = kind;
  s.eMin = emin;
  s.eMax = emax;
  for (size_t i = 0; i < par.size() && i < 8; ++i) s.par[i] = par[i];
  if (kind == NB200_SRC_WIMP) {
    if (!wprep) throw runtime_error("WIMP_spectrum: no WIMP_spectrum_prep");
    s.species = NB200_WIMP;
    s.wimp_base = wprep->base;
    s.wimp_exponent = wprep->exponent;
    s.wimp_xMax = wprep->xMax;
    s.wimp_divisor = wprep->divisor;
    s.wimp_mass = mass;
    s.wimp_day = day;
  }
  vector<double> out(n);
  NEST::Device::check(nb200_spectrum(NEST::Device::get(), &s, n, g_spec_seed, g_spec_index, NB200_MEM_HOST, out.data()));
  g_spec_index += n;
  return out;
}
namespace {
double next_energy(int kind, double emin, double emax, const vector<double>& par = {},
                   const TestSpectra::WIMP_spectrum_prep* w = nullptr, double mass = 0., double day = 0.) {
  SpectrumBlock& b = g_block;
  double p8[8] = {0., 0., 0., 0., 0., 0., 0., 0.};
  for (size_t i = 0; i < par.size() && i < 8; ++i) p8[i] = par[i];
  const double wk[4] = {mass, day, w ? w->xMax : 0., w ? w->base[0] : 0.};
  const bool same = b.kind == kind && b.emin == emin && b.emax == emax && std::memcmp(b.par, p8, sizeof p8) == 0 &&
                    std::memcmp(b.wkey, wk, sizeof wk) == 0;
  if (!same || b.next >= b.e.size()) {
    b.e = TestSpectra::spectrum_vec(kind, 16384, emin, emax, par, w, mass, day);
    b.kind = kind;
    b.emin = emin;
    b.emax = emax;
    std::memcpy(b.par, p8, sizeof p8);
    std::memcpy(b.wkey, wk, sizeof wk);
    b.next = 0;
  }
  return b.e[b.next++];
}
}  // namespace
double TestSpectra::CH3T_spectrum(double emin, double emax) { return next_energy(NB200_SRC_CH3T, emin, emax); }
double TestSpectra::C14_spectrum(double emin, double emax) { return next_energy(NB200_SRC_C14, emin, emax); }
double TestSpectra::B8_spectrum(double emin, double emax) { return next_energy(NB200_SRC_B8, emin, emax); }
double TestSpectra::AmBe_spectrum(double emin, double emax) { return PowLawFit_spectrum(emin, emax, -0.95351); }
double TestSpectra::PowLawFit_spectrum(double emin, double emax, double exp) {
  return next_energy(NB200_SRC_AMBE, emin, emax, {exp});
}
double TestSpectra::Cf_spectrum(double emin, double emax) { return next_energy(NB200_SRC_CF, emin, emax); }
double TestSpectra::DD_spectrum(double xMin, double xMax, double expFall, double peakFrac, double peakMu, double peakSig,
                                double peakSkew) {
  return next_energy(NB200_SRC_DD, xMin, xMax, {expFall, peakFrac, peakMu, peakSig, peakSkew});
}
double TestSpectra::ppSolar_spectrum(double emin, double emax) { return next_energy(NB200_SRC_PPSOLAR, emin, emax); }
double TestSpectra::atmNu_spectrum(double emin, double emax) { return next_energy(NB200_SRC_ATMNU, emin, emax); }
double TestSpectra::WIMP_dRate(double ER, double mWimp, double day) {
  // the reference draws a random Xe isotope inside (TestSpectra.cpp:278): natural abundances, host Philox stream
  static const int iso[9] = {124, 126, 128, 129, 130, 131, 132, 134, 136};
  static const double cum[9] = {0.090, 0.180, 2.100, 28.54, 32.62, 53.80, 80.69, 91.13, 100.};
  const double u = double(nb200_poisson(1e9, g_spec_seed + 77 + g_spec_index++) % 1000000007ull) / 1000000007. * 100.;
  int A = 136;
  for (int k = 0; k < 9; ++k)
    if (u <= cum[k]) {
      A = iso[k];
      break;
    }
  int ok = 1;
  const double v = nb200_wimp_drate(ER, mWimp, day, A, &ok);
  if (!ok) throw runtime_error("\tThe velocity integral in the WIMP generator broke!!!");
  return v;
}
TestSpectra::WIMP_spectrum_prep TestSpectra::WIMP_prep_spectrum(double mass, double eStep, double day) {
  WIMP_spectrum_prep p;
  int rc = nb200_wimp_prep(mass, eStep, day, g_spec_seed, p.base, p.exponent, &p.integral, &p.xMax, &p.divisor);
  if (rc) throw runtime_error(nb200_last_error(nullptr));
  return p;
}
double TestSpectra::WIMP_spectrum(WIMP_spectrum_prep wprep, double mass, double day) {
  return next_energy(NB200_SRC_WIMP, 0., 0., {}, &wprep, mass, day);
}

NESTObservableArray runNESTvec(VDetector* detector, INTERACTION_TYPE particleType, vector<double> eList,
                               vector<vector<double>> pos3dxyz, double inField, int seed,
                               vector<double> ERYieldsParam, vector<double> NRYieldsParam,
                               vector<double> NRERWidthsParam, S1CalculationMode s1mode, S2CalculationMode s2mode,
                               bool calculate_times) {
  if (calculate_times) throw runtime_error("runNESTvec: photon times are outside the GPU path (calculate_times=false)");
  if (eList.size() != pos3dxyz.size()) throw runtime_error("runNESTvec: eList and pos3dxyz differ in length");
  NESTcalc calc(detector);
  const size_t n = eList.size();
  NESTObservableArray r;
  if (n == 0) return r;
  const nb200_run_consts rc = calc.prepared(inField < 0. ? 180. : inField);
  nb200_model m;
  fill_model(&m, NRYieldsParam, ERYieldsParam, NRERWidthsParam);
  m.er_model = -1;
  m.SkewnessER = -999.;                  // FullCalculation NEST.cpp:33-34
  m.massNum = detector->get_molarMass();  // execNEST.cpp:378-379
  m.atomNum = 54.;
  Analysis a;
  a.s1CalculationMode = s1mode;
  a.s2CalculationMode = s2mode;
  nb200_analysis ana = a.flat();
  vector<double> x(n), y(n), z(n);
  for (size_t i = 0; i < n; ++i) {
    if (pos3dxyz[i].size() < 3) throw runtime_error("runNESTvec: a position needs x, y, z");
    x[i] = pos3dxyz[i][0];
    y[i] = pos3dxyz[i][1];
    z[i] = pos3dxyz[i][2];
  }
  nb200_source s;
  memset(&s, 0, sizeof s);
  s.kind = NB200_SRC_LIST;
  s.species = int(particleType);
  s.fixed_pos = 1;
  s.vec_mode = 1;
  s.inField = inField;
  s.list_mem_kind = NB200_MEM_HOST;
  s.E = eList.data();
  s.x = x.data();
  s.y = y.data();
  s.z = z.data();
  vector<int32_t> nph(n), ne(n);
  vector<vector<double>> s1(9, vector<double>(n)), s2(9, vector<double>(n));
  nb200_soa_out o;
  memset(&o, 0, sizeof o);
  o.mem_kind = NB200_MEM_HOST;
  o.photons = nph.data();
  o.electrons = ne.data();
  for (int k = 0; k < 9; ++k) {
    o.s1[k] = s1[k].data();
    o.s2[k] = s2[k].data();
  }
  beta_model_m5_notice(m, s.species);
  Device::check(nb200_run(Device::get(), &calc.flat(), &m, &ana, &s, &rc, uint64_t(seed), 0, n, &o, nullptr));
  r.energy_kev = eList;
  r.x_mm = x; r.y_mm = y; r.z_mm = z;
  r.n_photons.assign(nph.begin(), nph.end());
  r.n_electrons.assign(ne.begin(), ne.end());
  auto iabs = [&](const vector<double>& v) {  // store_signals: std::abs(int(.)) execNEST.cpp:284-311
    vector<int> q(n);
    for (size_t i = 0; i < n; ++i) q[i] = abs(int(v[i]));
    return q;
  };
  auto dabs = [&](const vector<double>& v) {
    vector<double> q(n);
    for (size_t i = 0; i < n; ++i) q[i] = fabs(v[i]);
    return q;
  };
  r.s1_nhits = iabs(s1[0]); r.s1_nhits_dpe = iabs(s1[1]); r.s1_nhits_thr = iabs(s1[8]);
  r.s1r_phe = dabs(s1[2]); r.s1c_phe = dabs(s1[3]); r.s1r_phd = dabs(s1[4]); r.s1c_phd = dabs(s1[5]);
  r.s1r_spike = dabs(s1[6]); r.s1c_spike = dabs(s1[7]);
  r.s2_Nee = iabs(s2[0]); r.s2_Nph = iabs(s2[1]); r.s2_nhits = iabs(s2[2]); r.s2_nhits_dpe = iabs(s2[3]);
  r.s2r_phe = dabs(s2[4]); r.s2c_phe = dabs(s2[5]); r.s2r_phd = dabs(s2[6]); r.s2c_phd = dabs(s2[7]);
  return r;
}

// ---- execNEST(): execNEST.cpp:412-1506 ---------------------------------------------------------------
namespace nest_b200_cli {
Analysis analysis;
vector<double> NRYieldsParam = default_NRYieldsParam, NRERWidthsParam = default_NRERWidthsParam,
               ERYieldsParam = default_ERYieldsParam;
double multFact = 1.;
}  // namespace nest_b200_cli

int execNEST(VDetector* detector, double numEvts, const string& type, double eMin, double eMax, double inField,
             string position, const string& posiMuon, double fPos, int seed, bool no_seed, double dayNumber) {
  using namespace nest_b200_cli;
  (void)posiMuon;
  (void)no_seed;
  Analysis& A = analysis;
  const int verbosity = A.verbosity;
  try {
    NESTcalc n(detector);
    if (detector->get_TopDrift() <= 0. || detector->get_anode() <= 0. || detector->get_gate() <= 0.) {  // :419-426
      if (verbosity > 0) cerr << "ERROR, unphysical value(s) of position within the detector geometry.";
      return 1;
    }
    const bool muon = (type == "muon" || type == "MIP" || type == "LIP" || type == "mu" || type == "mu-");
    if (nearlyEqual(eMin, -1.) && !muon) eMin = 0.;  // :452-466
    if (nearlyEqual(eMax, -1.) && nearlyEqual(eMin, 0.)) eMax = kHiEregime;
    if (nearlyEqual(eMax, 0.) && !muon) {
      if (verbosity > 0) cerr << "ERROR: The maximum energy (or Kr time sep) cannot be 0 keV (or 0 ns)!" << endl;
      return 1;
    }
    uint64_t useed = (seed == -1) ? uint64_t(time(nullptr)) : uint64_t(int64_t(seed));  // :442-450

    nb200_source S;
    memset(&S, 0, sizeof S);
    nb200_model M;
    fill_model(&M, NRYieldsParam, ERYieldsParam, NRERWidthsParam);
    M.SkewnessER = 0.;  // execNEST.cpp:1065
    M.er_model = -1;
    M.massNum = detector->get_molarMass();
    M.atomNum = 0.;
    vector<double> wbase(100), wexp(100);
    INTERACTION_TYPE type_num;
    // type aliases :468-576
    if (type == "NR" || type == "neutron" || type == "-1") type_num = NR;
    else if (type == "WIMP") {
      if (eMin < 0.44) {
        if (verbosity > 0) cerr << "WIMP mass too low, you're crazy!" << endl;
        return 1;
      }
      type_num = WIMP;
      double integral = 0.;
      int rc_ = nb200_wimp_prep(eMin, A.E_step, dayNumber, useed, wbase.data(), wexp.data(), &integral, &S.wimp_xMax,
                                &S.wimp_divisor);
      if (rc_) throw runtime_error(nb200_last_error(nullptr));
      S.wimp_base = wbase.data();
      S.wimp_exponent = wexp.data();
      S.wimp_mass = eMin;
      S.wimp_day = dayNumber;
      numEvts = double(nb200_poisson(integral * 1.0 * numEvts * eMax / 1e-36, useed));
    } else if (type == "B8" || type == "Boron8" || type == "8Boron" || type == "8B" || type == "Boron-8") {
      type_num = B8;
      numEvts = double(nb200_poisson(0.0026 * numEvts, useed));
    } else if (type == "DD" || type == "D-D") type_num = DD;
    else if (type == "AmBe") type_num = AmBe;
    else if (type == "Cf" || type == "Cf252" || type == "252Cf" || type == "Cf-252") type_num = Cf;
    else if (type == "ion" || type == "nucleus" || type == "alpha") {
      type_num = ion;
      if (type == "alpha") {
        M.atomNum = 2;
        M.massNum = 4;
      } else {  // :500-505
        cerr << "Atomic Number: ";
        cin >> M.atomNum;
        cerr << "Mass Number: ";
        cin >> M.massNum;
      }
      if (M.atomNum == 54.) type_num = NR;  // ATOM_NUM :506
    } else if (type == "gamma" || type == "gammaRay" || type == "x-ray" || type == "xray" || type == "xRay" ||
               type == "X-ray" || type == "Xray" || type == "XRay")
      type_num = gammaRay;
    else if (type == "Kr83m" || type == "83mKr" || type == "Kr83") type_num = Kr83m;
    else if (type == "CH3T" || type == "tritium") type_num = CH3T;
    else if (type == "C14" || type == "Carbon14" || type == "14C" || type == "C-14" || type == "Carbon-14")
      type_num = C14;
    else if (type == "beta" || type == "ER" || type == "Compton" || type == "compton" || type == "electron" ||
             type == "e-")
      type_num = NEST::beta;
    else if (type == "pp" || type == "ppsolar" || type == "ppSolar" || type == "pp_Solar" || type == "pp_solar" ||
             type == "pp-Solar" || type == "pp-solar") {
      type_num = ppSolar;
      numEvts = double(nb200_poisson(0.0011794 * numEvts, useed));  // counts per kg-day, 0-250 keVee :535-537
    } else if (type == "atmNu" || type == "AtmNu" || type == "atm_Nu" || type == "Atm_Nu" || type == "atm-Nu" ||
               type == "Atm-Nu" || type == "atm_nu" || type == "atm-nu") {
      type_num = atmNu;
      numEvts = double(nb200_poisson(1.5764e-7 * numEvts, useed));  // :541-542
    } else if (muon || type == "fullGamma") {
      if (verbosity > 0)
        cerr << "ERROR: particle type " << type << " is not part of the GPU path (nest-b200 DESIGN.md, scope); "
             << "use the reference execNEST for it." << endl;
      return 1;
    } else {
      if (verbosity > 0)
        cerr << "UNRECOGNIZED PARTICLE TYPE!! VALID OPTIONS ARE:\nNR or neutron,\nWIMP,\nB8 or Boron8 or 8Boron or "
                "8B or Boron-8,\nDD or D-D,\nAmBe,\nCf or Cf252 or 252Cf or Cf-252,\nion or nucleus,\nalpha,\ngamma or "
                "gammaRay,\nx-ray or xray or xRay or X-ray or Xray or XRay,\nKr83m or 83mKr or Kr83,\nCH3T or "
                "tritium,\nCarbon14 or 14C or C14 or C-14 or Carbon-14,\nbeta or ER or Compton or compton or "
                "electron or e-,\npp or ppSolar with many various underscore, hyphen and capitalization "
                "permutations permitted,\natmNu\n";
      return 1;
    }
    if (numEvts < 1. && type_num != WIMP && type_num != B8 && type_num != ppSolar && type_num != atmNu) {  // :578-581
      if (verbosity > 0) cerr << "ERROR, you must simulate at least one event, or non-zero kg*days" << endl;
      return 1;
    }
    double maxTimeSep = DBL_MAX;
    if (type_num == Kr83m) {  // :583-601
      if (((eMin > 9.35 && eMin < 9.45) || (eMin >= 32. && eMin < 32.2) || (eMin > 41. && eMin <= 41.6)) &&
          eMin != eMax) {
        maxTimeSep = eMax;
        if (eMax <= 0.) {
          if (verbosity > 0) cerr << "Max t sep must be +." << endl;
          return 1;
        }
      } else {
        if (verbosity > 0)
          cerr << "ERROR: For Kr83m, put E_min as 9.4, 32.1, or 41.5 keV and E_max as the max time-separation [ns] "
                  "between the two decays please."
               << endl;
        return 1;
      }
    }
    if ((eMin < 10. || eMax < 10.) && type_num == gammaRay && verbosity > 0) {  // :603-610
      cerr << "WARNING: Typically beta model works better for ER BG at low energies as in a WS." << endl;
      cerr << "ER data is often best matched by a weighted average of the beta & gamma models." << endl;
    }

    // per-run constants :612-631, 878-890
    nb200_run_consts rc;
    try {
      rc = n.prepared(inField, A.z_step);
    } catch (exception& e) {
      if (verbosity > 0) cerr << e.what() << endl;
      return 1;
    }
    const double rho = rc.rho, Wq_eV = rc.Wq_eV;
    if (inField != -1. && rc.vD_middle == 0.1) {
      // NEST.cpp:2504-2519: GetDriftVelocity_Liquid found a non-positive speed and took 0.1 mm/us (the per-run constants
      // take the same value); the reference says so at each of its two calls for a uniform field, whatever the verbosity
      for (int k = 0; k < 2; ++k) {
        cerr << "\nWARNING: DRIFT SPEED NON-POSITIVE. Setting to 0.1 mm/us\t"
             << "Line Number ~1950 of NEST.cpp, in function NESTcalc::GetDriftVelocity_Liquid\t"
             << "Stop bothering Matthew about this, and fix underlying cause in your code!\n";
        if (inField < 1e2 && inField >= 1.) cerr << "FIELD MAY BE TOO LOW. ";
        else if (inField > 1e4) cerr << "FIELD MAYBE TOO HIGH. ";
        else cerr << "Something unknown went wrong: are you in a noble element?? ";
        cerr << "EF = " << inField << " V/cm. T = " << detector->get_T_Kelvin() << " Kelvin" << endl;
      }
    }
    n.CalculateG2(verbosity);
    const double g2 = fabs(rc.g2_params[3]), g1 = detector->get_g1();
    // the mean yields at the top of the energy range, for the warnings about the S1 / S2 windows :635-668, 1441-1452
    YieldResult yieldsMax{};
    {
      const double centralZ = (detector->get_gate() * 0.8 + detector->get_cathode() * 1.03) / 2.;
      double centralField = detector->FitEF(0.0, 0.0, centralZ);
      if (inField != -1.) centralField = inField;
      try {
        if (type_num == WIMP)
          yieldsMax = n.GetYields(NR, 25.0, rho, centralField, detector->get_molarMass(), M.atomNum, NRYieldsParam, ERYieldsParam);
        else if (type_num == B8)
          yieldsMax = n.GetYields(NR, 5.00, rho, centralField, detector->get_molarMass(), M.atomNum, NRYieldsParam, ERYieldsParam);
        else {
          double energyMaximum = eMax < 0. ? 1. / fabs(eMax) : eMax;
          if (!energyMaximum || std::isnan(energyMaximum)) energyMaximum = 1.;
          if (type_num == Kr83m)
            yieldsMax = n.GetYields(Kr83m, eMin, rho, centralField, 400., 100., NRYieldsParam, ERYieldsParam);
          else
            yieldsMax = n.GetYields(type_num, energyMaximum, rho, centralField, M.massNum, M.atomNum, NRYieldsParam, ERYieldsParam);
        }
      } catch (exception&) {
        yieldsMax = YieldResult{};  // these yields only feed warnings: a refusal here must not end the run
      }
    }
    if ((g1 * yieldsMax.PhotonYield) > (2. * A.maxS1) && eMin != eMax && type_num != Kr83m && verbosity > 0)
      cerr << "\nWARNING: Your energy maximum may be too high given your maxS1.\n";
    if (type_num < 6) M.massNum = 0;  // :670
    if (type_num == Kr83m) {           // :671, 1052: max / min time separation ride in massNum / atomNum
      M.massNum = maxTimeSep;
      M.atomNum = 0.;  // minTimeSeparation :38
    }

    // source kind :681-790
    const bool mono = (nearlyEqual(eMin, eMax) && eMin >= 0. && eMax > 0.) || type_num == Kr83m;
    S.species = int(type_num);
    S.eMin = eMin;
    S.eMax = eMax;
    S.inField = inField;
    if (mono) S.kind = NB200_SRC_MONO;
    else if (type_num == CH3T) {
      S.kind = NB200_SRC_CH3T;
      if ((min(eMax, 18.5898) != 18.5898 || max(eMin, 0.) != 0.) && verbosity > 0)
        cerr << "WARNING: Recommended energy range is 0 to " << 18.5898 << " keV" << endl;  // TestSpectra.cpp:37-41
    } else if (type_num == B8) S.kind = NB200_SRC_B8;
    else if (type_num == C14) {
      S.kind = NB200_SRC_C14;
      if ((min(eMax, 156.) != 156. || max(eMin, 0.) != 0.) && verbosity > 0)
        cerr << "WARNING: Recommended energy range is 0 to " << 156. << " keV" << endl;  // TestSpectra.cpp:70-72
    } else if (type_num == Cf) S.kind = NB200_SRC_CF;
    else if (type_num == ppSolar) S.kind = NB200_SRC_PPSOLAR;
    else if (type_num == atmNu) S.kind = NB200_SRC_ATMNU;
    else if (type_num == AmBe) {
      S.kind = NB200_SRC_AMBE;
      S.par[0] = -0.95351;
    } else if (type_num == DD) {
      S.kind = NB200_SRC_DD;
      const double p[5] = {13., 0.12, 71.2, 20., -20.5};  // :705
      memcpy(S.par, p, sizeof p);
    } else if (type_num == WIMP) S.kind = NB200_SRC_WIMP;
    else {
      if (eMin < 0.) return 1;
      if (eMax > 0.) S.kind = (eMax > eMin) ? NB200_SRC_FLAT : NB200_SRC_GAUSS;
      else {
        if (nearlyEqual(eMin, 0.)) return 1;
        S.kind = NB200_SRC_EXP;
      }
    }
    if (type == "ER") {  // :1032-1048
      if (verbosity > 0)
        cerr << "CAUTION: Are you sure you don't want beta model instead of ER? This is a recomb-enhanced "
                "(Qy-reduced) version of it for non-betas e.g. EC, DEC. Seed is now Q reduction."
             << endl
             << ">1 means increase, and negative number means gamma model; use 0 for a weighted average of the "
                "gamma, beta models"
             << endl;
      if (multFact > 0.) { M.er_model = 0; M.multFact = multFact; }
      else if (multFact < 0.) { M.er_model = 1; M.multFact = -multFact; }
      else M.er_model = 2;
    } else if (seed < 0 && seed != -1 && type_num <= 5)
      M.massNum = detector->get_molarMass();  // :1050-1051

    // position :806-855
    if (nearlyEqual(fPos, -1.)) S.fixed_pos = 0;
    else {
      double px = 0., py = 0., pz = 0.;
      size_t c1 = position.find(','), c2 = position.rfind(',');
      if (c1 == string::npos || c1 == c2) {
        if (verbosity > 0) cerr << "ERROR: position must be -1 or x,y,z" << endl;
        return 1;
      }
      px = stof(position.substr(0, c1));
      py = stof(position.substr(c1 + 1, c2 - c1 - 1));
      pz = stof(position.substr(c2 + 1));
      const bool ranXY = nearlyEqual(py, -999.), ranZ = nearlyEqual(pz, -1.);
      if (!ranXY) {
        if ((px * px + py * py) > detector->get_radius() * detector->get_radius() && verbosity > 0)
          cerr << "WARNING: outside fiducial radius." << endl;
        if ((px * px + py * py) > detector->get_radmax() * detector->get_radmax()) {
          if (verbosity > 0) cerr << "\nERROR: outside physical radius!!!" << endl;
          return EXIT_FAILURE;
        }
      }
      S.fixed_pos = ranXY ? (ranZ ? 0 : 3) : (ranZ ? 2 : 1);
      S.xyz[0] = px;
      S.xyz[1] = py;
      S.xyz[2] = pz;
    }
    if ((inField < 0. && inField != -1.) || detector->get_E_gas() < 0.) {  // :863-868
      if (verbosity > 0)
        cerr << "\nERROR: Neg field is not permitted. We don't simulate field dir (yet). Put in magnitude.\n";
      return 1;
    }
    if (inField > 12e3 && verbosity > 0)
      cerr << "\nWARNING: Your field is >12,000 V/cm. No data out here. Are you sure about this?\n";

    if (mono && A.numBins == 1 && A.MCtruthE) {  // :903-909
      A.MCtruthE = false;
      if (verbosity > 0) fprintf(stderr, "Simulating a mono-E peak; setting MCtruthE false.\n");
    }
    // header :891-947
    if (verbosity > 0) {
      cout << "Density = " << rho << " g/mL" << "\t";
      cout << "central vDrift = " << rc.vD_middle << " mm/us\n";
      cout << "\t\t\tPRO TIP: neg s2_thr in detector->S2bot!\tW = " << Wq_eV
           << " eV\tNegative numbers are flagging things below threshold!   phe=(1+P_dphe)*phd & "
              "phd=phe/(1+P_dphe)\n";
      if (type_num == Kr83m && (eMin < 10. || eMin > 40.)) fprintf(stdout, "t [ns]\t\t");  // :900-901
      if (eMax > kHiEregime) fprintf(stdout, "Energy [keV]");
      else fprintf(stdout, A.MCtruthE ? "E_truth [keV]" : "E_recon [keV]");
      if (seed == -2)  // :918-935 (the dE/dx basis is not part of this path)
        printf("\tfield [V/cm]\ttDrift [us]\tX,Y,Z [mm]\tNph\t\tNe-\t\tS1 [PE or phe]\tS1_3Dcor [phd]\tspikeC(NON-INT)\tNe-"
               "Extr\tS2_rawArea [PE]\tS2_3Dcorr [phd]\n");
      else if (verbosity >= 2)
        printf("\tfield [V/cm]\ttDrift [us]\tZ [mm]\tNph\tNe-\tS1 [PE or phe]\tS1_3Dcor [phd]\tspikeC(NON-INT)\tNe-Extr"
               "\tS2_rawArea [PE]\tS2_3Dcorr [phd]\n");
      else
        printf("\tfield [V/cm]\ttDrift [us]\tX,Y,Z [mm]\tNph\tNe-\tS1 [PE or phe]\tS1_3Dcor [phd]\tspikeC(NON-INT)\tNe-"
               "Extr\tS2_rawArea [PE]\tS2_3Dcorr [phd]\n");
    }

    // the event loop :676-1414 on the device, in chunks; rows printed as the reference does :1346-1409
    nb200_analysis ana = A.flat();
    const uint64_t N = uint64_t(numEvts);
    const uint64_t chunk = 1u << 20;
    const uint64_t cap = min<uint64_t>(N, chunk);
    vector<double> cE(cap), cEtrue(cap), cF(cap), cDt(cap), cSx(cap), cSy(cap), cZ(cap), sig1(cap), sig2(cap);
    vector<int32_t> cNph(cap), cNe(cap);
    const bool means = seed < 0 && seed != -1;  // :1367: rows carry the mean yields in place of the quanta
    vector<double> cYph(means ? cap : 0), cYne(means ? cap : 0);
    vector<vector<double>> s1(9, vector<double>(cap)), s2(9, vector<double>(cap));
    vector<double> cDeltaT(cap), cKeVee(cap);
    double keVee = 0.;
    vector<double> all1, all2, allE;  // mono-energetic summaries need every event (GetBand resol, GetEnergyRes)
    nb200_soa_out o;
    memset(&o, 0, sizeof o);
    o.mem_kind = NB200_MEM_HOST;
    o.signalE = cE.data(); o.energy_kev = cEtrue.data(); o.field = cF.data(); o.driftTime = cDt.data();
    o.smear_x = cSx.data(); o.smear_y = cSy.data(); o.z_mm = cZ.data(); o.photons = cNph.data();
    o.electrons = cNe.data(); o.signal1 = sig1.data(); o.signal2 = sig2.data();
    o.deltaT_ns = cDeltaT.data(); o.keVee = cKeVee.data();
    if (means) { o.Nph_mean = cYph.data(); o.Ne_mean = cYne.data(); }
    for (int k = 0; k < 9; ++k) {
      o.s1[k] = s1[k].data();
      o.s2[k] = s2[k].data();
    }
    vector<uint64_t> acc(ana.numBins), rej(ana.numBins), h2(size_t(ana.numBins) * ana.logBins);
    vector<int64_t> sS1(ana.numBins), sY(ana.numBins), sY2(ana.numBins), sY3(ana.numBins);
    nb200_band_hist H;
    memset(&H, 0, sizeof H);
    H.numBins = ana.numBins; H.logBins = ana.logBins;
    H.accepted = acc.data(); H.rejected = rej.data(); H.sumS1 = sS1.data(); H.sumY = sY.data();
    H.sumY2 = sY2.data(); H.sumY3 = sY3.data(); H.hist2d = h2.data();
    bool warnedCathode = false;
    double lastField = inField;  // the reference's `field` after its event loop: the last event's
    beta_model_m5_notice(M, S.species);
    for (uint64_t done = 0; done < N; done += chunk) {
      const uint64_t m = min<uint64_t>(chunk, N - done);
      int rc_ = nb200_run(Device::get(), &n.flat(), &M, &ana, &S, &rc, useed, done, m, &o, mono ? nullptr : &H);
      if (rc_) {  // :1410-1413
        if (verbosity > 0) cerr << nb200_last_error(Device::get()) << endl;
        return 1;
      }
      for (uint64_t j = 0; j < m; ++j) {
        keVee += cKeVee[j];
        if (verbosity > 0) {  // per event, as the reference: :869-876
          if (nearlyEqual(cF[j], 0.) || std::isnan(cF[j]))
            cerr << "\nWARNING: A LITERAL ZERO (or undefined) FIELD MAY YIELD WEIRD RESULTS. USE A SMALL VALUE INSTEAD.\n";
          if (cF[j] > 12e3 || detector->get_E_gas() > 17e3)
            cerr << "\nWARNING: Your field is >12,000 V/cm. No data out here. Are you sure about this?\n";
          // :968-971, first event (for a FitEF field the drift speed at the centre stands in for the event's own)
          if (done + j == 0 && detector->get_dt_max() > (detector->get_TopDrift() - 0.) / rc.vD_middle && cF[j] >= 1.)
            cerr << "WARNING: dt_max is greater than max possible" << endl;
        }
        lastField = cF[j];
        if (mono && done + j < 10 && verbosity > 0) {  // :1160-1168
          if (s1[3][j] > A.maxS1 || s1[5][j] > A.maxS1 || s1[7][j] > A.maxS1)
            cerr << "WARNING: Some S1 pulse areas are greater than maxS1" << endl;
          if (s2[5][j] > A.maxS2 || s2[7][j] > A.maxS2)
            cerr << "WARNING: Some S2 pulse areas are greater than maxS2" << endl;
        }
        if (mono) {
          all1.push_back(sig1[j]);
          all2.push_back(sig2[j]);
          allE.push_back(cE[j]);
        }
        if (cZ[j] < detector->get_cathode() && verbosity > 0 && !warnedCathode) {  // :1326-1334
          warnedCathode = true;
          fprintf(stderr, "gamma-X i.e. MSSI may be happening. This may be why even high-E eff is <100%%. Check your "
                          "cathode position definition.\n\n");
        }
        bool print = (A.PrintSubThr == 1);
        if (!print) {  // :1336-1358
          double s1min[3] = {kPHE_MIN, kPHE_MIN, kPHE_MIN}, s1max[3] = {DBL_MAX, DBL_MAX, DBL_MAX};
          double s2min = kPHE_MIN, s2max = DBL_MAX;
          if (A.PrintSubThr == -1) {
            if (A.usePD >= 0 && A.usePD <= 2) {
              s1min[A.usePD] = A.minS1;
              s1max[A.usePD] = A.maxS1;
            }
            s2max = A.maxS2;
            s2min = detector->get_s2_thr();
          }
          print = true;
          const double lo1[9] = {kPHE_MIN, kPHE_MIN, kPHE_MIN, s1min[0], kPHE_MIN, s1min[1], kPHE_MIN, s1min[2], kPHE_MIN};
          const double hi1[9] = {DBL_MAX, DBL_MAX, DBL_MAX, s1max[0], DBL_MAX, s1max[1], DBL_MAX, s1max[2], DBL_MAX};
          for (int k = 0; k < 9 && print; ++k) {
            print = s1[k][j] > lo1[k] && s1[k][j] < hi1[k];
            double lo2 = (k == 4) ? s2min : kPHE_MIN, hi2 = (k == 7) ? s2max : DBL_MAX;
            print = print && s2[k][j] > lo2 && s2[k][j] < hi2;
          }
        }
        if (!print) continue;
        const double keV = A.MCtruthE ? cEtrue[j] : cE[j];
        if (type_num == Kr83m && verbosity > 0 && (eMin < 10. || eMin > 40.)) printf("%.6f\t", cDeltaT[j]);  // :1360-1361
        if (means)
          printf("%.6f\t%.6f\t%.6f\t%.0f, %.0f, %.0f\t%lf\t%lf\t", keV, cF[j], cDt[j], cSx[j], cSy[j], cZ[j], cYph[j],
                 cYne[j]);
        else if (verbosity >= 2)
          printf("%.6f\t%.6f\t%.6f\t%.0f\t%d\t%d\t", keV, cF[j], cDt[j], cZ[j], cNph[j], cNe[j]);
        else
          printf("%.6f\t%.6f\t%.6f\t%.0f, %.0f, %.0f\t%d\t%d\t", keV, cF[j], cDt[j], cSx[j], cSy[j], cZ[j], cNph[j],
                 cNe[j]);
        if (keV > 10. * kHiEregime || s1[5][j] > A.maxS1 || s2[7][j] > A.maxS2) {  // :1390-1396
          printf("%e\t%e\t%e\t", s1[2][j], s1[5][j], s1[7][j]);
          printf("%lli\t%e\t%e\n", (long long int)s2[0][j], s2[4][j], s2[7][j]);
        } else {
          printf("%.6f\t%.6f\t%.6f\t", s1[2][j], s1[5][j], s1[7][j]);
          printf("%i\t%.6f\t%.6f\n", (int)s2[0][j], s2[4][j], s2[7][j]);
        }
      }
    }

    if (verbosity >= 0) {  // :1422-1504
      if (!mono) {
        vector<double> band(size_t(ana.numBins) * 7);
        nb200_band_finalize(&ana, &H, band.data());
        fprintf(stderr, "Bin Center\tBin Actual\tHist Mean\tMean Error\tHist Sigma\tSkewness\t\tEff[%%>thr]\n");
        bool warned = false;  // (the reference sets eMax = -999 after the first hint: one per run)
        for (int j = 0; j < ana.numBins; ++j) {
          const double* b = &band[7 * size_t(j)];
          fprintf(stderr, "%lf\t%lf\t%lf\t%lf\t%lf\t%lf\t\t%lf\n", b[0], b[1], b[2], b[4], b[3], b[6], b[5] * 100.);
          bool bad = false;
          for (int k = 0; k < 6; ++k) bad = bad || b[k] <= 0.0 || std::isnan(b[k]);
          if (bad && !warned) {  // :1435-1454
            if (verbosity > 0) {
              if (((g1 * yieldsMax.PhotonYield) < A.maxS1 || (g2 * yieldsMax.ElectronYield) < A.maxS2) && j != 0)
                cerr << "WARNING: Insufficient number of high-energy events to populate highest bins is likely. Decrease "
                        "maxS1 and/or increase maxS2. Also: up the stats, change E range\n";
              else
                cerr << "WARNING: Insufficient number of low-energy events to populate lowest bins is likely. Increase "
                        "minS1 and/or decrease minS2,s2_thr. Up the stats, change E range\n";
            }
            warned = true;
          }
        }
      } else {  // GetBand(resol=true) + GetEnergyRes :1456-1503, 1582-1644
        // GetBand(S1s, S2s, resol = true, nFold): one bin of infinite width (:1533-1536)
        const int nFold = detector->get_coinLevel();
        const double border = (ana.useS2 == 2) ? ana.minS2 : ana.minS1, binWidth = DBL_MAX;
        const double s1c = border + binWidth / 2.;
        const double ReqS1 = nFold == 0 ? -DBL_MAX : 0.;
        double b0 = 0., b1 = 0., b2 = 0., b3 = 0.;
        vector<double> sel;
        for (size_t i = 0; i < all1.size(); ++i) {
          if ((fabs(all1[i]) > (s1c - binWidth / 2.) && fabs(all1[i]) <= (s1c + binWidth / 2.)) || !nFold) {
            if (all1[i] >= ReqS1 && all2[i] >= 0.) {
              sel.push_back(all2[i]);
              b2 += all2[i];
              b0 += all1[i];
            }
          }
        }
        const double total = double(all1.size());
        double numPts = double(sel.size());
        if (numPts <= 0) {  // :1585-1588
          for (size_t i = 0; i < all1.size(); ++i) b0 += fabs(all1[i]);
          numPts = total;
        }
        b0 /= numPts;
        b1 /= numPts;
        b2 /= numPts;
        if (numPts > double(sel.size())) numPts = double(sel.size());
        for (size_t i = 0; i < sel.size(); ++i) b3 += pow(sel[i] - b2, 2.);
        for (size_t i = 0; i < all1.size(); ++i)
          if (all1[i] > ReqS1 && all2[i] > 0.0) b1 += pow(all1[i] - b0, 2.);
        b3 = sqrt(b3 / (numPts - 1.));
        b1 = sqrt(b1 / (numPts - 1.));
        const double m1 = b0, v1 = b1, m2 = b2, v2 = b3;
        // GetEnergyRes :1623-1644
        double eN = 0, eM = 0, eV = 0;
        for (double e : allE)
          if (e > 0.) {
            eM += e;
            ++eN;
          }
        eM /= eN;
        for (double e : allE)
          if (e > 0.) eV += pow(eM - e, 2.);
        eV = sqrt(eV / (eN - 1.));
        const bool nrlike = type_num < 7;
        if (verbosity > 0)
          fprintf(stderr, nrlike ? "S1 Mean\t\tS1 Res [%%]\tS2 Mean\t\tS2 Res [%%]\tEc [keVnr]\tEc Res[%%]\tEff[%%>thr]\tEc "
                                   "[keVee]\n"
                                 : "S1 Mean\t\tS1 Res [%%]\tS2 Mean\t\tS2 Res [%%]\tEc Mean\t\tEc Res[%%]\tEff[%%>thr]\n");
        fprintf(stderr, "%lf\t%lf\t%lf\t%lf\t%lf\t%lf\t%lf\t", m1, v1 / m1 * 100., m2, v2 / m2 * 100., eM, eV / eM * 100.,
                (eN / total) * 100.);
        if (nrlike)  // :1465, 1477-1478
          fprintf(stderr, "%lf\n", (keVee / numEvts) / (eN / total));
        else
          fprintf(stderr, "\n");
        if ((m1 <= 0.0 || v1 <= 0.0 || m2 <= 0.0 || v2 <= 0.0 || std::isnan(m1) || std::isnan(v1) || std::isnan(m2) ||
             std::isnan(v2)) &&
            lastField >= 1.) {  // :1481-1495
          if (verbosity > 0) {
            if (numEvts > 1)
              cerr << "CAUTION: YOUR S1 and/or S2 MIN and/or MAX may be set to be too restrictive, please check.\n";
            else
              cerr << "CAUTION: Poor stats. You must have at least 2 events to calculate S1 and S2 and E "
                      "resolutions.\n";
          }
        } else if ((nearlyEqual(eM, eMin) || nearlyEqual(eM, eMax) || eV <= 1E-6) && lastField >= 1.)
          if (verbosity > 0)
            cerr << "If your energy resolution is 0% then you probably still have MC truth energy on." << endl;
      }
    }
    return 0;
  } catch (exception& e) {
    if (analysis.verbosity > 0) cerr << e.what() << endl;
    return 1;
  }
}

// ---- main(): execNEST.cpp:40-256 (the non-loopNEST branch) ------------------------------------------
int nest_b200_cli::main_like(VDetector* detector, int argc, char** argv) {
  Analysis& A = analysis;
  if (A.verbosity > 0) cerr << "*** Detector definition message ***" << endl;
  if (A.verbosity > 0) cerr << "You are currently using the " << detector->getName() << " detector." << endl << endl;
  if (argc < 7) {
    cout << "This program takes 6 (or 7) inputs, with Z position in mm from bottom of detector:" << endl;
    cout << "\t./execNEST numEvts type_interaction E_min[keV] E_max[keV] field_drift[V/cm] x,y,z-position[mm] "
            "{optional:seed}"
         << endl << endl;
    cout << "For Kr83m time-dependent 9.4, 32.1, or 41.5 keV yields: " << endl;
    cout << "\t ./execNEST numEvents Kr83m Energy[keV] maxTimeDiff[ns] field_drift[V/cm] x,y,z-position[mm] "
            "{optional:seed}"
         << endl << endl;
    cout << "For 8B or pp or atmNu, numEvts is kg-days of exposure with everything else same. For WIMPs:" << endl;
    cout << "\t./execNEST exposure[kg-days] {WIMP} m[GeV] x-sect[cm^2] field_drift[V/cm] x,y,z-position[mm] "
            "{optional:seed}"
         << endl << endl;
    return 1;
  }
  double numEvts = atof(argv[1]);
  if (numEvts <= 0.) {
    if (A.verbosity > 0) cerr << "ERROR, you must simulate at least one event, or non-zero kg*days" << endl;
    return 1;
  }
  string type = argv[2];
  double eMin = atof(argv[3]), eMax = atof(argv[4]), inField = atof(argv[5]);
  string position = argv[6], posiMuon = argv[4];
  double fPos = atof(argv[6]);
  int seed = 1;
  bool no_seed = false;
  if (argc == 8) {  // the 7th argument is both seed and multFact (:177-180)
    multFact = atof(argv[7]);
    seed = atoi(argv[7]);
  } else
    no_seed = true;
  // the CLI's own model vectors (:185-249), not the NEST.hh defaults
  NRERWidthsParam = {0.4, 0.4, 0.04, 0.50, 0.19, 2.25, -0.00050, 0.03290, 2.629, 0.2964, -.3343, 0.0000, 0.0000};
  NRYieldsParam = {11., 1.1, 0.0480, -0.0533, 12.6, 0.3, 2., 0.3, 2., 0.5, 1.0, 1.0};
  ERYieldsParam = default_ERYieldsParam;
  return execNEST(detector, numEvts, type, eMin, eMax, inField, position, posiMuon, fPos, seed, no_seed, 0.00);
}
