// NEST_b200.hh -- host-side mirror of the reference's operator API for the accelerated path.
//
// Same names, argument meaning and error behaviour as the reference (NESTCollaboration/nest):
//   VDetector              include/Detectors/VDetector.hh:18-225  (getters/setters + Fit* virtuals)
//   INTERACTION_TYPE, YieldResult, S1/S2CalculationMode   include/NEST/NEST.hh:133-178
//   NESTObservableArray, runNESTvec(), execNEST()          include/NEST/execNEST.hh:10-87
//   NESTcalc::{SetDensity, SetDriftVelocity, CalculateG2, WorkFunction, GetYields}  NEST.hh:214-457
// so that a program written against the reference recompiles against this header and runs the
// event loop on the GPU through the C-ABI (include/nest_b200.h).  User detectors subclass VDetector
// exactly as before; their virtual position maps are tabulated once (nest_b200_flatten.hh).
#pragma once
// the standard headers NEST.hh brings in ahead of a detector header (include/NEST/NEST.hh:70-79, RandomGen.hh:10-13)
#include <algorithm>
#include <array>
#include <cassert>
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <numeric>
#include <random>
#include <string>
#include <vector>

#include "nest_b200.h"

// What detector headers written for the reference take from NEST.hh, which the reference's translation
// units include ahead of them (include/NEST/NEST.hh:105-111): the waveform constants ...
#ifndef FIELD_MIN  // include/NEST/NEST.hh:81-100
#define HIGH_E_NR 330.
#define W_DEFAULT 13.4
#define W_SCINT 8.5e-3
#define NEST_AVO 6.0221409e+23
#define ATOM_NUM 54.
#define PHE_MIN 1e-6
#define FIELD_MIN 1.
#define DENSITY 2.90
#define EPS_GAS 1.00126
#define EPS_LIQ 1.85
#endif
#ifndef SAMPLE_SIZE
#define SAMPLE_SIZE 10  // ns
#define PULSE_WIDTH 10  // ns
#define PULSEHEIGHT 0.005
#define SPIKES_MAXM 120
#define PHE_MAX 140
#endif

// ... and the RandomGen singleton (include/NEST/RandomGen.hh:16-36).  The event loop never draws from
// it -- the GPU chain owns its counter-based streams -- but user hooks such as OptTrans /
// SinglePEWaveForm call it, so the interface is kept (host SplitMix64; same distributions as
// src/RandomGen.cpp:39-85).
class RandomGen {
 public:
  static RandomGen* rndm() {
    static RandomGen r;
    return &r;
  }
  void SetSeed(uint64_t s) { state_ = s; }
  void LockSeed() {}
  void UnlockSeed() {}
  double rand_uniform() {
    uint64_t z = (state_ += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (double(z >> 11) + 0.5) * (1.0 / 9007199254740992.0);
  }
  double rand_gauss(double mean, double sigma, bool zero_min = false) {
    double u = rand_uniform(), v = rand_uniform();
    double d = mean + sigma * std::sqrt(2.) * std::sqrt(-std::log(u)) * std::cos(2. * M_PI * v);
    return zero_min ? std::max(d, 0.) : d;
  }
  double rand_zero_trunc_gauss(double mean, double sigma) {
    double r = rand_gauss(mean, sigma);
    while (r <= 0.) r = rand_gauss(mean, sigma);
    return r;
  }
  int SelectRanXeAtom() {  // src/RandomGen.cpp:159-182: natural abundances, cumulative per cent
    const double iso = rand_uniform() * 100.;
    static const double cum[8] = {0.090, 0.180, 2.100, 28.54, 32.62, 53.80, 80.69, 91.13};
    static const int A[9] = {124, 126, 128, 129, 130, 131, 132, 134, 136};
    for (int k = 0; k < 8; ++k)
      if (iso <= cum[k]) return A[k];
    return A[8];
  }
  double rand_exponential(double half_life, double t_min = -1, double t_max = -1) {
    double r = rand_uniform(), r_min = 0, r_max = 1;
    if (t_min >= 0) r_min = 1 - std::exp(-t_min * std::log(2.) / half_life);
    if (t_max >= 0) r_max = 1 - std::exp(-t_max * std::log(2.) / half_life);
    r = (r_max - r_min) * r + r_min;
    return -std::log(1 - r) * half_life / std::log(2.);
  }

 private:
  RandomGen() = default;
  uint64_t state_ = 0x853C49E6748FEA9Bull;
};

#define NB_DET_SCALAR(T, name, dflt)       \
 protected:                                \
  T name = dflt;                           \
                                           \
 public:                                   \
  T get_##name() const { return name; }    \
  void set_##name(T v) { name = v; }

class VDetector {
 public:
  VDetector() = default;
  virtual ~VDetector() = default;
  virtual void Initialization() {}
  std::string getName() const { return name; }

  typedef enum { fold = 0, unfold = 1 } LCE;
  virtual double FitS1(double, double, double, LCE) { return 1.; }
  virtual double FitEF(double, double, double) { return 730.; }
  virtual std::vector<double> FitDirEF(double, double, double) { return {0, 0, 1}; }
  virtual double FitS2(double, double, LCE) { return 1.; }
  virtual std::vector<double> FitTBA(double, double, double) { return {0.5, 0.5}; }
  // photon-timing hooks (VDetector.hh:150-154): declared so that detector headers overriding them
  // compile unchanged; the waveform modes that call them are outside this path
  virtual double OptTrans(double, double, double) { return 0.; }
  virtual std::vector<double> SinglePEWaveForm(double, double) { return {0.}; }

  const double* get_noiseBaseline() const { return noiseBaseline; }
  const double* get_noiseLinear() const { return noiseLinear; }
  const double* get_noiseQuadratic() const { return noiseQuadratic; }
  void set_noiseBaseline(double a, double b, double c, double d) {
    noiseBaseline[0] = a; noiseBaseline[1] = b; noiseBaseline[2] = c; noiseBaseline[3] = d;
  }
  void set_noiseLinear(double a, double b) { noiseLinear[0] = a; noiseLinear[1] = b; }
  void set_noiseQuadratic(double a, double b) { noiseQuadratic[0] = a; noiseQuadratic[1] = b; }

  // defaults: VDetector.hh:157-224
  NB_DET_SCALAR(double, g1, 0.0760)
  NB_DET_SCALAR(double, sPEres, 0.58)
  NB_DET_SCALAR(double, sPEthr, 0.35)
  NB_DET_SCALAR(double, sPEeff, 1.00)
  NB_DET_SCALAR(double, P_dphe, 0.2)
  NB_DET_SCALAR(double, coinWind, 100)
  NB_DET_SCALAR(int, coinLevel, 2)
  NB_DET_SCALAR(int, numPMTs, 89)
  NB_DET_SCALAR(bool, OldW13eV, true)
  NB_DET_SCALAR(double, g1_gas, 0.06)
  NB_DET_SCALAR(double, s2Fano, 3.61)
  NB_DET_SCALAR(double, s2_thr, 300.)
  NB_DET_SCALAR(double, E_gas, 12.)
  NB_DET_SCALAR(double, eLife_us, 2200.)
  NB_DET_SCALAR(bool, inGas, false)
  NB_DET_SCALAR(double, T_Kelvin, 177.)
  NB_DET_SCALAR(double, p_bar, 2.14)
  NB_DET_SCALAR(double, dtCntr, 40.)
  NB_DET_SCALAR(double, dt_min, 20.)
  NB_DET_SCALAR(double, dt_max, 60.)
  NB_DET_SCALAR(double, radius, 50.)
  NB_DET_SCALAR(double, radmax, 50.)
  NB_DET_SCALAR(double, TopDrift, 150.)
  NB_DET_SCALAR(double, anode, 152.5)
  NB_DET_SCALAR(double, gate, 147.5)
  NB_DET_SCALAR(double, cathode, 1.00)
  NB_DET_SCALAR(double, PosResFlat, 3.00)
  NB_DET_SCALAR(double, PosResBase, 70.8364)
  NB_DET_SCALAR(double, molarMass, 131.293)

 protected:
  std::string name = "VDetector";
  double noiseBaseline[4] = {0., 0., 0., 0.};
  double noiseLinear[2] = {3e-2, 3e-2};
  double noiseQuadratic[2] = {0., 0.};
};

namespace NEST {

typedef enum {
  NR = 0, WIMP = 1, B8 = 2, DD = 3, AmBe = 4, Cf = 5, ion = 6, gammaRay = 7, beta = 8, CH3T = 9, C14 = 10,
  Kr83m = 11, ppSolar = 12, atmNu = 13, fullGamma = 14, fullGamma_PE = 15, fullGamma_Compton_PP = 16,
  NoneType = 17
} INTERACTION_TYPE;

enum class S1CalculationMode { Full, Parametric, Hybrid, Waveform };
enum class S2CalculationMode { Full, Waveform, WaveformWithEtrain };

struct YieldResult {
  double PhotonYield, ElectronYield, ExcitonRatio, Lindhard, ElectricField, DeltaT_Scint;
};

struct QuantaResult {  // NEST.hh:171-178
  int photons, electrons, ions, excitons;
  double recombProb, Variance;
};
typedef std::vector<double> photonstream;
struct NESTresult {  // NEST.hh:182-186
  YieldResult yields;
  QuantaResult quanta;
  photonstream photon_times;
};

// analysis.hh's compile-time globals as a runtime object (include/NEST/analysis.hh:7-77)
struct Analysis {
  int verbosity = 1;
  short PrintSubThr = 1;
  bool MCtruthE = false, MCtruthPos = false;
  S1CalculationMode s1CalculationMode = S1CalculationMode::Hybrid;
  S2CalculationMode s2CalculationMode = S2CalculationMode::Full;
  int usePD = 2, useS2 = 0;
  double minS1 = 1.5, maxS1 = 99.5;
  int numBins = 98;
  double minS2 = 42., maxS2 = 1e4, logMax = 3.6, logMin = 0.6;
  int logBins = 30;
  double z_step = 0.1, E_step = 5.0;
  static Analysis G2();  // analysis_G2.hh values
  nb200_analysis flat() const;
};

const std::vector<double> default_NRYieldsParam = {11., 1.1, 0.0480, -0.0533, 12.6, 0.3, 2., 0.3, 2., 0.5, 1., 1.};
const std::vector<double> default_NRERWidthsParam = {0.4, 0.4, 0.04, 0.5, 0.19, 2.25, -0.001,
                                                     0.046452, 0.45, 0.205, -0.2, 0., 0.};
const std::vector<double> default_ERYieldsParam = {-1., -1., -1., -1., -1., -1., -1., -1., -1., -1.};

// one GPU context per process (the reference's RandomGen / NESTcalc are process-wide state too)
class Device {
 public:
  static nb200_ctx* get();        // throws std::runtime_error without a CUDA device: no CPU fallback
  static void check(int rc);      // nonzero -> throw std::runtime_error(nb200_last_error)
};

class NESTcalc {
 public:
  struct Wvalue {
    double Wq_eV, alpha;
  };
  explicit NESTcalc(VDetector* detector);
  NESTcalc(const NESTcalc&) = delete;
  NESTcalc& operator=(const NESTcalc&) = delete;

  double SetDensity(double Kelvin, double bara);                                // NEST.cpp:2270
  double SetDriftVelocity(double Kelvin, double Density, double eField, double Bar = 1.5);  // :2352
  // the static forms (NEST.hh:373-402), liquid phase; gas-phase conditions are outside this path (std::runtime_error)
  static double GetDriftVelocity(double Kelvin, double Density, double eField, bool inGas, double Bar);
  static double GetDriftVelocity_Liquid(double Kelvin, double eField, double Density, double Bar = 1.5, short SD = -1);
  static double GetDensity(double Kelvin, double bara, bool& inGas, uint64_t evtNum = 0, double molarMass = 131.293);
  double NexONi(double energy, double density);                                 // :2850
  void SetDetector(VDetector* detector) { fdetector = detector; flattened = false; }
  std::vector<double> SetDriftVelocity_NonUniform(double rho, double zStep, double Kelvin, double Bar, double dx,
                                                  double dy);                   // :2664 (liquid; host, through FitEF)
  std::vector<double> CalculateG2(int verbosity = 1);                           // :2065 (prints the banner)
  static Wvalue WorkFunction(double rho, double MolarMass, bool OldW13eV = true);  // :2829

  // NEST.cpp:1211-1252 on the device; the vector form evaluates n (energy, field) pairs in one launch
  YieldResult GetYields(INTERACTION_TYPE species, double energy, double density, double dfield, double massNum,
                        double atomNum, const std::vector<double>& NRYieldsParam = default_NRYieldsParam,
                        const std::vector<double>& ERYieldsParam = default_ERYieldsParam);
  std::vector<YieldResult> GetYields(INTERACTION_TYPE species, const std::vector<double>& energy, double density,
                                     const std::vector<double>& dfield, double massNum, double atomNum,
                                     const std::vector<double>& NRYieldsParam = default_NRYieldsParam,
                                     const std::vector<double>& ERYieldsParam = default_ERYieldsParam);

  // The model methods GetYields dispatches to (NEST.hh:256-303), public and virtual in the reference and called
  // directly by nestpy and execNEST.cpp:1032-1048: each is nb200_yields with the matching switch; scalar as in the
  // reference, and a vector form (one launch for n energies / fields).  GetYieldKr83m draws its decay-time
  // separation (the reference: RandomGen) from the Philox stream of this object's next call number.
  YieldResult GetYieldGamma(double energy, double density, double dfield, double multFact = 1.);
  YieldResult GetYieldERWeighted(double energy, double density, double dfield,
                                 const std::vector<double>& ERYieldsParam = default_ERYieldsParam);
  YieldResult GetYieldNR(double energy, double density, double dfield, double massNum,
                         const std::vector<double>& NRYieldsParam = default_NRYieldsParam);
  YieldResult GetYieldIon(double energy, double density, double dfield, double massNum, double atomNum,
                          const std::vector<double>& NRYieldsParam = default_NRYieldsParam);
  YieldResult GetYieldKr83m(double energy, double density, double dfield, double maxTimeSeparation = 1000.,
                            double minTimeSeparation = 300.);
  YieldResult GetYieldBetaGR(double energy, double density, double dfield,
                             const std::vector<double>& ERYieldsParam = default_ERYieldsParam, double multFact = 1.);
  std::vector<YieldResult> GetYieldGamma(const std::vector<double>& energy, double density,
                                         const std::vector<double>& dfield, double multFact = 1.);
  std::vector<YieldResult> GetYieldERWeighted(const std::vector<double>& energy, double density,
                                              const std::vector<double>& dfield,
                                              const std::vector<double>& ERYieldsParam = default_ERYieldsParam);
  std::vector<YieldResult> GetYieldNR(const std::vector<double>& energy, double density, const std::vector<double>& dfield,
                                      double massNum, const std::vector<double>& NRYieldsParam = default_NRYieldsParam);
  std::vector<YieldResult> GetYieldIon(const std::vector<double>& energy, double density, const std::vector<double>& dfield,
                                       double massNum, double atomNum,
                                       const std::vector<double>& NRYieldsParam = default_NRYieldsParam);
  std::vector<YieldResult> GetYieldBetaGR(const std::vector<double>& energy, double density,
                                          const std::vector<double>& dfield,
                                          const std::vector<double>& ERYieldsParam = default_ERYieldsParam,
                                          double multFact = 1.);
  // NEST.cpp:1254-1271: the clamps every GetYield* applies to its result (in place, and returned)
  YieldResult YieldResultValidity(YieldResult& res, const double energy, const double Wq_eV);
  // NEST.cpp:164-246 on the device (nb200_widths): the width model GetQuanta uses; vector forms = one launch
  double RecombOmegaNR(double elecFrac, const std::vector<double>& NRERWidthsParam = default_NRERWidthsParam);
  double RecombOmegaER(double efield, double elecFrac, double numQuanta,
                       const std::vector<double>& NRERWidthsParam = default_NRERWidthsParam);
  double FanoER(double density, double Nq_mean, double efield,
                const std::vector<double>& NRERWidthsParam = default_NRERWidthsParam);
  // rows of {omega_NR, omega_ER, Fano_ER} for n (efield, elecFrac, numQuanta) records
  std::vector<std::vector<double>> Widths(const std::vector<double>& efield, const std::vector<double>& elecFrac,
                                          const std::vector<double>& numQuanta, double density,
                                          const std::vector<double>& NRERWidthsParam = default_NRERWidthsParam);
  // NEST.cpp:2243-2268 (NEST.hh:345) on the device: member scratch of two values, valid until the next call
  const std::vector<double>& GetSpike(int Nph, double dx, double dy, double dz, double driftSpeed, double dS_mid,
                                      const std::vector<double>& origScint);
  // NEST.cpp:2699-2736 (NEST.hh:404) on the device; the vector form smears n positions in one launch (rows {x, y})
  std::vector<double> xyResolution(double xPos_mm, double yPos_mm, double A_top);
  std::vector<std::vector<double>> xyResolution(const std::vector<double>& xPos_mm, const std::vector<double>& yPos_mm,
                                                const std::vector<double>& A_top);

  // NEST.cpp:248-432 on the device.  The reference draws from its process-wide RandomGen; here call number k of this
  // object draws from the Philox streams of event k under the seed given to SetSeed (default 0), so a sequence of
  // calls is reproducible and the vector form equals the same calls made one by one.
  QuantaResult GetQuanta(const YieldResult& yields, double density,
                         const std::vector<double>& NRERWidthsParam = default_NRERWidthsParam, double SkewnessER = -999.);
  std::vector<QuantaResult> GetQuanta(const std::vector<YieldResult>& yields, double density,
                                      const std::vector<double>& NRERWidthsParam = default_NRERWidthsParam,
                                      double SkewnessER = -999.);
  // NEST.cpp:17-40.  Photon times (do_times = true, the reference's default) are outside the GPU path: they are refused
  // with std::runtime_error rather than returned as zeros; with do_times = false photon_times is photons x 0.0 as in
  // the reference.
  NESTresult FullCalculation(INTERACTION_TYPE species, double energy, double density, double dfield, double A, double Z,
                             const std::vector<double>& NRYieldsParam = default_NRYieldsParam,
                             const std::vector<double>& NRERWidthsParam = default_NRERWidthsParam,
                             const std::vector<double>& ERYieldsParam = default_ERYieldsParam, bool do_times = true);
  // NEST.cpp:1288-1722 on the device (NEST.hh:335-340: the result is member scratch, valid until the next call): prelude,
  // the chain's hit kernel, epilogue.  Full, Parametric and Hybrid modes; Waveform is refused.  driftSpeed / dS_mid
  // replace the detector's own drift speeds for this call, as in the reference.
  const std::vector<double>& GetS1(const QuantaResult& quanta, double truthPosX, double truthPosY, double truthPosZ,
                                   double smearPosX, double smearPosY, double smearPosZ, double driftSpeed, double dS_mid,
                                   INTERACTION_TYPE species, uint64_t evtNum, double dfield, double energy,
                                   S1CalculationMode mode, int outputTiming, std::vector<int64_t>& wf_time,
                                   std::vector<double>& wf_amp);
  // n records at once: rows of nine (scintillation[0..8]); record i uses event number firstEvt + i
  std::vector<std::vector<double>> GetS1(const std::vector<int>& photons, const std::vector<std::vector<double>>& truthPos,
                                         const std::vector<std::vector<double>>& smearPos, double driftSpeed, double dS_mid,
                                         uint64_t firstEvt, double dfield, const std::vector<double>& energy,
                                         S1CalculationMode mode);
  // NEST.cpp:1724-2063, S2CalculationMode::Full, on the device (NEST.hh:351-356: the result is member scratch, valid
  // until the next call).  evtNum selects the Philox streams (seed from SetSeed); the Waveform modes are refused.
  const std::vector<double>& GetS2(int Ne, double truthPosX, double truthPosY, double truthPosZ, double smearPosX,
                                   double smearPosY, double smearPosZ, double dt, double driftSpeed, uint64_t evtNum,
                                   double dfield, S2CalculationMode mode, int outputTiming, std::vector<int64_t>& wf_time,
                                   std::vector<double>& wf_amp, const std::vector<double>& g2_params);
  // the same for n records at once: rows of nine (ionization[0..8]); record i uses event number firstEvt + i
  std::vector<std::vector<double>> GetS2(const std::vector<int>& Ne, const std::vector<std::vector<double>>& truthPos,
                                         const std::vector<std::vector<double>>& smearPos, const std::vector<double>& dt,
                                         uint64_t firstEvt, double dfield, const std::vector<double>& g2_params);
  void SetSeed(uint64_t seed) { call_seed = seed; call_index = 0; }  // RandomGen::rndm()->SetSeed

  const nb200_detector& flat();   // the detector flattened for the device (maps tabulated on first use)
  const nb200_run_consts& prepared(double inField, double z_step = 0.);
  VDetector* detector() const { return fdetector; }

 protected:
  VDetector* fdetector;
  uint64_t call_seed = 0, call_index = 0;  // Philox seed and next event number of the per-call stage API

 private:
  nb200_detector fdet{};
  std::vector<double> lut[4];
  bool flattened = false;
  nb200_run_consts rc{};
  double rc_field = -12345.;
  std::vector<double> ionization = std::vector<double>(9, 0.), scintillation = std::vector<double>(9, 0.);
  std::vector<double> newSpike = std::vector<double>(2, 0.);
  std::vector<YieldResult> yields_by(int er_model, int species, const std::vector<double>& energy, double density,
                                     const std::vector<double>& dfield, double massNum, double atomNum,
                                     const std::vector<double>& nr, const std::vector<double>& er, double multFact);
};

// ---- liquid argon: include/NEST/LArNEST.hh:19-42, LArParameters.hh:27-34 ----
enum class LArInteraction { NR = 0, ER = 1, Alpha = 2, dEdx = 3, LeptonLET = 4, LET = 5 };
struct LArYieldResult {
  double TotalYield, QuantaYield, LightYield, Nph, Ne, Nex, Nion, Lindhard, ElectricField;
};
struct LArYieldFluctuationResult {
  double NphFluctuation, NeFluctuation, NexFluctuation, NionFluctuation;
};
struct LArNESTResult {
  LArYieldResult yields;
  LArYieldFluctuationResult fluctuations;
  photonstream photon_times;
};

// LArNEST::FullCalculation / GetYields (src/LArNEST.cpp:17-33, 35-212) for the NR, ER and Alpha models with the
// parameters of LArParameters.hh, on the device (nb200_lar -> k_lar).  The dE/dx and LET models and the parameter
// setters of the reference class are not part of this path (std::runtime_error / absent).
class LArNEST : public NESTcalc {
 public:
  explicit LArNEST(VDetector* detector) : NESTcalc(detector) {}
  LArNESTResult FullCalculation(LArInteraction species, double energy, double dx, double efield, double density,
                                bool do_times);
  std::vector<LArNESTResult> FullCalculation(LArInteraction species, const std::vector<double>& energy,
                                             const std::vector<double>& efield, double density);
  LArYieldResult GetYields(LArInteraction species, double energy, double dx, double efield, double density);
};

}  // namespace NEST

// include/NEST/TestSpectra.hh:38-73: the energy samplers, static as in the reference.  Each scalar call hands out
// the next energy of a block the device drew in one launch (nb200_spectrum; 16384 at a time, redrawn when the
// arguments change), so that reference-style loops `for (...) keV = TestSpectra::CH3T_spectrum(0, 18.6)` do not pay a
// launch per energy; the *_vec forms return n energies of one launch.  TestSpectra::SetSeed replaces
// RandomGen::rndm()->SetSeed for these streams.  Gamma_spectrum (line tables of GammaHandler) and ZeplinBackground
// are outside the hot path.
class TestSpectra {
 public:
  TestSpectra() = default;
  struct WIMP_spectrum_prep {
    double base[100] = {1.};
    double exponent[100] = {0.};
    double integral = 0.;
    double xMax = 0.;
    double divisor = 1.;
  };
  WIMP_spectrum_prep wimp_spectrum_prep;

  static double CH3T_spectrum(double emin, double emax);
  static double C14_spectrum(double emin, double emax);
  static double B8_spectrum(double emin, double emax);
  static double AmBe_spectrum(double emin, double emax);
  static double PowLawFit_spectrum(double emin, double emax, double exp);
  static double Cf_spectrum(double emin, double emax);
  static double DD_spectrum(double xMin, double xMax, double expFall, double peakFrac, double peakMu, double peakSig,
                            double peakSkew);
  static double ppSolar_spectrum(double emin, double emax);
  static double atmNu_spectrum(double emin, double emax);
  static double WIMP_dRate(double ER, double mWimp, double day);  // random Xe isotope per call, as the reference
  static WIMP_spectrum_prep WIMP_prep_spectrum(double mass, double eStep, double day);
  static double WIMP_spectrum(WIMP_spectrum_prep wprep, double mass, double day);
  // n energies in one launch: kind = NB200_SRC_*, par = the source's shape parameters (nb200_source::par)
  static std::vector<double> spectrum_vec(int kind, size_t n, double emin, double emax,
                                          const std::vector<double>& par = {}, const WIMP_spectrum_prep* wprep = nullptr,
                                          double mass = 0., double day = 0.);
  static void SetSeed(uint64_t seed);
};

// include/NEST/execNEST.hh:10-65 (vectors filled by store_signals, :259-326)
class NESTObservableArray {
 public:
  std::vector<double> energy_kev, x_mm, y_mm, z_mm;
  std::vector<int> n_electrons, n_photons;
  std::vector<int> s1_nhits, s1_nhits_dpe, s1_nhits_thr;
  std::vector<double> s1r_phe, s1c_phe, s1r_phd, s1c_phd, s1r_spike, s1c_spike;
  std::vector<int> s2_Nee, s2_Nph, s2_nhits, s2_nhits_dpe;
  std::vector<double> s2r_phe, s2c_phe, s2r_phd, s2c_phd;
};

// execNEST.hh:72-87, unchanged signatures (the Waveform modes and calculate_times=true are refused)
NESTObservableArray runNESTvec(VDetector* detector, NEST::INTERACTION_TYPE particleType, std::vector<double> eList,
                               std::vector<std::vector<double>> pos3dxyz, double inField = -1.0, int seed = 0,
                               std::vector<double> ERYieldsParam = NEST::default_ERYieldsParam,
                               std::vector<double> NRYieldsParam = NEST::default_NRYieldsParam,
                               std::vector<double> NRERWidthsParam = NEST::default_NRERWidthsParam,
                               NEST::S1CalculationMode s1mode = NEST::S1CalculationMode::Hybrid,
                               NEST::S2CalculationMode s2mode = NEST::S2CalculationMode::Full,
                               bool calculate_times = false);

int execNEST(VDetector* detector, double numEvts, const std::string& type, double eMin, double eMax, double inField,
             std::string position, const std::string& posiMuon, double fPos, int seed, bool no_seed,
             double dayNumber);

// state the reference keeps in file-scope globals of execNEST.cpp (:30-38, :185-249)
namespace nest_b200_cli {
extern NEST::Analysis analysis;
extern std::vector<double> NRYieldsParam, NRERWidthsParam, ERYieldsParam;
extern double multFact;
int main_like(VDetector* detector, int argc, char** argv);  // execNEST.cpp:40-256
}  // namespace nest_b200_cli
