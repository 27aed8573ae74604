/*
 * nest_b200.h -- C-ABI of the B200-native NEST detector-response chain.
 *
 * The reference (NESTCollaboration/nest) has no FFI layer: its operator API is the
 * C++ class surface NESTcalc/VDetector plus the execNEST()/runNESTvec() drivers.
 * Every entry point below names the reference interface it replaces (file:line,
 * relative to the reference root).  Plain pointers and sizes only; no C++ or
 * torch types cross this boundary.  INTEGRATION.md shows the reference-side
 * binding (a header-only flattener VDetector* -> nb200_detector and the
 * ctypes stub used by the tests).
 *
 * Conventions
 *   - every function returns 0 on success, a negative NB200_E* code otherwise
 *     (the reference's "cerr << msg; return 1" / throw std::runtime_error,
 *     src/execNEST.cpp:1410-1413); nb200_last_error() holds the message.
 *   - one context per device and per host thread; calls are stream-ordered on
 *     the context's stream (nb200_set_stream) and synchronous at return unless
 *     the function name ends in _async.
 *   - results for event id e depend only on (seed, e, parameters): they are
 *     independent of batch split, launch geometry and GPU count.
 *   - there is no CPU fallback: without a CUDA device every compute entry
 *     point fails with NB200_ENODEV.
 */
#ifndef NEST_B200_H
#define NEST_B200_H 1

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NB200_ABI_VERSION 1

/* error codes */
#define NB200_OK 0
#define NB200_EINVAL (-1)   /* bad argument (reference: cerr + return 1)        */
#define NB200_ENODEV (-2)   /* no CUDA device / wrong architecture              */
#define NB200_ECUDA (-3)    /* CUDA runtime error, see nb200_last_error         */
#define NB200_EPHYSICS (-4) /* reference would throw std::runtime_error         */
#define NB200_ENOMEM (-5)
#define NB200_EFIT (-6)     /* a histogram fit found no minimum (too few filled bins) */

/* INTERACTION_TYPE, include/NEST/NEST.hh:133-156 (same numeric values) */
enum {
  NB200_NR = 0, NB200_WIMP = 1, NB200_B8 = 2, NB200_DD = 3, NB200_AmBe = 4,
  NB200_Cf = 5, NB200_ion = 6, NB200_gammaRay = 7, NB200_beta = 8,
  NB200_CH3T = 9, NB200_C14 = 10, NB200_Kr83m = 11, NB200_ppSolar = 12,
  NB200_atmNu = 13, NB200_NoneType = 17
};

/* S1CalculationMode / S2CalculationMode, NEST.hh:158-160.  The Waveform modes
 * are out of scope and rejected with NB200_EINVAL. */
enum { NB200_S1_FULL = 0, NB200_S1_PARAMETRIC = 1, NB200_S1_HYBRID = 2 };
enum { NB200_S2_FULL = 0 };

/* ---- position maps: the four VDetector virtuals on the path ------------------
 * VDetector.hh:131-154 (FitS1, FitEF, FitS2, FitTBA).  Virtual C++ functions
 * cannot run on the device, so a detector is flattened to one of:
 *   CONST      value = c[0]                                (VDetector defaults)
 *   LUX_RUN03  the closed forms of include/Detectors/LUX_Run03.hh:90-167,
 *              coefficients in c[] (layout documented in nest_b200/csrc/maps.cuh)
 *   LUT_RZ     bilinear table over (r [mm], z [mm]) tabulated from the
 *              virtual at flatten time; lut is a HOST pointer to n0*n1 doubles
 *              (row = r index), axes r in [c[0],c[1]], z in [c[2],c[3]]
 */
enum { NB200_MAP_CONST = 0, NB200_MAP_LUX_RUN03 = 1, NB200_MAP_LUT_RZ = 2 };

typedef struct nb200_map {
  int32_t kind;
  int32_t n0, n1;
  int32_t _pad;
  double c[16];
  const double* lut;
} nb200_map;

/* ---- VDetector flattened: include/Detectors/VDetector.hh:157-224 -------------
 * Same member names and units as the reference class. */
typedef struct nb200_detector {
  double g1, sPEres, sPEthr, sPEeff;
  double noiseBaseline[4];
  double P_dphe;
  double coinWind;
  int32_t coinLevel, numPMTs;
  int32_t OldW13eV, inGas;
  double noiseLinear[2], noiseQuadratic[2];
  double g1_gas, s2Fano, s2_thr, E_gas, eLife_us;
  double T_Kelvin, p_bar;
  double dtCntr, dt_min, dt_max;
  double radius, radmax, TopDrift, anode, cathode, gate;
  double PosResFlat, PosResBase;
  double molarMass;
  nb200_map FitS1, FitS2, FitEF, FitTBA; /* FitTBA: value = BotTotRat[1] */
} nb200_detector;

/* Built-in flattenings of shipped detectors (closed-form maps).
 * name: "LUX_Run03" (LUX_Run03.hh:30-88) or "VDetector" (VDetector.hh:157-224
 * defaults; also what LArDetector really is). */
int nb200_detector_builtin(const char* name, nb200_detector* out);

/* ---- model parameter vectors ---------------------------------------------------
 * NRYieldsParam[12], ERYieldsParam[10], NRERWidthsParam[17]: NEST.hh:120-124,
 * src/execNEST.cpp:185-249.  n_widths is the reference vector's size() (13..17:
 * RecombOmegaER reads [13..16] only when present, NEST.cpp:206-212).
 * SkewnessER: GetQuanta's 4th argument (NEST.cpp:248-250; -999 = LUX model).
 * er_model: which ER yield the "ER" CLI type uses (execNEST.cpp:1032-1048):
 *   0 GetYieldBetaGR(multFact), 1 GetYieldGamma(-multFact), 2 GetYieldERWeighted
 *   and -1 = plain GetYields(species) dispatch (NEST.cpp:1211-1252). */
typedef struct nb200_model {
  double NRYieldsParam[12];
  double ERYieldsParam[10];
  double NRERWidthsParam[17];
  int32_t n_widths;
  int32_t er_model;
  double SkewnessER;
  double multFact;
  double massNum;  /* 0 => random Xe isotope per NR event (execNEST.cpp:670) */
  double atomNum;
} nb200_model;

/* which==0: default_* vectors of NEST.hh:120-124 (what runNESTvec uses);
 * which==1: the hard-coded CLI vectors of execNEST.cpp:188-248, SkewnessER=0. */
int nb200_model_default(int which, nb200_model* out);

/* ---- analysis.hh compile-time globals as a runtime struct ----------------------
 * include/NEST/analysis.hh:7-77 / analysis_G2.hh. */
typedef struct nb200_analysis {
  int32_t usePD, useS2;
  int32_t numBins, logBins;
  double minS1, maxS1, minS2, maxS2, logMin, logMax;
  int32_t MCtruthE, MCtruthPos;
  int32_t s1mode, s2mode;
  double z_step;
} nb200_analysis;

/* which==0: analysis.hh (LUX band); which==1: analysis_G2.hh (as shipped). */
int nb200_analysis_default(int which, nb200_analysis* out);

/* ---- energy source: TestSpectra / execNEST.cpp:681-790 -------------------------
 * kind FLAT/GAUSS/EXP are the "default:" branch (execNEST.cpp:763-781), MONO is
 * eMin==eMax (:681-683); CH3T TestSpectra.cpp:31-58, DD :192-211 with the five
 * shape parameters of execNEST.cpp:705 in par[0..4], B8 :110-128, AMBE = PowLawFit
 * :151-167 with the exponent of execNEST.cpp:698 in par[0].
 * LIST: energies (and optionally positions) supplied by the caller, the
 * runNESTvec form (execNEST.cpp:329-409). */
enum {
  NB200_SRC_MONO = 0, NB200_SRC_FLAT = 1, NB200_SRC_GAUSS = 2, NB200_SRC_EXP = 3,
  NB200_SRC_CH3T = 4, NB200_SRC_DD = 5, NB200_SRC_B8 = 6, NB200_SRC_LIST = 7,
  NB200_SRC_WIMP = 8, NB200_SRC_AMBE = 9,
  /* C14_spectrum :60-108, Cf_spectrum :169-190, ppSolar_spectrum :213-227, atmNu_spectrum :229-245 */
  NB200_SRC_C14 = 10, NB200_SRC_CF = 11, NB200_SRC_PPSOLAR = 12, NB200_SRC_ATMNU = 13
};

typedef struct nb200_source {
  int32_t kind;
  int32_t species;   /* INTERACTION_TYPE */
  double eMin, eMax;
  double par[8];
  /* position: fixed_pos==0 => random (z,r,phi) with drift-time window retry
   * (execNEST.cpp:806-814, 962-967); else xyz[] used for every event (a -1 z or -999 x,y
   * on the reference CLI = the random-z / random-xy variants 2 and 3, :845-853). */
  int32_t fixed_pos; /* 0 random; 1 fixed xyz; 2 fixed xy, random z; 3 random xy, fixed z */
  int32_t vec_mode;  /* 1 = runNESTvec semantics (execNEST.cpp:357-406): positions as given,
                      * no xy smearing, no cathode cut, no drift-window retry, vD from the
                      * per-event field, yields through FullCalculation (NEST.cpp:22-42) */
  double xyz[3];
  double inField;    /* V/cm; -1 => FitEF(x,y,z) per event (execNEST.cpp:857-861) */
  /* LIST only: n_events entries each, host or device memory per list_mem_kind
   * (NB200_MEM_*); x/y/z may be NULL (then fixed_pos/xyz apply) */
  int32_t list_mem_kind;
  int32_t _pad;
  const double* E;
  const double* x;
  const double* y;
  const double* z;
  /* WIMP only: table from nb200_wimp_prep (TestSpectra.hh:43-49) */
  const double* wimp_base;
  const double* wimp_exponent;
  double wimp_xMax, wimp_divisor, wimp_mass, wimp_day;
} nb200_source;

/* ---- per-run constants computed on the host once -------------------------------
 * SetDensity (NEST.cpp:2270-2350), WorkFunction (:2829-2848), SetDriftVelocity
 * (:2352-2521), CalculateG2 (:2065-2235), FindNewMean (RandomGen.cpp:63-70). */
typedef struct nb200_run_consts {
  double rho, Wq_eV, alpha;
  double vD, vD_middle;
  double g2_params[5]; /* elYield, ExtEff, SE, g2, gasGap */
  double NewMean;
  double field;        /* uniform drift field used, V/cm */
} nb200_run_consts;

/* ---- outputs --------------------------------------------------------------------
 * Structure-of-arrays record per event.  Any pointer may be NULL (column not
 * written).  Column meaning = the reference vectors:
 *   GetS1 scintillation[0..8] NEST.cpp:1632-1645, GetS2 ionization[0..8]
 *   :2019-2060 (sign = below-threshold flag, :1689-1716, :2051-2057),
 *   quanta NEST.hh:171-178, truth as NESTObservableArray execNEST.hh:30-44,
 *   signal1/signal2/signalE execNEST.cpp:1121-1158,1245-1250. */
enum { NB200_MEM_HOST = 0, NB200_MEM_DEVICE = 1 };
/* nb200_soa_out.flags: NB200_SOA_F32 = the packed record -- every `double*` column below is an array of FLOAT
 * (4 bytes per entry; the int32 columns are unchanged): half the bytes over PCIe for callers that histogram or plot
 * the signals (7 significant digits; the reference CLI prints %.6f).  Default 0: doubles, as the reference's own
 * vectors. */
enum { NB200_SOA_F32 = 1 };

typedef struct nb200_soa_out {
  int32_t mem_kind;
  int32_t flags;
  double *energy_kev, *x_mm, *y_mm, *z_mm, *field, *driftTime;
  double *smear_x, *smear_y;
  int32_t *photons, *electrons, *ions, *excitons;
  double *s1[9];
  double *s2[9];
  double *signal1, *signal2, *signalE; /* band inputs; keV_recon */
  double *Nph_mean, *Ne_mean;          /* yields.PhotonYield / ElectronYield */
  double *deltaT_ns;                   /* yields.DeltaT_Scint: Kr83m decay separation (-999 otherwise) */
  double *keVee;                       /* this event's term of the keVee sum, execNEST.cpp:1239-1240 */
} nb200_soa_out;

/* Band accumulator = what GetBand (execNEST.cpp:1508-1621) reduces, kept as exact
 * integers so that any GPU count / batch split gives bit-identical results:
 * per S1 bin: accepted count, rejected count, and fixed-point sums of S1, y,
 * y^2, y^3 (y = log10(S2/S1) or the 0 the reference pushes when y is outside
 * (logMin,logMax), :1551-1555); plus the numBins x logBins 2-D histogram of y.
 * All arrays are caller-owned (HOST), sized by nb200_band_hist_bytes(). */
/* sumS1 (the band's X: S1, or S2 for useS2 == 2) is in units of 1 / nb200_band_scale_x(ana): 2^-20 (NB200_FX_S1)
 * while the analysis keeps X below 1024, else the largest power of two with X * scale < 2^30.  A bin's 64-bit
 * sums hold at least 2^63 / (6.4 * 2^30) = 1.3e9 accepted events for |log10| < 6.4 (2^33 for sumS1);
 * nb200_band_finalize refuses (NB200_EINVAL) an accumulator in which a bin's count exceeds what its sums can
 * carry.  Split longer runs over several accumulators. */
#define NB200_FX_S1 1048576.0        /* 2^20  */
#define NB200_FX_Y 1073741824.0      /* 2^30  */
#define NB200_FX_Y2 67108864.0       /* 2^26  */
#define NB200_FX_Y3 4194304.0        /* 2^22  */

typedef struct nb200_band_hist {
  int32_t numBins, logBins;
  uint64_t n_events;    /* events simulated into this accumulator */
  uint64_t* accepted;   /* [numBins] */
  uint64_t* rejected;   /* [numBins] */
  int64_t* sumS1;       /* [numBins] units 2^-20 */
  int64_t* sumY;        /* [numBins] units 2^-30 */
  int64_t* sumY2;       /* [numBins] units 2^-26 */
  int64_t* sumY3;       /* [numBins] units 2^-22 */
  uint64_t* hist2d;     /* [numBins*logBins], log axis (logMin,logMax) */
} nb200_band_hist;

typedef struct nb200_ctx nb200_ctx;

int nb200_abi_version(void);
/* sizeof() of the ABI structs as this library was compiled, for FFI layout checks:
 * which = 0 nb200_map, 1 nb200_detector, 2 nb200_model, 3 nb200_analysis, 4 nb200_source,
 * 5 nb200_run_consts, 6 nb200_soa_out, 7 nb200_band_hist; -1 for anything else */
int nb200_sizeof(int which);
/* One context per device and host thread.  Device memory: the stage workspace and the deferred-hit queue grow with the
 * largest call so far, up to one internal chunk of 2^24 events (188 B/event: 3.2 GB); a device that cannot give that
 * falls back to smaller chunks, and the environment variable NB200_CHUNK_LOG2 (16..28, read here) sets the size.
 * Results do not depend on the chunk size. */
int nb200_create(int device, nb200_ctx** out);
void nb200_destroy(nb200_ctx* ctx);
const char* nb200_last_error(const nb200_ctx* ctx);
/* cudaStream_t as void*; NULL = the context's own stream */
int nb200_set_stream(nb200_ctx* ctx, void* cuda_stream);
int nb200_synchronize(nb200_ctx* ctx);

/* Host-only: replaces NESTcalc::SetDensity + WorkFunction + SetDriftVelocity +
 * CalculateG2(0) + FindNewMean as called at execNEST.cpp:612-631,878-890. */
/* Sets det->inGas as SetDensity does (NEST.cpp:2270-2276, execNEST.cpp:618-619).
 * inField == -1: vD_middle is read from the SetDriftVelocity_NonUniform table at centralZ
 * (execNEST.cpp:635-637, 882-883) built with z_step (0 => analysis.hh's 0.1 mm). */
int nb200_prepare(nb200_detector* det, double inField, double z_step, nb200_run_consts* out);

/* Host-only: TestSpectra::WIMP_prep_spectrum (TestSpectra.cpp:392-454).  The reference
 * draws a random Xe isotope inside every WIMP_dRate call (:278); those draws come from the
 * Philox host stream keyed by seed.  base/exponent: caller arrays of 100 doubles. */
int nb200_wimp_prep(double mass_GeV, double eStep, double dayNum, uint64_t seed, double* base,
                    double* exponent, double* integral, double* xMax, double* divisor);
/* The same with the isotope of every WIMP_dRate call given by the caller, in call order: the table's points first
 * (int(100/eStep) + 1, or int(1000/eStep) + 1 below 10 GeV; fewer if the spectrum ends in 100 zeros), then the
 * 1e6 terms of the integral.  With the reference's own draws (RandomGen::SelectRanXeAtom under the reference's seed)
 * the table equals the reference's to rounding.  NB200_EINVAL when the list is too short. */
int nb200_wimp_prep_iso(double mass_GeV, double eStep, double dayNum, const int32_t* isotopes, size_t n_isotopes,
                        double* base, double* exponent, double* integral, double* xMax, double* divisor);
/* Host-only: TestSpectra::WIMP_dRate (TestSpectra.cpp:251-390) [events/kg/day/keV] with the Xe isotope the
 * reference draws at :278 made an argument; *ok = 0 where the reference throws ("velocity integral broke"). */
double nb200_wimp_drate(double ER_keV, double mass_GeV, double dayNum, int isotope_A, int* ok);
/* Poisson draw for the WIMP/8B event count (execNEST.cpp:482-488), host stream keyed by seed */
uint64_t nb200_poisson(double mean, uint64_t seed);

/* Deterministic mean yields on the device: replaces NESTcalc::GetYields
 * (NEST.cpp:1211-1252) + YieldResultValidity (:1254-1271) for n independent
 * (E, field) pairs.  out columns [6][n] = PhotonYield, ElectronYield,
 * ExcitonRatio, Lindhard, ElectricField, DeltaT_Scint (NEST.hh:162-169).
 * massNum must be non-zero here (no RNG on this entry point). E/field/out are
 * pointers of kind mem_kind. */
int nb200_yields(nb200_ctx* ctx, const nb200_detector* det, const nb200_model* model,
                 int species, size_t n, const double* E, const double* field,
                 double density, int mem_kind, double* out);

/* Fluctuated quanta on the device: replaces NESTcalc::GetQuanta (NEST.cpp:248-432) for n independent yield records --
 * the stage call behind FullCalculation (NEST.cpp:17-40) and bareNEST.cpp:186.  yields [6][n] as nb200_yields writes
 * them; quanta [4][n] int32 = photons, electrons, ions, excitons; prob_var [2][n] = recombProb, Variance
 * (QuantaResult, NEST.hh:171-178).  model->NRERWidthsParam / n_widths / SkewnessER are the call's NRERWidthsParam and
 * SkewnessER arguments.  Record i draws from Philox(seed; event first_event + i): results do not depend on how a
 * list is split over calls.  NB200_EPHYSICS where the reference throws (fewer than 13 width parameters, quanta not
 * conserved).  Pointers are of kind mem_kind. */
int nb200_quanta(nb200_ctx* ctx, const nb200_detector* det, const nb200_model* model, size_t n, const double* yields,
                 double density, uint64_t seed, uint64_t first_event, int mem_kind, int32_t* quanta, double* prob_var);

/* S2 on the device as a stage call: replaces NESTcalc::GetS2, S2CalculationMode::Full (NEST.cpp:1724-1776, 1975-2063;
 * bareNEST.cpp:210) for n independent records.  Ne [n] int32 = electrons leaving the interaction site; pos [5][n] =
 * truth x, y, z and smeared x, y [mm]; dt [n] drift time [us]; field [n] drift field [V/cm] (below 1 V/cm, or a g2
 * parameter <= 0, gives nine zeros as in the reference, :1746-1750); rc = nb200_prepare's constants (g2_params).
 * out [9][n] = ionization[0..8]: Nee, Nph, nHits, Nphe, area, areaC, phd, phdC (bottom-array values for s2_thr < 0;
 * negated below |s2_thr|), g2.  Record i draws from Philox(seed; event first_event + i). */
int nb200_s2(nb200_ctx* ctx, const nb200_detector* det, const nb200_run_consts* rc, size_t n, const int32_t* Ne,
             const double* pos, const double* dt, const double* field, uint64_t seed, uint64_t first_event,
             int mem_kind, double* out);

/* S1 on the device as a stage call: replaces NESTcalc::GetS1 (NEST.cpp:1288-1722; bareNEST.cpp:203) for n independent
 * records in the Full, Parametric or Hybrid mode (s1mode = NB200_S1_*; Hybrid: Full below 100 keV).  Nph [n] int32 =
 * quanta.photons; pos [6][n] = truth x, y, z and smeared x, y, z [mm]; energy [n] keV (the Hybrid switch); rc =
 * nb200_prepare's constants (vD_middle fixes the map centre z_c).  out [9][n] = scintillation[0..8]: nHits, Nphe,
 * area, areaC, phd, phdC, spike, spikeC, spike count (sign conventions of :1689-1716).  Three launches per 4.2e6
 * records: prelude, the hit kernel of the chain, epilogue.  Record i draws from Philox(seed; event first_event + i). */
int nb200_s1(nb200_ctx* ctx, const nb200_detector* det, const nb200_run_consts* rc, int s1mode, size_t n,
             const int32_t* Nph, const double* pos, const double* energy, uint64_t seed, uint64_t first_event,
             int mem_kind, double* out);

/* The width model of GetQuanta as a call of its own: replaces NESTcalc::RecombOmegaNR (NEST.cpp:164-171),
 * RecombOmegaER (:173-225) and FanoER (:227-246) for n records (efield [V/cm], elecFrac, numQuanta).
 * out [3][n] = omega_NR, omega_ER, Fano_ER.  NB200_EPHYSICS where the reference throws (fewer than 13 width
 * parameters; NRERWidthsParam[9] > [8]). */
int nb200_widths(nb200_ctx* ctx, const nb200_detector* det, const nb200_model* model, size_t n, const double* efield,
                 const double* elecFrac, const double* numQuanta, double density, int mem_kind, double* out);

/* Replaces NESTcalc::GetSpike (NEST.cpp:2243-2268; NEST.hh:345) for n records: pos [3][n] = the dx, dy, dz the
 * reference takes [mm], scint [9][n] = the scintillation[] GetS1 returned; rc->vD_middle is the reference's dS_mid.
 * out [2][n] = smeared spike count, and the same corrected to the detector centre.  Record i draws from
 * Philox(seed; event first_event + i). */
int nb200_spike(nb200_ctx* ctx, const nb200_detector* det, const nb200_run_consts* rc, size_t n, const double* pos,
                const double* scint, uint64_t seed, uint64_t first_event, int mem_kind, double* out);

/* Replaces NESTcalc::xyResolution (NEST.cpp:2699-2736; NEST.hh:404) for n records: true x, y [mm] and the S2
 * size A_top [phd] -> out [2][n] smeared x, y.  Record i draws from Philox(seed; event first_event + i). */
int nb200_xy_resolution(nb200_ctx* ctx, const nb200_detector* det, size_t n, const double* x, const double* y,
                        const double* A_top, uint64_t seed, uint64_t first_event, int mem_kind, double* out);

/* The energy samplers alone: replaces TestSpectra::CH3T_spectrum / C14_spectrum / B8_spectrum / PowLawFit_spectrum
 * (AmBe) / Cf_spectrum / DD_spectrum / ppSolar_spectrum / atmNu_spectrum / WIMP_spectrum (src/TestSpectra.cpp:31-499,
 * include/NEST/TestSpectra.hh:50-69) and the flat / Gauss / exponential defaults of execNEST.cpp:769-781 for n
 * draws: src->kind, eMin, eMax, par[] (and the WIMP table) as for nb200_run; out [n] keV.  Draw i is the energy
 * event first_event + i of an nb200_run with the same source and seed would get. */
int nb200_spectrum(nb200_ctx* ctx, const nb200_source* src, size_t n, uint64_t seed, uint64_t first_event,
                   int mem_kind, double* out);

/* The per-event chain: replaces the loop body of execNEST() (execNEST.cpp:676-1414)
 * and of runNESTvec() (:357-406) for event ids [first_event, first_event+n).
 * soa and hist may each be NULL.  hist is ACCUMULATED into (caller zeroes it). */
int nb200_run(nb200_ctx* ctx, const nb200_detector* det, const nb200_model* model,
              const nb200_analysis* ana, const nb200_source* src,
              const nb200_run_consts* rc, uint64_t seed, uint64_t first_event,
              uint64_t n_events, nb200_soa_out* soa, nb200_band_hist* hist);

/* Batched parameter sweep (BASELINE config 4): n_points grid points -- one detector, model, source and set of
 * run constants each (arrays of n_points), one shared analysis -- every point simulating event ids
 * [first_event, first_event + n_events) under `seed`, one band accumulator per point (bands[n_points],
 * ACCUMULATED into).  Replaces a shell loop of execNEST processes over a parameter grid (the reference's
 * loopNEST mode, src/execNEST.cpp:103-159 driven by examples/loopNEST.in:10-54): the grid-point index is part
 * of the launch, the per-point parameter blocks live in device memory, and blocks of every point share the
 * GPU, so points of 1e4..1e6 events run at the rate of one large run instead of being launch- and tail-bound.
 * A point's band is bit for bit the band of nb200_run with that point's parameters and the same seed / event
 * range.  List sources and vec_mode are not accepted here. */
int nb200_run_sweep(nb200_ctx* ctx, size_t n_points, const nb200_detector* dets, const nb200_model* models,
                    const nb200_analysis* ana, const nb200_source* srcs, const nb200_run_consts* rcs, uint64_t seed,
                    uint64_t first_event, uint64_t n_events, nb200_band_hist* bands);

/* Band post-processing for the bands of the LAST nb200_run_sweep on this context, on the device (they are still in
 * device memory): per grid point GetBand's statistics (execNEST.cpp:1582-1618; stats [n_points][numBins][7] as
 * nb200_band_finalize), rootNEST's goodness of fit against target_band [numBins][7] (examples/rootNEST.cpp:600-643;
 * gof [n_points][6] as nb200_band_chi2) and the events below the line centroid[numBins] (:880-915; leak
 * [n_points][3] = below, all, ratio, as nb200_band_leakage's total).  Any output may be NULL; all pointers HOST.
 * What a fit over the grid reads back is then 6 numbers per point instead of a 28 KB accumulator. */
int nb200_sweep_band_post(nb200_ctx* ctx, size_t n_points, const nb200_analysis* ana, const double* target_band,
                          int free_param, const double* centroid, double* stats, double* gof, double* leak);

/* Per-bin band fits on the device (k_band_fit, one thread per grid point and S1 bin), for the bands of
 * the LAST nb200_run_sweep on this context (nb200_sweep_band_fit) or for the context's own device accumulator
 * (nb200_band_fit_device, after nb200_run_async):
 *   model 0 -- GetBand_Gaussian (examples/rootNEST.cpp:1309-1359): fit [n_points][numBins][8], each row exactly what
 *     nb200_band_fit_gauss writes for that slice (the same statements run on the device);
 *   model 1 -- the skew-Gaussian band (:1360-1507): "skewband" (:1361-1364) fitted to the slice's row over the log bins
 *     whose centres lie in mean +- 3 sigma (GetBand's moments; widened by 0.5 as :1322-1327 when the window reaches
 *     beyond 2 / 3), from EstimateSkew's start values (:1567-1585, :1367-1374): fit [n_points][numBins][12] = {A, xi,
 *     omega, alpha, xi error, omega error, alpha error, cov(xi, omega), cov(xi, alpha), cov(omega, alpha), chi^2 per
 *     degree of freedom, bins used} -- what nb200_hist_fit_cov(1, ...) gives for that window and start.
 * Rows without a fit (too few filled bins, no minimum) are zeros.  line [numBins] (or NULL) with leak [n_points][numBins]
 * (or NULL): the fraction below the line read off each slice's fit -- (1 - erf(n_sigma / sqrt 2)) / 2 (:848) or the
 * skew-normal CDF with the reference's Owen's-T quadrature (:700-735, 1587-1602), as nb200_band_leakage_gauss /
 * nb200_band_leakage_skew.  All pointers HOST.  A fit over a 250-point grid is one launch of 24 500 threads instead of
 * 250 accumulators copied back and fitted on the host. */
int nb200_sweep_band_fit(nb200_ctx* ctx, size_t n_points, const nb200_analysis* ana, int model, const double* line,
                         double* fit, double* leak);
int nb200_band_fit_device(nb200_ctx* ctx, const nb200_analysis* ana, int model, const double* line, double* fit,
                          double* leak);

/* Device-resident variant used for throughput measurement: the band accumulator
 * stays in device memory owned by ctx between calls; fetch adds it into hist. */
int nb200_run_async(nb200_ctx* ctx, const nb200_detector* det, const nb200_model* model,
                    const nb200_analysis* ana, const nb200_source* src,
                    const nb200_run_consts* rc, uint64_t seed, uint64_t first_event,
                    uint64_t n_events, nb200_soa_out* soa_device_or_null);
int nb200_band_reset(nb200_ctx* ctx, const nb200_analysis* ana);
int nb200_band_fetch(nb200_ctx* ctx, nb200_band_hist* hist);
/* Use a caller-owned device buffer of nb200_band_hist_bytes(ana) bytes as the accumulator
 * (e.g. a torch tensor that torch.distributed all-reduces); the caller zeroes it. */
int nb200_band_attach(nb200_ctx* ctx, const nb200_analysis* ana, void* device_buffer);
/* device pointer + element count (uint64) of the packed accumulator, for an
 * external all-reduce (torch.distributed / NCCL) over the event shards */
int nb200_band_device_buffer(nb200_ctx* ctx, void** dptr, size_t* n_u64);
/* ncclComm_t as void*; sums the device accumulator over ranks in place */
int nb200_allreduce_hist(nb200_ctx* ctx, void* nccl_comm);

/* GetBand's final statistics (execNEST.cpp:1582-1618) from the accumulator:
 * band[j][0..6] = centre, mean S1, mean y, sigma (N-1), sigma/sqrt(N), eff, skew */
size_t nb200_band_hist_bytes(const nb200_analysis* ana);
double nb200_band_scale_x(const nb200_analysis* ana);
int nb200_band_finalize(const nb200_analysis* ana, const nb200_band_hist* hist,
                        double* band /* [numBins][7] */);

/* Band post-processing on the host, the two uses rootNEST makes of a band without ROOT fits:
 *
 * nb200_band_chi2 -- goodness of fit between two bands (examples/rootNEST.cpp:600-643): a = the simulated
 * band, b = the band compared against, both [numBins][7] as nb200_band_finalize writes them.  Bins with
 * |mean| >= 5 are ignored for useS2 == 0 as in the reference; the error of a width, which rootNEST reads
 * from its Gaussian fits, is sigma/sqrt(2N) = band[.][4]/sqrt(2); empty bins are skipped.
 * DoF = numBins - |free_param|.  out[6] = reduced chi^2 of the means, of the widths, their arithmetic and
 * geometric means, and the mean %-differences (a - b)/b of means and widths.  NB200_EINVAL when the bin
 * centres differ by more than 0.05 (the reference's "Binning doesn't match for GoF calculation").
 *
 * nb200_band_leakage -- fraction of the accumulator's events below a centroid, by counting
 * (examples/rootNEST.cpp:880-915 with the per-bin centroid, NRbandCenter = 1): centroid[numBins] in the
 * band's y units; counted on the 2-D histogram (the log bin holding the centroid is split linearly, so the
 * resolution is (logMax - logMin)/logBins); events the reference would have pushed as y = 0 count as
 * below.  leak[numBins]; err_hi / err_lo[numBins] = the reference's Poisson error bars (:898-903) or NULL;
 * total[3] = {events below, all events, ratio} or NULL. */
int nb200_band_chi2(int numBins, int useS2, const double* a, const double* b, int free_param, double* out);
int nb200_band_leakage(const nb200_analysis* ana, const nb200_band_hist* hist, const double* centroid,
                       double* leak, double* err_hi, double* err_lo, double* total);

/* nb200_band_fit_gauss -- GetBand_Gaussian without ROOT (examples/rootNEST.cpp:1309-1359, skewness == 0): a
 * Gaussian A exp(-(y - mean)^2 / (2 sigma^2)) fitted to each S1 bin's row of the 2-D histogram as TH1::Fit
 * does by default (function at the bin centres, weights 1/content, empty bins left out; Levenberg-Marquardt,
 * errors from the inverse of J^T W J).  fit[numBins][8] = {A, mean, sigma, mean error, sigma error, chi^2 per
 * degree of freedom of that fit, the reference's own goodness figure band[j][8] (:1346-1357, in its indexing:
 * ROOT bin k with k = 0 the underflow, model at logMin + k w), points used}; a row of zeros where fewer than
 * four log bins are filled.  Needs logBins >= 4; the resolution is the caller's choice of logBins.
 *
 * nb200_band_leakage_gauss -- the "Leak Frac Fit" columns of rootNEST's leakage table (:700-738, :848, :908-914):
 * er, nr = [numBins][8] as nb200_band_fit_gauss writes them, num_sig_above = NUMSIGABV (:39; 0 in the reference).
 * out[numBins][4] = {number of ER sigmas between the two means, leakage (1 - erf(n/sqrt 2))/2, + error bar,
 * - error bar}; discrimination = 1 - leakage. */
int nb200_band_fit_gauss(const nb200_analysis* ana, const nb200_band_hist* hist, double* fit);

/* nb200_hist_fit -- one histogram, one chi^2 fit (same minimiser and conventions as nb200_band_fit_gauss), for callers
 * that keep their own histograms: rootNEST's skew-Gaussian mode re-bins every S1 slice on mean +- 3 sigma
 * (examples/rootNEST.cpp:1319-1330).  model 0 = A exp(-(x-mu)^2/(2 s^2)), par/err/start[3] = {A, mu, s};
 * model 1 = rootNEST's "skewband" (:1361-1364) A/(w sqrt(2 pi)) exp(-(x-xi)^2/(2 w^2)) (1 + erf(al (x-xi)/(w sqrt 2))),
 * par/err/start[4] = {A, xi, w, al}.  counts[bins] on (lo, hi).  NB200_EFIT when no minimum was found.
 *
 * nb200_band_leakage_skew -- leakage below line[i] of a skew-Gaussian band (xi, omega, alpha per S1 bin), i.e. the
 * skew-normal CDF with Owen's T evaluated by the reference's own quadrature (:700-735, :1587-1602); err_hi / err_lo
 * (or NULL) = the same with omega +- omega_err, the reference's "quick + dirty" error bars (:739-771). */
int nb200_hist_fit(int model, const uint64_t* counts, int bins, double lo, double hi, const double* start, double* par,
                   double* err, double* chi2_ndf);
/* nb200_hist_fit_cov -- nb200_hist_fit returning the parameters' covariance matrix (J^T W J)^-1, row-major
 * cov[npar][npar] (3 or 4 squared), as rootNEST's skewness = 2 reads it from TFitResult (:1399-1405).
 *
 * nb200_band_leakage_skew_cov -- the skewness = 2 leakage (:774-845): skew-normal CDF at line[i] and its 1-sigma error
 * propagated from the fit (skew [numBins][9] = {xi, xi error, omega, omega error, alpha, alpha error, cov(xi, omega),
 * cov(xi, alpha), cov(omega, alpha)}) and from the line's own error line_err[i]; the reference's error bars are
 * leak +- err. */
int nb200_hist_fit_cov(int model, const uint64_t* counts, int bins, double lo, double hi, const double* start,
                       double* par, double* cov, double* chi2_ndf);
int nb200_band_leakage_skew_cov(int numBins, const double* skew, const double* line, const double* line_err,
                                double* leak, double* err);

/* nb200_graph_fit -- TGraph::Fit without ROOT for the two curves rootNEST puts through the NR band's means (:855-878):
 * model 2 = the Woods function p0/(x + p1) + p2 x + p3, model 3 = its fallback p0 + (p1 - p0)/(1 + (x/p2)^p3);
 * unweighted least squares over (x[i], y[i]), par/start[4]; *chi2_ndf = sum of squared residuals / (n - 4), the figure
 * the reference tests against 1.5 and 2.  NB200_EFIT when no minimum was found. */
int nb200_graph_fit(int model, int n, const double* x, const double* y, const double* start, double* par, double* chi2_ndf);
int nb200_band_leakage_skew(int numBins, const double* xi, const double* omega, const double* omega_err,
                            const double* alpha, const double* line, double* leak, double* err_hi, double* err_lo);
int nb200_band_leakage_gauss(int numBins, const double* er, const double* nr, double num_sig_above, double* out);

/* LAr: replaces LArNEST::FullCalculation (src/LArNEST.cpp:17-33) for species
 * NR=0, ER=1, Alpha=2 (LArParameters.hh:27-34).  yields [9][n] in the field order
 * of LArYieldResult (LArNEST.hh:19-29), fluct [4][n] LArYieldFluctuationResult
 * (:31-36); either may be NULL.  Event i uses Philox stream (seed, first_event+i). */
int nb200_lar(nb200_ctx* ctx, int species, size_t n, const double* E, const double* field,
              double density, uint64_t seed, uint64_t first_event, int mem_kind,
              double* yields, double* fluct);

/* Per-stage device time of the chain kernels (k_pre, k_hits, k_post): with timing enabled each
 * launch is bracketed by CUDA events on the context's stream; nb200_timing_read synchronises
 * the stream, returns the milliseconds and launch counts accumulated since the last read and
 * clears them.  ms[3], launches[3]. */
int nb200_timing_enable(nb200_ctx* ctx, int on);
int nb200_timing_read(nb200_ctx* ctx, double* ms, uint64_t* launches);

/* Single-electron size Monte Carlo of CalculateG2(verbosity > 0) (NEST.cpp:2169-2225): n single
 * electrons at random (r, phi); returns the mean and the width about the analytic SE the banner
 * "g1 = ... SE_mean,width = ..." prints. */
int nb200_se_stats(nb200_ctx* ctx, const nb200_detector* det, const nb200_run_consts* rc, int n, uint64_t seed,
                   double* se_mean, double* se_width);

/* Measured FP64 pipe peak of this device: a register-resident DFMA chain kernel timed with
 * CUDA events (the roofline denominator of the chain kernel). */
int nb200_fp64_peak(nb200_ctx* ctx, double* tflops);

/* Test hook: the S1 hit loop builds -ln(u) and (sin, cos) straight from random bits with
 * short polynomials instead of libm's generic log/sincospi.  bits3 = n triples {hi, lo, z}
 * (u = (hi:lo)/2^64, angle = 2 pi (z+1/2)/2^32), out3 = n triples {-ln u, sin, cos}; host
 * pointers.  Used by tests/test_gpu_hitmath.py to hold them to 1e-15 against libm. */
int nb200_debug_hitmath(nb200_ctx* ctx, size_t n, const uint32_t* bits3, double* out3);

/* Test hook: the standard normals of the S1 hit loop (2048-layer ziggurat, replacing the
 * Box-Muller of RandomGen::rand_gauss, src/RandomGen.cpp:44-53, for the per-hit draws of
 * NEST.cpp:1367-1429), drawn by the very device function the hit kernel and its finisher call: slot s =
 * the three photo-electrons of Philox block s % 1024 of event first_event + s / 1024.  out = 3*slots doubles,
 * kind = 3*slots bytes or NULL (0: layer core, 1: wedge / tail completion); host pointers.
 * tests/test_gpu_samplers.py holds them to N(0,1) (KS, moments, tail mass). */
int nb200_debug_normal(nb200_ctx* ctx, uint64_t seed, uint64_t first_event, size_t slots, double* out,
                       unsigned char* kind);

/* Test hook: `count` draws of the chain's binomial sampler (the device replacement of
 * RandomGen::binom_draw, src/RandomGen.cpp:132-134) with n trials of probability p, draw i on
 * the stream of event first_event + i; out = count int64 on the host.
 * tests/test_gpu_samplers.py compares the histogram with the exact pmf. */
int nb200_debug_binomial(nb200_ctx* ctx, uint64_t seed, uint64_t first_event, size_t count, int64_t n,
                         double p, int64_t* out);

/* Test hook: the same sampler with a trial count per draw (n[count], HOST), on the binomial stream of a given caller:
 * which_stream 0 = the chain's S1 hit count, 1 = LArNEST's fluctuations.  k_lar makes the latter's draws regrouped
 * (queued per warp, one candidate per lane and pass); tests/test_gpu_lar.py requires them to equal these. */
int nb200_debug_binomial_list(nb200_ctx* ctx, uint64_t seed, uint64_t first_event, size_t count, const int64_t* n,
                              double p, int which_stream, int64_t* out);

/* Test hook: the same for the alias-table sampler that draws an event's number of double-phe hits,
 * Binomial(nHits, P_dphe) (the hit-by-hit decision of NEST.cpp:1391, taken once per event; also the
 * second photo-electrons of S2, NEST.cpp:1984): 0 < p < 1, 0 <= n < 65536. */
int nb200_debug_binomial_fixed(nb200_ctx* ctx, uint64_t seed, uint64_t first_event, size_t count, int n,
                               double p, int64_t* out);

/* Host-only test hooks (no device needed; tests/test_capi_cpu.py): the ziggurat layers
 * (x[n+1], f[n+1], k[n], w[n], n_layers must equal the library's 2048) and the alias tables of
 * Binomial(2^b, p), b = 0..12 (8204 entries: table b at offset 2^b - 1 + b). */
int nb200_debug_zig_tables(int n_layers, double* x, double* f, uint32_t* k, double* w);
int nb200_debug_alias_tables(double p, int n_entries, uint32_t* thr, uint32_t* alias);
/* Host-only: NESTcalc::GetDensity (NEST.cpp:2278-2350): xenon density [g/cm^3] at (Kelvin, bara); *in_gas in/out as the
 * reference's inGas (liquid, solid below 161.4 K, or the real-gas law; the gas phase is reported here even though the
 * event chain refuses gas-phase detectors). */
double nb200_density(double Kelvin, double bara, double molarMass, int* in_gas);

/* Host-only: NESTcalc::GetDriftVelocity_Liquid (NEST.cpp:2366-2521): electron drift speed [mm/us] in liquid xenon at
 * (Kelvin, field [V/cm]); NaN outside the tabulated 100-230 K. */
double nb200_drift_velocity_liquid(double Kelvin, double field);

/* Host-only: NESTcalc::SetDriftVelocity_NonUniform (NEST.cpp:2664-2697, liquid): the table of effective drift speeds
 * [mm/us] from z = k z_step (k = 0, 1, ... while z < TopDrift) to the liquid surface along the line (x, y), through the
 * detector's FitEF map -- each entry the same sum, in the same order, as the reference's O(n^2) loop.  Call with
 * table == NULL to get the length in *n; then with a buffer of that many doubles. */
int nb200_drift_table(const nb200_detector* det, double z_step, double x, double y, double* table, size_t* n);

/* Host-only test hook: the flattened position maps exactly as the kernels evaluate them (closed form or bilinear
 * (r, z) table), at n points: s1 = FitS1(x, y, z), s2 = FitS2(x, y), ef = FitEF(x, y, z), tba = FitTBA(x, y, z)[1]
 * (VDetector.hh:131-224); any output may be NULL.  Tables must be host-resident (as nb200_flatten leaves them). */
int nb200_debug_maps(const nb200_detector* det, size_t n, const double* x, const double* y, const double* z, double* s1,
                     double* s2, double* ef, double* tba);

/* counters for the bench: kernels launched by this context since creation */
uint64_t nb200_launch_count(const nb200_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* NEST_B200_H */
