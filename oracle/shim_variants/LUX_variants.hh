// TEST INFRASTRUCTURE ONLY (oracle build).  Detector-parameter branches of GetS1 / GetS2 / GetQuanta that neither
// shipped LUX_Run03 nor LZ settings reach, as user detectors: each class is the reference's own
// DetectorExample_LUX_RUN03 (include/Detectors/LUX_Run03.hh, unmodified) with a few parameters changed through the
// VDetector setters, the way a user derives a detector.  Compiled twice by oracle/Makefile: into the unmodified
// src/execNEST.cpp (reference side) and into nest-b200's command line against the VDetector mirror.
#pragma once
#include "LUX_Run03.hh"

// bottom-array S2 (s2_thr < 0: NEST.cpp:2010-2049) and a 3-fold S1 coincidence (nCr sums, :1647-1675)
class LUX_BottomS2_3Fold : public DetectorExample_LUX_RUN03 {
 public:
  LUX_BottomS2_3Fold() {
    set_s2_thr(-150.);
    set_coinLevel(3);
  }
};
// single-photo-electron efficiency below one (:1339-1347, 1381-1386, 1403), another threshold and double-phe
// probability, linear and quadratic noise on both signals (:1613-1623, 1994-2003), S1 / S2 baseline noise
class LUX_Noisy : public DetectorExample_LUX_RUN03 {
 public:
  LUX_Noisy() {
    set_sPEeff(0.85);
    set_sPEthr(0.5);
    set_P_dphe(0.25);
    set_noiseLinear(3.e-2, 2.e-2);
    set_noiseQuadratic(1.e-4, 2.e-5);
    set_noiseBaseline(0.02, 0.1, 0.0, 0.0);
  }
};
// no coincidence requirement (coinLevel 0: the Nph floor of execNEST.cpp:1214, GetBand's ReqS1), a short electron
// lifetime, a weaker extraction field, another S2 Fano factor, a wider coincidence window
class LUX_NoCoincidence : public DetectorExample_LUX_RUN03 {
 public:
  LUX_NoCoincidence() {
    set_coinLevel(0);
    set_eLife_us(250.);
    set_E_gas(5.0);
    set_s2Fano(1.5);
    set_coinWind(50.);
  }
};
// another thermodynamic point and work function (density, W, drift speed, extraction: the per-run constants), coarser
// position reconstruction, and a single-fold S1
class LUX_WarmNewW : public DetectorExample_LUX_RUN03 {
 public:
  LUX_WarmNewW() {
    set_T_Kelvin(170.);
    set_p_bar(1.7);
    set_OldW13eV(false);
    set_PosResFlat(5.0);
    set_PosResBase(120.);
    set_coinLevel(1);
  }
};
