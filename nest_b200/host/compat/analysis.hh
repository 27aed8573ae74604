// Compatibility header for include/NEST/analysis.hh:7-77: the reference configures an analysis through file-scope
// globals defined in this header (one translation unit includes it).  Same names; the values are the defaults of
// NEST::Analysis (nest_b200/host/NEST_b200.hh), which are analysis.hh's.  Programs using the mirror's own drivers set
// nest_b200_cli::analysis instead.
#pragma once
#include "NEST.hh"

namespace nb200_compat {
inline const NEST::Analysis& A() {
  static const NEST::Analysis a{};
  return a;
}
}  // namespace nb200_compat

int verbosity = nb200_compat::A().verbosity;
unsigned loopNEST = 0;
short PrintSubThr = nb200_compat::A().PrintSubThr;
bool MCtruthE = nb200_compat::A().MCtruthE;
bool MCtruthPos = nb200_compat::A().MCtruthPos;
NEST::S1CalculationMode s1CalculationMode = nb200_compat::A().s1CalculationMode;
NEST::S2CalculationMode s2CalculationMode = nb200_compat::A().s2CalculationMode;
int usePD = nb200_compat::A().usePD;
int useS2 = nb200_compat::A().useS2;
double minS1 = nb200_compat::A().minS1;
double maxS1 = nb200_compat::A().maxS1;
int numBins = nb200_compat::A().numBins;
double minS2 = nb200_compat::A().minS2;
double maxS2 = nb200_compat::A().maxS2;
double logMax = nb200_compat::A().logMax;
double logMin = nb200_compat::A().logMin;
int logBins = nb200_compat::A().logBins;
double z_step = nb200_compat::A().z_step;
double E_step = nb200_compat::A().E_step;
int freeParam = 2;
int skewness = 1;
int mode = 0;
int NRbandCenter = 3;
