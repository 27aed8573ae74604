// TEST INFRASTRUCTURE ONLY (oracle build).  Shadows include/NEST/analysis_G2.hh
// with the analysis.hh settings (usePD=2, useS2=0, 98 bins ...) that go with
// LUX_Run03 (execNEST.cpp:17).
#include "analysis.hh"
