// TEST INFRASTRUCTURE ONLY (oracle build).  Like oracle/shim_lux/analysis_G2.hh -- the analysis.hh settings for the
// unmodified src/execNEST.cpp -- but the switches a user edits in analysis.hh before compiling (usePD, useS2, MCtruthE,
// MCtruthPos, the S1 / S2 windows, the band's axis) can be given in the environment instead: they are plain mutable
// globals there, set here before main() runs.  No reference file is edited; with no variable set the binary is
// execNEST_ref_lux.  Used to pin the analysis modes other than the two shipped ones against the reference itself
// (tests/test_gpu_analysis_modes.py).
#include <cstdlib>

#include "analysis.hh"

namespace {
struct nb200_ref_analysis_from_env {
  static void i(const char* n, int& v) {
    if (const char* s = std::getenv(n)) v = std::atoi(s);
  }
  static void b(const char* n, bool& v) {
    if (const char* s = std::getenv(n)) v = std::atoi(s) != 0;
  }
  static void d(const char* n, double& v) {
    if (const char* s = std::getenv(n)) v = std::atof(s);
  }
  nb200_ref_analysis_from_env() {
    i("NEST_REF_VERBOSITY", verbosity);
    if (const char* s = std::getenv("NEST_REF_PRINTSUBTHR")) PrintSubThr = short(std::atoi(s));
    i("NEST_REF_USEPD", usePD);
    i("NEST_REF_USES2", useS2);
    b("NEST_REF_MCTRUTHE", MCtruthE);
    b("NEST_REF_MCTRUTHPOS", MCtruthPos);
    d("NEST_REF_MINS1", minS1);
    d("NEST_REF_MAXS1", maxS1);
    d("NEST_REF_MINS2", minS2);
    d("NEST_REF_MAXS2", maxS2);
    d("NEST_REF_LOGMIN", logMin);
    d("NEST_REF_LOGMAX", logMax);
    i("NEST_REF_NUMBINS", numBins);
    if (const char* s = std::getenv("NEST_REF_S1MODE")) s1CalculationMode = NEST::S1CalculationMode(std::atoi(s));
  }
} nb200_ref_analysis_from_env_instance;
}  // namespace
