// execNEST command line on the GPU chain: same argv as the reference (src/execNEST.cpp:40-256),
// LUX_Run03 detector and analysis.hh settings by default (README.md:170-219 of the reference);
// NB200_ANALYSIS=G2 selects the analysis_G2.hh settings; NB200_USEPD, NB200_USES2, NB200_MCTRUTHE, NB200_MCTRUTHPOS,
// NB200_MINS1 / MAXS1 / MINS2 / MAXS2 / LOGMIN / LOGMAX set single switches of analysis.hh.
#include <cstdlib>
#include <cstring>

#include "LUX_Run03_b200.hh"

int main(int argc, char** argv) {
  if (const char* a = std::getenv("NB200_ANALYSIS"))
    if (!std::strcmp(a, "G2")) nest_b200_cli::analysis = NEST::Analysis::G2();
  if (const char* v = std::getenv("NB200_VERBOSITY")) nest_b200_cli::analysis.verbosity = std::atoi(v);
  // the switches a user of the reference edits in analysis.hh before compiling (include/NEST/analysis.hh:16-48)
  {
    NEST::Analysis& A = nest_b200_cli::analysis;
    if (const char* v = std::getenv("NB200_PRINTSUBTHR")) A.PrintSubThr = short(std::atoi(v));
    if (const char* v = std::getenv("NB200_USEPD")) A.usePD = std::atoi(v);
    if (const char* v = std::getenv("NB200_USES2")) A.useS2 = std::atoi(v);
    if (const char* v = std::getenv("NB200_MCTRUTHE")) A.MCtruthE = std::atoi(v) != 0;
    if (const char* v = std::getenv("NB200_MCTRUTHPOS")) A.MCtruthPos = std::atoi(v) != 0;
    if (const char* v = std::getenv("NB200_MINS1")) A.minS1 = std::atof(v);
    if (const char* v = std::getenv("NB200_MAXS1")) A.maxS1 = std::atof(v);
    if (const char* v = std::getenv("NB200_MINS2")) A.minS2 = std::atof(v);
    if (const char* v = std::getenv("NB200_MAXS2")) A.maxS2 = std::atof(v);
    if (const char* v = std::getenv("NB200_LOGMIN")) A.logMin = std::atof(v);
    if (const char* v = std::getenv("NB200_LOGMAX")) A.logMax = std::atof(v);
    if (const char* v = std::getenv("NB200_NUMBINS")) A.numBins = std::atoi(v);
    if (const char* v = std::getenv("NB200_S1MODE")) A.s1CalculationMode = NEST::S1CalculationMode(std::atoi(v));  // Full, Parametric, Hybrid
  }
  auto* detector = new DetectorExample_LUX_RUN03();
  int rc = nest_b200_cli::main_like(detector, argc, argv);
  delete detector;
  return rc;
}
