// execNEST command line on the GPU chain: same argv as the reference (src/execNEST.cpp:40-256),
// LUX_Run03 detector and analysis.hh settings by default (README.md:170-219 of the reference);
// NB200_ANALYSIS=G2 selects the analysis_G2.hh settings.
#include <cstdlib>
#include <cstring>

#include "LUX_Run03_b200.hh"

int main(int argc, char** argv) {
  if (const char* a = std::getenv("NB200_ANALYSIS"))
    if (!std::strcmp(a, "G2")) nest_b200_cli::analysis = NEST::Analysis::G2();
  if (const char* v = std::getenv("NB200_VERBOSITY")) nest_b200_cli::analysis.verbosity = std::atoi(v);
  auto* detector = new DetectorExample_LUX_RUN03();
  int rc = nest_b200_cli::main_like(detector, argc, argv);
  delete detector;
  return rc;
}
