// TEST INFRASTRUCTURE.  oracle/lz_cli_main.cpp for any detector header the reference ships: the execNEST command
// line of nest-b200 (nest_b200/host) with the reference's OWN, unmodified include/Detectors/<header> compiled against
// the VDetector mirror (its position maps become (r,z) lookup tables, include/nest_b200_flatten.hh).  Built by
// oracle/Makefile as oracle/_ref/execNEST_b200_<name> with -DNB_DET_HEADER='"<header>"' -DNB_DET_CLASS=<class>, next to
// oracle/_ref/execNEST_ref_<name> = the unmodified src/execNEST.cpp instantiating the same class (a two-line shadow of
// LZ_WS2024.hh); compared by tests/test_gpu_analysis_modes.py.
#include <cstdlib>
#include <cstring>

#include "NEST_b200.hh"
#define VDetector_hh 1  // the reference header's own VDetector.hh include becomes a no-op: the mirror stands in
#include NB_DET_HEADER

int main(int argc, char** argv) {
#ifndef NB_DET_LUX_ANALYSIS
  nest_b200_cli::analysis = NEST::Analysis::G2();  // src/execNEST.cpp is compiled with analysis_G2.hh
#endif  // (else the analysis.hh settings, the mirror's default: the reference side is built with oracle/shim_lux_ana)
  if (const char* v = std::getenv("NB200_VERBOSITY")) nest_b200_cli::analysis.verbosity = std::atoi(v);
  auto* detector = new NB_DET_CLASS();
  int rc = nest_b200_cli::main_like(detector, argc, argv);
  delete detector;
  return rc;
}
