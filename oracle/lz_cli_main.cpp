// TEST INFRASTRUCTURE.  The execNEST command line of nest-b200 (nest_b200/host) with the reference's OWN
// detector header include/Detectors/LZ_WS2024.hh -- the detector the shipped execNEST instantiates
// (src/execNEST.cpp:43) -- compiled against the VDetector mirror: demonstrates that an unmodified user
// detector header recompiles against nest_b200/host/NEST_b200.hh and runs on the GPU through (r,z)
// lookup tables built by include/nest_b200_flatten.hh.  Built by oracle/Makefile into
// oracle/_ref/execNEST_b200_lz only where the reference tree is present; compared with
// oracle/_ref/execNEST_ref (the reference as shipped) by tests/test_gpu_cli.py.
#include <cstdlib>
#include <cstring>

#include "NEST_b200.hh"
#define VDetector_hh 1  // the reference header's own VDetector.hh include becomes a no-op: the mirror stands in
#include "LZ_WS2024.hh"

int main(int argc, char** argv) {
  nest_b200_cli::analysis = NEST::Analysis::G2();  // the shipped binary is compiled with analysis_G2.hh
  if (const char* v = std::getenv("NB200_VERBOSITY")) nest_b200_cli::analysis.verbosity = std::atoi(v);
  auto* detector = new LZ_Detector_2024();
  int rc = nest_b200_cli::main_like(detector, argc, argv);
  delete detector;
  return rc;
}
