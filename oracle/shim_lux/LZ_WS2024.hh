// TEST INFRASTRUCTURE ONLY (oracle build).  Shadows include/Detectors/LZ_WS2024.hh
// so that the unmodified src/execNEST.cpp:20,43 instantiates the LUX_Run03
// detector that README.md:170-180 and BASELINE.json config 1 describe.
#pragma once
#include "LUX_Run03.hh"
using LZ_Detector_2024 = DetectorExample_LUX_RUN03;
