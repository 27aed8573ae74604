// TEST INFRASTRUCTURE ONLY (oracle build).  oracle/shim_lux/analysis_G2.hh without its detector shadow: the analysis.hh
// settings (every event printed, band in spike counts) for the LUX-derived user detectors of oracle/shim_variants.
#include "analysis.hh"
